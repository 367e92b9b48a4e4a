"""CPU-only: matlab/cnmfe_b200_mex.cpp built against a FUNCTIONAL stand-in for mex.h (tests/stubs/mex_impl.cpp: malloc-backed
mxArrays) and driven through the command sequence of the drop-in methods (tests/stubs/mex_driver.cpp):
  * linked with a recording fake of the C ABI -> argument marshalling of every gateway command (options struct incl. bg_ssub /
    background_model / nb, CSC conversion, output shapes, handle lifetime incl. the mexAtExit hook, error forwarding);
  * linked with the real libcnmfe_b200.so -> on a box without a GPU 'create' fails loudly through mexErrMsgIdAndTxt.
Also: every function the drop-in .m files call exists as a .m file or a gateway command."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STUBS = os.path.join(ROOT, "tests", "stubs")
MEX = os.path.join(ROOT, "matlab", "cnmfe_b200_mex.cpp")


def _build(tmp_path, extra, name):
    exe = str(tmp_path / name)
    cmd = ["g++", "-std=c++17", "-O1", "-I" + STUBS, "-I" + os.path.join(ROOT, "include"), MEX, os.path.join(STUBS, "mex_impl.cpp"),
           os.path.join(STUBS, "mex_driver.cpp")] + extra + ["-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_gateway_marshalling_against_recording_library(tmp_path):
    exe = _build(tmp_path, [os.path.join(STUBS, "fake_cnmfe.cpp")], "mex_fake")
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    out = r.stdout.splitlines()
    expect = [
        "create d1=6 d2=4 T=40 np=2 pp0=[1 3 1 4] bp1=[2 6 1 4] owned=null rr=2 nn=0 dev=0",
        "upload_block ip=0 dtype=1 first=1234",
        # demo_large_data_1p.m: nnls, bg_ssub = 2, deconv_options = struct('type','ar1','method','foopsi','smin',-5,'optimize_pars',true,'optimize_b',true,'max_tau',100)
        "set_options alg=2 maxIter=5 deconv_flag=1 accel=1 quirk=1 tensor=1 model=0 nb=1 ssub=2 outlier=2.5 | type=1 method=0 optb=1 optp=1 maxIter=10 smin=-5 lambda=0 max_tau=100 tau_range=0",
        "ssub_dims -> d1s=3 d2s=2 nnb=4 r_shift[3]=3 c_shift[3]=-3",
        "ring_offsets -> n=4 r_shift[0]=-2 c_shift[0]=2",
        "set_ring ip=1 W=set b0=set W1=1.5 b00=2.5",
        "set_ring ip=0 W=null b0=set W1=-1 b00=2.5",
        "set_neurons K=2 nnz=3 ir0=3 pr0=1.5 C0=42",
        "get_ring -> W 4x6 W[5]=105 b0 12x1 b0[0]=7.5",
        "set_prev K=2 nnz=3 ir0=3 pr0=1.5 C0=43",
        "set_search K=2 nnz=3 ir0=3 pr0=-1 C0=-1",
        "update_spatial update_sn=1",
        "update_spatial -> n=3 v1=0.5 sn 6x4 sn0=11",
        "post_process_spatial -> sparse=1 pr0=0 pr1=2.5",
        "set_use_c_hat 0",
        "update_temporal -> C 2x40 C0=1 Craw0=2 S0=3 kp 2x2 kp0=0.95 nsn0=0.1",
        "set_bf ip=0 b0=0.5 f1=0.25 b0vec=null",
        "get_bf -> b 12x1 f 1x40 b0 12x1",
        "estimate_noise -> 6x4 sn0=9",
        "compute_rss 1 40 b0=3.25 b0new=4.5",
        "compute_rss -> 2x1 rss1=11",
        "reconstruct_background ip=1 3 12 b0=3.25 b0new=4.5",
        "reconstruct_background -> 12x10 y0=5.5",
        "deconvolve T=40 N=3 type=1 method=0 sn=null pars=null y0=0.125 dev=0",
        "deconvolve -> c 40x3 pars 2x3 lam0=7",
        "rejected: cnmfe:options: background_model 'pca' unknown (ring, svd, nmf)",
        "locks=2", "destroy (1 of 2)", "locks=1", "destroy (2 of 2)",
        "error path: cnmfe:b200: create: cnmfe_create: device 99 out of range",
    ]
    pos = 0
    for line in expect:
        assert line in out[pos:], "missing or out of order: %r\n--- got ---\n%s" % (line, r.stdout)
        pos = out.index(line, pos) + 1


def test_gateway_with_real_library_fails_loudly_without_gpu(tmp_path, built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    libdir = os.path.join(ROOT, "cnmf_e_b200")
    exe = _build(tmp_path, ["-L" + libdir, "-lcnmfe_b200", "-Wl,-rpath," + libdir], "mex_real")
    r = subprocess.run([exe, "real"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "mex error: cnmfe:b200: create:" in r.stdout and "no CUDA device" in r.stdout, r.stdout


def test_dropin_methods_call_only_existing_functions():
    """Every cnmfe_b200_* function the MATLAB side calls is a .m file under matlab/, and every gateway command used exists in
    cnmfe_b200_mex.cpp (round 1 shipped helpers that were only comments)."""
    mdir = os.path.join(ROOT, "matlab")
    files = [os.path.join(dp, f) for dp, _, fs in os.walk(mdir) for f in fs if f.endswith(".m")]
    assert len(files) >= 9
    have = {os.path.splitext(os.path.basename(f))[0] for f in files}
    commands = set(re.findall(r'c == "([a-z_]+)"', open(MEX).read()))
    for f in files:
        code = "\n".join(l.split("%")[0] for l in open(f).read().splitlines())          # strip comments
        assert code.strip(), "%s has no code" % f
        for fn in set(re.findall(r"\b(cnmfe_b200_[A-Za-z0-9_]+)\s*\(", code)):
            assert fn == "cnmfe_b200_mex" or fn in have, "%s calls %s which does not exist" % (os.path.basename(f), fn)
        for cmd in set(re.findall(r"cnmfe_b200_mex\('([a-z_]+)'", code)):
            assert cmd in commands, "%s uses gateway command %r which cnmfe_b200_mex.cpp does not implement" % (os.path.basename(f), cmd)
