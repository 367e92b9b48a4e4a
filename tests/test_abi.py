"""CPU-only: the C-ABI library builds for sm_100a, loads, and exports every symbol include/cnmfe_b200.h declares."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_all_header_symbols_exported(built_lib):
    from cnmf_e_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "cnmfe_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(cnmfe_[a-z0-9_]+)\s*\(", hdr))
    assert names, "no declarations parsed"
    for n in sorted(names):
        assert hasattr(built_lib, n), "symbol %s declared in the header but not exported" % n
        assert n in _lib.SYMBOLS, "symbol %s has no ctypes prototype" % n


def test_defaults_match_reference_parser(built_lib):
    """deconvolveCa.m:212-230 / CNMFSetParms.m defaults."""
    import ctypes
    from cnmf_e_b200 import _lib
    d = _lib.DeconvOpts()
    built_lib.cnmfe_deconv_defaults(ctypes.byref(d))
    assert (d.type, d.method, d.maxIter, d.optimize_b, d.optimize_pars) == (1, 1, 10, 0, 0)
    assert (d.smin, d.lam, d.b, d.max_tau, d.thresh_factor, d.p_noise) == (0.0, 0.0, 0.0, 100.0, 1.0, 0.9999)
    o = _lib.Options()
    built_lib.cnmfe_options_defaults(ctypes.byref(o))
    assert o.maxIter_temporal == 5 and o.deconv_flag == 1 and o.spatial_algorithm == 0


def test_mex_gateway_type_checks_against_the_header():
    """matlab/cnmfe_b200_mex.cpp cannot be built here (no MATLAB), but it must stay in sync with the C ABI: compile it
    (-fsyntax-only) against a stub mex.h and the real include/cnmfe_b200.h."""
    import subprocess
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-I" + os.path.join(ROOT, "tests", "stubs"),
                        "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "matlab", "cnmfe_b200_mex.cpp")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
