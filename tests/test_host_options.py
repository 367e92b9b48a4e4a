"""CPU-only: host-side option parsing of the deconvolveCa mirror (OASIS_matlab/deconvolveCa.m:208-355: defaults, struct then
name/value merge, rejected values) and error behaviour of compute entry points without a CUDA device."""
import numpy as np
import pytest


def test_make_deconv_opts_defaults_and_merge(built_lib):
    from cnmf_e_b200.oasis import make_deconv_opts
    d, pars, sn = make_deconv_opts()
    assert (d.type, d.method, d.maxIter, d.optimize_b, d.optimize_pars, d.has_tau_range) == (1, 1, 10, 0, 0, 0)   # ar1, constrained
    # the demo's deconv_options (demo_large_data_1p.m:36-42) + the name/value pairs HALS_temporal adds (:92)
    demo = dict(type="ar1", method="foopsi", smin=-5, optimize_pars=True, optimize_b=True, max_tau=100)
    d, pars, sn = make_deconv_opts(demo, maxIter=20, sn=3.5, pars=[0.95])
    assert (d.type, d.method, d.maxIter, d.optimize_b, d.optimize_pars) == (1, 0, 20, 1, 1)
    assert (d.smin, d.max_tau) == (-5.0, 100.0) and sn == 3.5 and pars == [0.95]
    # name/value pairs win over the struct (deconvolveCa.m:233-246)
    d, _, _ = make_deconv_opts(demo, method="thresholded", type="ar2", tau_range=(2, 50))
    assert (d.type, d.method, d.has_tau_range) == (2, 2, 1) and tuple(d.tau_range) == (2.0, 50.0)
    d, _, _ = make_deconv_opts({"lambda": 2.5})
    assert d.lam == 2.5


def test_make_deconv_opts_rejects_unknown(built_lib):
    from cnmf_e_b200.oasis import make_deconv_opts
    for bad in (dict(type="ar3"), dict(method="mcmc"), dict(nonsense=1), dict(remove_large_residuals=True)):
        with pytest.raises(ValueError):
            make_deconv_opts(bad)
    make_deconv_opts(dict(window=200, shift=100, extra_params=None))          # accepted and ignored, as by the reference's methods here


def test_compute_entry_points_fail_loudly_without_gpu(built_lib):
    """No CPU fallback: on a box without a CUDA device a compute call raises with the driver's message."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from cnmf_e_b200 import oasis as G, _lib as L
    with pytest.raises(L.CnmfeError):
        G.GetSn(np.random.default_rng(0).normal(size=(2, 500)))
    with pytest.raises(L.CnmfeError):
        G.deconvolveCa(np.random.default_rng(0).normal(size=500), dict(type="ar1", method="foopsi"))


def test_bench_full_size_invariants_on_oracle_output():
    """bench.py's post-run invariants (A >= 0 inside the mask, oasisAR1 pool algebra of C/S) hold for the oracle's iteration and
    flag a corrupted trace."""
    import bench
    from oracle import gen, cnmfe as OC
    D = gen.make_synthetic(48, 40, 600, 5, seed=3, nblob=2)
    o = OC.OracleSources2D(D["Y"], (48, 40), ring_radius=6, options=dict(spatial_algorithm="nnls"))
    o.A, o.C = D["A0"].copy(), D["C0"].copy()
    o.P["sn"] = np.full((48, 40), 10.0)
    o.update_background_parallel(); o.update_spatial_parallel(IND=D["IND"]); o.update_temporal_parallel()
    kp = np.array([p[0] for p in o.P["kernel_pars"]])
    r = bench.full_size_invariants(o.A, D["IND"], o.C, o.S, kp, o.P["neuron_sn"])
    assert r["ok"] and r["n_spikes"] > 0, r
    C2 = o.C.copy(); C2[0, 10] += 1.0
    assert not bench.full_size_invariants(o.A, D["IND"], C2, o.S, kp, o.P["neuron_sn"])["ok"]


def test_smin_invariant_allows_spike_after_clipped_pool():
    """oasisAR1.m:64-65: the forward test uses the UNCLIPPED value of the preceding pool, so a pool that follows a negative
    pool (clipped to c = 0 at :105) may open with a spike below smin.  The bench invariant must accept exactly that case and
    still reject a sub-smin spike after a positive pool (round-1 VERDICT weak #2)."""
    import bench
    from oracle import oasis as O
    import scipy.sparse as sp
    g, smin = 0.9, 1.0
    # negative pool | one sample at 0.3 (passes the forward test against the UNCLIPPED negative pool) | a large transient
    y = np.concatenate([[-5.0, -5.0, -5.0, 0.3], 6.0 * g ** np.arange(0, 60)])
    c, s, _ = O.oasisAR1(y, g, 0.0, smin)[:3]
    assert 0 < s[3] < smin and c[2] == 0.0          # the oracle (restating the reference) produces the sub-smin spike
    c_py, s_py, _ = O.oasisAR1_py(y, g, 0.0, smin)
    assert np.allclose(c, c_py) and np.allclose(s, s_py)
    A = sp.csc_matrix(np.ones((1, 1))); IND = sp.csc_matrix(np.ones((1, 1), dtype=bool))
    r = bench.full_size_invariants(A, IND, c[None, :], s[None, :], np.array([g]), np.array([smin / 5.0]))
    assert r["ok"] and r["n_spikes_below_smin_after_clipped_pool"] == 1, r
    # the same small spike after a POSITIVE pool violates the rule
    c2 = np.concatenate([1.0 * g ** np.arange(3), [g ** 3 + 0.3], (g ** 3 + 0.3) * g ** np.arange(1, 61)])
    s2 = np.zeros_like(c2); s2[3] = 0.3
    r2 = bench.full_size_invariants(A, IND, c2[None, :], s2[None, :], np.array([g]), np.array([smin / 5.0]))
    assert not r2["ok"] and r2["max_rel_spike_below_smin_after_positive_pool"] > 0, r2
