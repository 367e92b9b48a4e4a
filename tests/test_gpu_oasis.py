"""GPU parity: per-trace kernels (deconvolveCa variants, GetSn, HALS_temporal) vs the float64 oracle.
Tolerances: all kernels are fp64 -> rtol 1e-7 on continuous outputs (differences come only from summation order),
spike support (indices where s>0) must be IDENTICAL."""
import numpy as np
import pytest

import gpu_cases as GC

pytestmark = pytest.mark.gpu


def _traces(name):
    return GC.traces(name)


def _compare(Y, opts, built_lib, rtol=1e-7):
    from cnmf_e_b200 import oasis as G
    from oracle import oasis as O
    r = G.deconvolveCa_batch(Y, opts)
    for n in range(Y.shape[0]):
        c, s, o = O.deconvolveCa(Y[n], opts)
        scale = max(1.0, np.abs(c).max())
        assert np.allclose(r["c"][n], c, rtol=rtol, atol=rtol * scale), (n, np.abs(r["c"][n] - c).max())
        assert np.array_equal(np.nonzero(r["s"][n] > 0)[0], np.nonzero(s > 0)[0]), "spike support differs, trace %d" % n
        assert np.allclose(r["s"][n], s, rtol=rtol, atol=rtol * scale)
        assert np.allclose(r["b"][n], o["b"], rtol=rtol, atol=1e-9)
        assert np.allclose(r["pars"][n][:np.size(o["pars"])], np.atleast_1d(o["pars"]), rtol=rtol)
        assert np.allclose(r["sn"][n], o["sn"], rtol=1e-9)


def test_getsn_matches_oracle(built_lib):
    from cnmf_e_b200 import oasis as G
    from oracle import oasis as O
    for T in (1000, 3000, 10000, 20011):
        Y, _, _ = _traces("ar1_getsn_%d" % T)
        sn = G.GetSn(Y)
        ref = O.GetSn(Y)
        assert np.allclose(sn, ref, rtol=1e-10), (T, sn, ref)


def test_foopsi_ar1_demo_options(built_lib):
    """deconv_options of demos/demo_large_data_1p.m:41-47"""
    Y, _, _ = _traces("ar1_n01")
    _compare(Y, dict(type="ar1", method="foopsi", smin=-5, optimize_pars=True, optimize_b=True, max_tau=100), built_lib)


def test_foopsi_ar1_fixed_g_lambda(built_lib):
    Y, _, _ = _traces("ar1_n03")
    _compare(Y, dict(type="ar1", method="foopsi", pars=[0.95], **{"lambda": 0.5}), built_lib)
    _compare(Y, dict(type="ar1", method="foopsi", pars=[0.95], smin=0.4, optimize_pars=True), built_lib)


def test_constrained_ar1(built_lib):
    Y, _, _ = _traces("ar1_n03")
    _compare(Y, dict(), built_lib)                                             # deconvolveCa.m:220 default method
    _compare(Y, dict(optimize_b=True, optimize_pars=True), built_lib, rtol=1e-6)


def test_thresholded_ar1(built_lib):
    Y, _, _ = _traces("ar1_n03")
    _compare(Y, dict(method="thresholded", optimize_pars=True), built_lib, rtol=1e-6)


def test_thresholded_ar1_optimize_b(built_lib):
    """thresholded_oasisAR1 with optimize_b (estimate_baseline_noise + fit_gauss1, thresholded_oasisAR1.m:141-181), with and
    without optimize_pars, on traces with a non-zero baseline."""
    Y, _, _ = _traces("ar1_baseline")
    _compare(Y, dict(method="thresholded", optimize_b=True), built_lib, rtol=1e-6)
    _compare(Y, dict(method="thresholded", optimize_b=True, optimize_pars=True), built_lib, rtol=1e-6)


def test_foopsi_ar2(built_lib):
    Y, _, _ = _traces("ar2_4")
    _compare(Y, dict(type="ar2", method="foopsi", pars=[1.7, -0.712], smin=-3), built_lib)
    _compare(Y, dict(type="ar2", method="foopsi", smin=-3), built_lib, rtol=1e-6)


def test_thresholded_ar2(built_lib):
    Y, _, _ = _traces("ar2_3")
    _compare(Y, dict(type="ar2", method="thresholded", pars=[1.7, -0.712]), built_lib)
    # optimize_b (thresholded_oasisAR2.m:127-163): baseline from estimate_baseline_noise, one pass on y - b
    _compare(Y + 7.5, dict(type="ar2", method="thresholded", pars=[1.7, -0.712], optimize_b=True), built_lib)


def test_ar2_at_c5_length(built_lib):
    """BASELINE configs[4] trace length (T = 100000, AR2 traces of functions/gen_data.m with g = [1.7, -0.712], noise 1, seed 3, restated): parity with the
    oracle, not only invariants -- deconvolveCa(y, 'ar2', 'foopsi', pars, 'smin', -3) and the 'thresholded' variant."""
    Y, _, _ = _traces("ar2_c5")
    _compare(Y, dict(type="ar2", method="foopsi", pars=[1.7, -0.712], smin=-3), built_lib)
    _compare(Y[:3], dict(type="ar2", method="thresholded", pars=[1.7, -0.712]), built_lib)


def test_pav_invariants_large(built_lib):
    """Size-independent properties at a BASELINE-scale T: s>=smin or 0, c_t = g c_{t-1} off spikes."""
    from cnmf_e_b200 import oasis as G
    Y, _, _ = _traces("ar1_long")
    r = G.deconvolveCa_batch(Y, dict(type="ar1", method="foopsi", pars=[0.95], smin=0.5))
    for n in range(Y.shape[0]):
        c, s = r["c"][n], r["s"][n]
        nz = s > 0
        assert np.all(s[nz] >= 0.5 - 1e-9)
        resid = c[1:] - 0.95 * c[:-1]
        assert np.allclose(resid[~nz[1:]], 0, atol=1e-9)
        assert np.allclose(resid[nz[1:]], s[1:][nz[1:]], atol=1e-9)


def test_hals_temporal_uv(built_lib):
    from cnmf_e_b200 import oasis as G
    from oracle import cnmfe as OC
    D = GC.synthetic("hals_uv")
    Y = D["Y"].reshape(-1, 1500, order="F").astype(np.float64)
    Y = Y - Y.mean(axis=1, keepdims=True)
    A = D["A0"].toarray()
    C0 = D["C0"]
    opts = dict(type="ar1", method="foopsi", smin=-5, optimize_pars=True, optimize_b=True, max_tau=100)
    Cr, Craw_r, res_r, S_r = OC.HALS_temporal(Y, A, C0, 3, opts)
    C, Craw, res, S = G.HALS_temporal_uv(A.T @ Y, A.T @ A, C0, 3, opts)
    assert np.allclose(C, Cr, rtol=1e-7, atol=1e-7)
    assert np.allclose(Craw, Craw_r, rtol=1e-7, atol=1e-7)
    assert np.array_equal(S > 0, S_r > 0)
    assert np.allclose(res["sn"], res_r["sn"], rtol=1e-9)
    # no-deconvolution branch (HALS_temporal.m:64-68)
    Cr2, Craw_r2, _, _ = OC.HALS_temporal(Y, A, C0, 2, None)
    C2, Craw2, _, _ = G.HALS_temporal_uv(A.T @ Y, A.T @ A, C0, 2, None)
    assert np.allclose(C2, Cr2, rtol=1e-9, atol=1e-9)

