"""CPU: pin the oracle's OASIS stack with solver-independent known-answer tests; the reference
has no golden vectors of its own."""
import numpy as np
import pytest
from scipy.optimize import nnls
from scipy import signal

from oracle import oasis as O


def _trace(T=600, g=0.95, noise=0.3, seed=13, n=0):
    Y, truth, sp = O.gen_data(gam=g, noise=noise, T=T, N=n + 1, seed=seed)
    return Y[n], truth[n], sp[n]


def test_c_core_matches_pure_python():
    y, _, _ = _trace()
    for lam, smin in [(0.0, 0.0), (0.5, 0.0), (0.0, 0.4), (0.3, 0.2)]:
        c, s, P = O.oasisAR1(y, 0.95, lam, smin)
        c2, s2, pools = O.oasisAR1_py(y, 0.95, lam, smin)
        assert len(P) == len(pools)
        assert np.allclose(c, c2, atol=1e-12) and np.allclose(s, s2, atol=1e-12)


def test_ar1_foopsi_solves_the_convex_problem():
    """oasisAR1(y,g,lam,0) = argmin 1/2|c-y|^2 + lam|s|_1, s=Gc>=0 (the check the reference draws by eye against CVX,
    examples/ar1_foopsi.m:19-20): compare with NNLS on the impulse-response dictionary."""
    y, _, _ = _trace(T=300)
    g, lam, T = 0.95, 0.4, 300
    K = np.tril(g ** (np.arange(T)[:, None] - np.arange(T)[None, :]).clip(min=0))
    # with c = K s (s_1 = c_1): 1/2|c-y|^2 + lam*[(1-g) sum_{t<T} c_t + c_T] = 1/2|Ks-y|^2 + lam*sum(s)  (oasisAR1.m:46-49)
    L = np.linalg.cholesky(K.T @ K)
    rhs = np.linalg.solve(L, K.T @ y - lam * np.ones(T))
    s_ref, _ = nnls(L.T, rhs, maxiter=50 * T)
    c, s, _ = O.oasisAR1(y, g, lam, 0.0)
    s_full = s.copy()
    s_full[0] = c[0]
    obj = lambda sv: 0.5 * np.sum((K @ sv - y) ** 2) + lam * np.sum(sv)
    assert abs(obj(s_full) - obj(s_ref)) < 1e-8 * max(1.0, obj(s_ref))
    assert np.allclose(K @ s_full, c, atol=1e-9)
    assert np.allclose(s_full, s_ref, atol=1e-6)


def test_pav_invariants_with_smin():
    y, _, _ = _trace(T=2000, noise=0.2)
    c, s, P = O.oasisAR1(y, 0.95, 0.0, 0.5)
    assert P.t[0] == 0 and np.all(P.t[1:] == P.t[:-1] + P.l[:-1]) and P.t[-1] + P.l[-1] == 2000
    nz = s > 0
    assert np.all(s[nz] >= 0.5 - 1e-12)
    r = c[1:] - 0.95 * c[:-1]
    assert np.allclose(r[~nz[1:]], 0, atol=1e-12)


def test_constrained_hits_the_noise_level():
    y, _, _ = _trace(T=3000)
    c, s, o = O.deconvolveCa(y)                       # method 'constrained' (deconvolveCa.m:220)
    assert abs(np.sum((y - c) ** 2) - o["sn"] ** 2 * y.size) < 1e-3
    c, s, o = O.deconvolveCa(y, optimize_b=True, optimize_pars=True)
    assert abs(np.sum((y - c - o["b"]) ** 2) - o["sn"] ** 2 * y.size) < 1e-3


def test_pwelch_restatement_matches_scipy():
    y, _, _ = _trace(T=3000)
    L = int(3000 / 4.5)
    f, P = signal.welch(y, fs=1, window=signal.get_window("hamming", L, fftbins=False), noverlap=L // 2, nfft=1024,
                        detrend=False, scaling="density")
    Pm, ff = O.pwelch_psd(y)
    assert np.allclose(ff, f) and np.allclose(Pm, P, rtol=1e-12)
    sn = O.GetSn(y)
    assert 0.25 < sn < 0.35                              # true noise 0.3


def test_fminbnd_matches_scipy_fmm():
    from scipy.optimize import fminbound
    f = lambda x: (x - 0.613) ** 2 + 0.1 * np.sin(7 * x)
    x, _ = O.fminbnd(f, 0.0, 1.0)
    xs = fminbound(f, 0.0, 1.0, xtol=1e-4)
    assert abs(x - xs) < 2e-4


def test_time_constant_recovers_g():
    Y, _, _ = O.gen_data(gam=0.95, noise=0.1, T=20000, N=1, seed=2, firerate=2.0)
    g = O.estimate_time_constant(Y[0], 1, O.GetSn(Y[0]))
    assert abs(g[0] - 0.95) < 0.02


def test_ar2_recursion_and_threshold():
    Y, truth, sp = O.gen_data([1.7, -0.712], 0.5, 3000, 30, 0.5, 0, 1, 3)
    c, s, o = O.deconvolveCa(Y[0], type="ar2", pars=[1.7, -0.712], method="foopsi", smin=-3)
    nz = s > 0
    assert np.all(s[nz] >= o["smin"] - 1e-12)
    assert np.corrcoef(c, truth[0])[0, 1] > 0.9


def test_ar2_thresholded_optimize_b_is_one_pass_on_the_baseline_corrected_trace():
    """thresholded_oasisAR2.m:127-163 with optimize_g = false: b = estimate_baseline_noise(y), then ONE oasisAR2 pass on y - b
    (the loop exits at its first RSS test).  A constant added to the trace moves b by (nearly) that constant -- the histogram
    bins of estimate_baseline_noise are not exactly shift invariant in floating point."""
    Y, _, _ = O.gen_data([1.7, -0.712], 0.5, 3000, 30, 0.5, 0, 1, 7)
    y = Y[0]
    opts = dict(type="ar2", pars=[1.7, -0.712], method="thresholded", optimize_b=True)
    c0, s0, o0 = O.deconvolveCa(y, opts)
    c1, s1, o1 = O.deconvolveCa(y + 12.5, opts)
    assert abs((o1["b"] - o0["b"]) - 12.5) < 0.05
    assert np.mean((s0 > 0) != (s1 > 0)) < 0.01
    b_ref, _ = O.estimate_baseline_noise(y)
    assert abs(o0["b"] - b_ref) < 1e-12
    smin = O.choose_smin([1.7, -0.712], o0["sn"], 0.99999999)
    c_ref, s_ref, _ = O.oasisAR2(y - b_ref, [1.7, -0.712], None, smin)
    assert np.array_equal(c0, c_ref) and np.array_equal(s0, s_ref)


def test_gen_data_spikes_follow_matlab_rand_stream():
    """MT19937 + column-major fill: first uniform of RandomState(13) is MATLAB's rand after rng(13)."""
    rs = np.random.RandomState(13)
    assert abs(rs.rand() - 0.7777024105738202) < 1e-15


def test_fit_gauss1_and_baseline_noise_known_answers():
    """fit_gauss1 recovers an exact Gaussian; hist_centers equals an explicit-edge histogram; estimate_baseline_noise recovers
    the baseline / noise of a Gaussian sample with sparse positive transients (functions/estimate_baseline_noise.m)."""
    from oracle import oasis as O
    x = np.linspace(-3, 5, 161)
    mu, sig, A = O.fit_gauss1(x, 7.0 * np.exp(-(x - 1.2) ** 2 / 2 / 0.8 ** 2), 0.3, 3)
    assert abs(mu - 1.2) < 1e-9 and abs(sig - 0.8) < 1e-9 and abs(A - 7.0) < 1e-8
    rs = np.random.RandomState(1)
    y = rs.randn(4000)
    c = np.linspace(-2, 2, 9)
    edges = np.concatenate([[min(c[0] - 0.25, y.min())], (c[:-1] + c[1:]) / 2, [max(c[-1], y.max())]])
    ref = np.histogram(y, edges)[0]
    assert np.array_equal(O.hist_centers(y, c), ref)
    y = 2.5 + 0.3 * rs.randn(20000)
    y[::40] += 4.0
    b, sn = O.estimate_baseline_noise(y)
    assert abs(b - 2.5) < 0.02 and abs(sn - 0.3) < 0.02
    assert np.allclose(O._solve_small([[2.0, 1, 0], [1, 3, 1], [0, 1, 4]], [1.0, 2, 3]), np.linalg.solve([[2.0, 1, 0], [1, 3, 1], [0, 1, 4]], [1.0, 2, 3]))
