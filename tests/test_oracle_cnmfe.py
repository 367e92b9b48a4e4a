"""CPU: known-answer tests pinning oracle/cnmfe.py."""
import numpy as np
import scipy.sparse as sp
from scipy.optimize import nnls

from oracle import cnmfe as OC, gen, oasis as O


def _small(seed=3, d=40, T=400, K=4, rr=6):
    D = gen.make_synthetic(d, d, T, K, seed=seed, nblob=3)
    return D


def test_get_nhood_counts():
    for r, n in [(18, 120), (9, 56)]:          # get_nhood(18) = 120, get_nhood(9) = 56
        rs, cs = OC.get_nhood(r)
        assert rs.size == n
        R = np.sqrt(rs ** 2 + cs ** 2)
        assert np.all((R >= r) & (R < r + 1))


def test_patch_geometry_matches_survey_example():
    pp, bp = OC.patch_geometry(1024, 1024, (256, 256), 18)
    assert pp.shape[:2] == (4, 4)
    assert list(pp[0, 0]) == [1, 256, 1, 256] and list(pp[3, 3]) == [769, 1024, 769, 1024]
    assert list(bp[1, 1]) == [257 - 19, 513 + 18, 257 - 19, 513 + 18]


def test_ring_fit_satisfies_normal_equations_and_keeps_pattern():
    D = _small()
    Y = D["Y"].reshape(-1, 400, order="F")
    W0 = OC.ring_W_init([1, 40, 1, 40], [1, 40, 1, 40], 40, 40, 6)
    A, C = D["A0"], D["C0"]
    W, b0 = OC.fit_ring_model(Y, A, C, W0, np.nan, None, np.ones((40, 40), bool), True)
    assert (W != 0).nnz == (W0 != 0).nnz
    Ad = A.toarray()
    Bf = (Y - Y.mean(1, keepdims=True)) - Ad @ (C - C.mean(1, keepdims=True))
    for m in (0, 417, 1599):
        ring = W0[m].indices
        X = np.vstack([Bf[ring], np.ones((1, 400))])
        G = X @ X.T
        w_full = np.linalg.solve(G + 1e-5 * np.trace(G) * np.eye(G.shape[0]), X @ Bf[m])
        assert np.allclose(W[m, ring].toarray().ravel(), w_full[:-1] + 1e-100, rtol=1e-9, atol=1e-12)
    assert np.allclose(b0, Y.mean(1) - Ad @ C.mean(1))


def test_bg_subtract_identity_after_bg_update():
    """Right after a BG update (b0 = Ybar - A_prev*Cbar_prev): mean_t(Ysig) = A_prev*mean(C_prev)."""
    D = _small()
    o = OC.OracleSources2D(D["Y"], (40, 40), ring_radius=6)
    o.A, o.C = D["A0"], D["C0"].copy()
    o.update_background_parallel()
    Ysig = o._ysig((0, 0), "temporal")
    assert np.allclose(Ysig.mean(1), o.A_prev @ o.C_prev.mean(1), atol=1e-8)


def test_hals_spatial_monotone_and_nnls_vs_scipy():
    D = _small()
    T = 400
    Y = D["Y"].reshape(-1, T, order="F").astype(float)
    A0 = D["A0"].toarray()
    C = D["C0"]
    mask = D["IND"].toarray()
    Yc = Y - Y.mean(1, keepdims=True)
    Cc = C - C.mean(1, keepdims=True)
    obj = lambda A: np.sum((Yc - A @ Cc) ** 2)
    A1 = OC.HALS_spatial(Y, A0, C, mask, 1)
    A3 = OC.HALS_spatial(Y, A0, C, mask, 3)
    A0m = A0 * mask
    assert obj(A1) <= obj(A0m) + 1e-6 and obj(A3) <= obj(A1) + 1e-6 and A3.min() >= 0
    An = OC.nnls_spatial(Y, A0, C, mask, 20)
    CC, YC = Cc @ Cc.T, Cc @ Yc.T
    for px in np.nonzero(mask.sum(1) > 0)[0][::37]:
        ind = mask[px].astype(bool)
        L = np.linalg.cholesky(CC[np.ix_(ind, ind)])
        ref, _ = nnls(L.T, np.linalg.solve(L, YC[ind, px]))
        assert np.allclose(An[px, ind], ref, atol=2e-4 * max(1.0, ref.max()))    # tol = the solver's 1e-4 threshold


def test_temporal_merge_weights():
    """C_raw = sum_p aa_p C_raw,p / sum_p aa_p (update_temporal_parallel.m:269-280) on a 2x2-patch layout reproduces the
    single-patch answer when footprints do not straddle patches."""
    rng = np.random.default_rng(0)
    aa = rng.uniform(1, 2, (3, 4))
    Cp = rng.standard_normal((3, 4, 50))
    num = (aa[..., None] * Cp).sum(0)
    den = aa.sum(0)
    out = num / den[:, None]
    assert np.allclose(out, np.einsum("pk,pkt->kt", aa / den, Cp))


def test_oracle_iteration_recovers_ground_truth():
    D = gen.make_synthetic(64, 64, 1000, 6, seed=1)
    o = OC.OracleSources2D(D["Y"], (64, 64), ring_radius=9)
    o.A, o.C = D["A0"], D["C0"].copy()
    o.options["spatial_algorithm"] = "hals_thresh"
    o.P["sn"] = O.GetSn(D["Y"].reshape(-1, 1000, order="F").astype(float)).reshape(64, 64, order="F")
    o.update_background_parallel()
    o.update_spatial_parallel(IND=D["IND"])
    o.update_temporal_parallel()
    cc = [np.corrcoef(D["C_true"][k], o.C[k])[0, 1] for k in range(6)]
    assert np.median(cc) > 0.95
