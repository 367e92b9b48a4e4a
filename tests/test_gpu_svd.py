"""GPU parity of the svd background model (2p path, demo_large_data_2p.m:46) against oracle/svd_bg.py.
b and f are defined up to a common sign per component (eigs), so parity is stated on the product b*f and on b0."""
import numpy as np
import pytest
import scipy.sparse as sp

import gpu_cases as GC

pytestmark = pytest.mark.gpu


def _close(a, b, tol=1e-7):
    scale = max(1.0, float(np.abs(b).max()))
    err = float(np.abs(np.asarray(a) - np.asarray(b)).max())
    assert err <= tol * scale, "max abs err %g (scale %g)" % (err, scale)


@pytest.mark.parametrize("shape", [("svd_48x40", (48, 40), 1), ("svd_64x60_patches", (32, 30), 2)])
def test_svd_background_chain(built_lib, shape):
    from oracle import oasis as O
    from oracle.svd_bg import OracleSources2DSVD
    from cnmf_e_b200.sources2d import Sources2D
    case, patch, nb = shape
    D = GC.synthetic(case)
    d1, d2, T = D["Y"].shape
    sn = O.GetSn(D["Y"].reshape(-1, T, order="F").astype(np.float64)).reshape(d1, d2, order="F")
    orc = OracleSources2DSVD(D["Y"], patch, ring_radius=6, nb=nb, options=dict(spatial_algorithm="hals"))
    gpu = Sources2D(d1, d2, T, patch, ring_radius=6, options=dict(background_model="svd", nb=nb, spatial_algorithm="hals"))
    gpu.load_video(D["Y"])
    for o in (orc, gpu):
        o.A, o.C = D["A0"].copy(), D["C0"].copy()
        o.P["sn"] = sn
    orc.update_background_parallel()
    gpu.update_background_parallel()
    for i, mp in enumerate(orc.patches()):
        _close(gpu.b[i] @ gpu.f[i], orc.b[mp] @ orc.f[mp], 1e-7)
        _close(gpu.b0[i], orc.b0[mp], 1e-7)
    orc.update_spatial_parallel(IND=D["IND"])
    gpu.update_spatial_parallel(IND=D["IND"])
    _close(gpu.A.toarray(), orc.A.toarray(), 1e-6)
    orc.update_temporal_parallel()
    gpu.update_temporal_parallel()
    _close(gpu.C_raw, orc.C_raw, 1e-6)
    _close(gpu.C, orc.C, 1e-6)
    assert np.array_equal(gpu.S > 0, orc.S > 0)
    gpu.close()


def test_nmf_background_subtraction(built_lib):
    """nmf model with b, f supplied by the caller (cnmfe_set_bf): the BG subtraction Y - b*f of the spatial and
    temporal updates (update_spatial_parallel.m:179-182, update_temporal_parallel.m:165-168) must match the oracle."""
    from oracle.svd_bg import OracleSources2DSVD
    from cnmf_e_b200.sources2d import Sources2D
    d1, d2, T, K = 40, 36, 400, 4
    D = GC.synthetic("nmf_sub")
    Yf = D["Y"].reshape(-1, T, order="F").astype(np.float64)
    u, s, vt = np.linalg.svd(Yf, full_matrices=False)
    b = np.abs(u[:, :1] * s[0]); f = np.abs(vt[:1])                 # a deterministic non-negative rank-1 factorisation

    class NmfOracle(OracleSources2DSVD):
        def _ysig(self, mp, phase):
            return self._get_block(self.patch_pos[mp]).astype(np.float64) - self.b[mp] @ self.f[mp]

    orc = NmfOracle(D["Y"], (d1, d2), ring_radius=6, nb=1, options=dict(spatial_algorithm="hals"))
    gpu = Sources2D(d1, d2, T, (d1, d2), ring_radius=6, options=dict(background_model="nmf", nb=1, spatial_algorithm="hals"))
    gpu.load_video(D["Y"])
    for o in (orc, gpu):
        o.A, o.C = D["A0"].copy(), D["C0"].copy()
    orc.b[(0, 0)], orc.f[(0, 0)] = b, f
    gpu.b[0], gpu.f[0], gpu.b0[0] = b, f, np.zeros(d1 * d2)
    orc.update_spatial_parallel(IND=D["IND"]); gpu.update_spatial_parallel(IND=D["IND"])
    _close(gpu.A.toarray(), orc.A.toarray(), 1e-6)
    orc.update_temporal_parallel(); gpu.update_temporal_parallel()
    _close(gpu.C_raw, orc.C_raw, 1e-6)
    assert np.array_equal(gpu.S > 0, orc.S > 0)
    gpu.close()


@pytest.mark.parametrize("shape", [("nmf_fit", (48, 40), 1), ("svd_64x60_patches", (32, 30), 2)])
def test_nmf_background_fit_chain(built_lib, shape):
    """update_background_parallel with background_model = 'nmf' (fit_nmf_model.m -> nnmf ALS from the shared fixed start) and
    the spatial / temporal updates on Y - b*f, against oracle/nmf_bg.py.  Tolerance 1e-6: the ALS runs to the same iteration
    count on both sides and is a contraction, but sums over T and d are ordered differently."""
    from oracle.nmf_bg import OracleSources2DNMF
    from cnmf_e_b200.sources2d import Sources2D
    case, patch, nb = shape
    D = GC.synthetic(case)
    d1, d2, T = D["Y"].shape
    orc = OracleSources2DNMF(D["Y"], patch, ring_radius=6, nb=nb, options=dict(spatial_algorithm="hals"))
    gpu = Sources2D(d1, d2, T, patch, ring_radius=6, options=dict(background_model="nmf", nb=nb, spatial_algorithm="hals"))
    gpu.load_video(D["Y"])
    for o in (orc, gpu):
        o.A, o.C = D["A0"].copy(), D["C0"].copy()
    orc.update_background_parallel()
    gpu.update_background_parallel()
    for i, mp in enumerate(orc.patches()):
        assert np.allclose(np.sum(gpu.f[i] ** 2, axis=1), 1.0, rtol=1e-12)          # rows of f have unit length
        _close(gpu.b[i] @ gpu.f[i], orc.b[mp] @ orc.f[mp], 1e-6)
        _close(gpu.f[i], orc.f[mp], 1e-6)
    if len(orc.patches()) == 1:      # the reference's nmf BG subtraction only has consistent shapes for block == patch
        orc.update_spatial_parallel(IND=D["IND"]); gpu.update_spatial_parallel(IND=D["IND"])
        _close(gpu.A.toarray(), orc.A.toarray(), 1e-6)
        orc.update_temporal_parallel(); gpu.update_temporal_parallel()
        _close(gpu.C_raw, orc.C_raw, 1e-6)
        assert np.array_equal(gpu.S > 0, orc.S > 0)
    gpu.close()
