"""GPU parity of the ring model with spatial down-sampling (bg_ssub = 2, the setting of demo_large_data_1p.m:30)
against oracle/ssub.py (imresize restated).  fp64 everywhere -> tight tolerances; spike support identical."""
import numpy as np
import pytest
import scipy.sparse as sp

import gpu_cases as GC

pytestmark = pytest.mark.gpu


def _close(a, b, tol=1e-7):
    scale = max(1.0, float(np.abs(b).max()))
    err = float(np.abs(np.asarray(a) - np.asarray(b)).max())
    assert err <= tol * scale, "max abs err %g (scale %g)" % (err, scale)


@pytest.mark.parametrize("shape", [("ssub_60x52", (60, 52), 2), ("ssub_75x64_patches", (38, 32), 2), ("ssub3_48x45", (48, 45), 3)])
def test_bg_ssub_chain(built_lib, shape):
    from oracle import oasis as O
    from oracle.ssub import OracleSources2DSsub
    from cnmf_e_b200.sources2d import Sources2D
    case, patch, ssub = shape
    D = GC.synthetic(case)
    d1, d2, T = D["Y"].shape
    sn = O.GetSn(D["Y"].reshape(-1, T, order="F").astype(np.float64)).reshape(d1, d2, order="F")
    orc = OracleSources2DSsub(D["Y"], patch, ring_radius=10, bg_ssub=ssub, options=dict(spatial_algorithm="hals_thresh"))
    gpu = Sources2D(d1, d2, T, patch, ring_radius=10, options=dict(bg_ssub=ssub, spatial_algorithm="hals_thresh"))
    gpu.load_video(D["Y"])
    for o in (orc, gpu):
        o.A, o.C = D["A0"].copy(), D["C0"].copy()
        o.P["sn"] = sn
    for it in range(2):      # first run (uniform W) and a steady-state run
        orc.update_background_parallel()
        gpu.update_background_parallel()
        for i, mp in enumerate(orc.patches()):
            _close(gpu.ring_as_sparse_ssub(i).toarray(), sp.csr_matrix(orc.W[mp]).toarray())
            _close(gpu.b0[i], orc.b0[mp])
        orc.update_spatial_parallel(IND=D["IND"])
        gpu.update_spatial_parallel(IND=D["IND"])
        _close(gpu.A.toarray(), orc.A.toarray())
        orc.update_temporal_parallel()
        gpu.update_temporal_parallel()
        _close(gpu.C_raw, orc.C_raw)
        _close(gpu.C, orc.C)
        assert np.array_equal(gpu.S > 0, orc.S > 0)
        # keep both sides on exactly the same state for the second round
        gpu.A, gpu.C = orc.A.copy(), orc.C.copy()
    gpu.close()
