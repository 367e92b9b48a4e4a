"""GPU parity of compute_RSS / reconstruct_background (Sources2D.m:1247-1510) with the oracle: after a full
iteration (so that A, C differ from A_prev, C_prev and b0_new differs from b0), single patch and 2 x 2 patches."""
import numpy as np
import pytest

import gpu_cases as GC

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [("chain_64", (64, 64)), ("chain_96x80_patches", (48, 40))])
def test_compute_rss_and_reconstruct_background(built_lib, shape):
    from oracle import cnmfe as OC, oasis as O
    from cnmf_e_b200.sources2d import Sources2D
    case, patch = shape
    D = GC.synthetic(case)
    d1, d2, T = D["Y"].shape
    sn = O.GetSn(D["Y"].reshape(-1, T, order="F").astype(np.float64)).reshape(d1, d2, order="F")
    orc = OC.OracleSources2D(D["Y"], patch, ring_radius=9, options=dict(spatial_algorithm="hals_thresh"))
    gpu = Sources2D(d1, d2, T, patch, ring_radius=9, options=dict(spatial_algorithm="hals_thresh"))
    gpu.load_video(D["Y"])
    for o in (orc, gpu):
        o.A, o.C = D["A0"].copy(), D["C0"].copy()
        o.P["sn"] = sn
        o.update_background_parallel()
        o.update_spatial_parallel(IND=D["IND"])
        o.update_temporal_parallel()
    tot_o, rss_o = orc.compute_RSS()
    tot_g, rss_g = gpu.compute_RSS()
    ref = np.array([rss_o[mp] for mp in orc.patches()])
    assert np.allclose(rss_g, ref, rtol=1e-9), (rss_g, ref)
    assert abs(tot_g - tot_o) <= 1e-9 * tot_o and gpu.P["RSS"] == tot_g
    tot_o2, _ = orc.compute_RSS((11, 300))
    tot_g2, _ = gpu.compute_RSS((11, 300))
    assert abs(tot_g2 - tot_o2) <= 1e-9 * tot_o2
    Yo = orc.reconstruct_background((5, 44))
    Yg = gpu.reconstruct_background((5, 44))
    assert Yg.shape == Yo.shape == (d1, d2, 40)
    assert np.abs(Yg - Yo).max() <= 1e-8 * np.abs(Yo).max()
    gpu.close()
