"""GPU parity of Sources2D.estimate_noise (per-pixel GetSn of the raw video).  Kept in its own module, collected after the
core parity tests."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_estimate_noise_matches_oracle(built_lib):
    """Sources2D.estimate_noise (Sources2D.m:328-379): per-pixel GetSn of the raw video on the default frame range."""
    from oracle import gen, oasis as O
    from cnmf_e_b200.sources2d import Sources2D
    d1, d2, T = 20, 16, 400
    D = gen.make_synthetic(d1, d2, T, 3, seed=8, nblob=2)
    obj = Sources2D(d1, d2, T, (d1, d2), ring_radius=6)
    sn = obj.estimate_noise(D["Y"], chunk=100)
    ref = O.GetSn(D["Y"].reshape(-1, T, order="F").astype(np.float64)).reshape(d1, d2, order="F")
    assert sn.shape == (d1, d2) and np.allclose(sn, ref, rtol=1e-8, atol=1e-12)
    sn2 = obj.estimate_noise(D["Y"], frame_range=(11, 310))
    ref2 = O.GetSn(D["Y"].reshape(-1, T, order="F")[:, 10:310].astype(np.float64)).reshape(d1, d2, order="F")
    assert np.allclose(sn2, ref2, rtol=1e-8, atol=1e-12)
