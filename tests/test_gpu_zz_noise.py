"""GPU parity of Sources2D.estimate_noise (per-pixel GetSn of the raw video).  Kept in its own module, collected after the
core parity tests."""
import numpy as np
import pytest

import gpu_cases as GC

pytestmark = pytest.mark.gpu


def test_estimate_noise_matches_oracle(built_lib):
    """Sources2D.estimate_noise (Sources2D.m:328-379): per-pixel GetSn of the raw video on the default frame range, from a host
    array and from the resident video, with and without the reference's block-border quirk."""
    from oracle import cnmfe as OC, oasis as O
    from cnmf_e_b200.sources2d import Sources2D
    d1, d2, T, K, rr = GC.NOISE_CASE
    D = GC.synthetic("noise")
    obj = Sources2D(d1, d2, T, (d1, d2), ring_radius=rr)
    plain = O.GetSn(D["Y"].reshape(-1, T, order="F").astype(np.float64)).reshape(d1, d2, order="F")
    sn = obj.estimate_noise(D["Y"], chunk=100, replicate_block_quirk=False)
    assert sn.shape == (d1, d2) and np.allclose(sn, plain, rtol=1e-8, atol=1e-12)
    ref = OC.estimate_noise(D["Y"], (d1, d2), rr)
    assert ref.shape == (d1, d2) and not np.array_equal(ref, plain)          # the quirk moves border rows
    assert np.allclose(obj.estimate_noise(D["Y"]), ref, rtol=1e-8, atol=1e-12)
    sn2 = obj.estimate_noise(D["Y"], frame_range=(11, 310), replicate_block_quirk=False)
    ref2 = O.GetSn(D["Y"].reshape(-1, T, order="F")[:, 10:310].astype(np.float64)).reshape(d1, d2, order="F")
    assert np.allclose(sn2, ref2, rtol=1e-8, atol=1e-12)
    # from the resident video (cnmfe_estimate_noise)
    obj.load_video(D["Y"])
    assert np.allclose(obj.estimate_noise(), ref, rtol=1e-8, atol=1e-12)
    assert np.allclose(obj.estimate_noise(frame_range=(11, 310), replicate_block_quirk=False), ref2, rtol=1e-8, atol=1e-12)
    obj.close()
