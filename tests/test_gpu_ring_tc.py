"""The tcgen05 INT8 second-moment kernel must equal the exact SIMT kernel bit for bit (integers), and both must equal a
direct numpy evaluation on a small block."""
import ctypes
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _moments(d1, d2, T, rr, seed, hi=65535):
    from cnmf_e_b200.sources2d import Sources2D
    from cnmf_e_b200 import _lib as L
    rng = np.random.default_rng(seed)
    Y = rng.integers(0, hi + 1, size=(d1, d2, T), dtype=np.uint16)
    g = Sources2D(d1, d2, T, (d1, d2), ring_radius=rr)
    g.load_video(Y)
    ND = 2 * rr * (4 * rr + 1) + 2 * rr + 1
    out = {}
    for use_tensor in (0, 1):
        S2 = np.zeros((d1 * d2, ND))
        L.check(L.lib().cnmfe_debug_second_moments(g._h, 0, use_tensor, S2.ctypes.data_as(ctypes.c_void_p)))
        out[use_tensor] = S2
    g.close()
    return Y, out, ND


def _disp_id(dr, dc, rr):
    return dr if dc == 0 else (2 * rr + 1) + (dc - 1) * (4 * rr + 1) + (dr + 2 * rr)


@pytest.mark.parametrize("shape", [(64, 48, 700, 9), (100, 37, 1300, 18)])
def test_tensor_equals_simt_and_numpy(built_lib, shape):
    d1, d2, T, rr = shape
    Y, out, ND = _moments(d1, d2, T, rr, seed=5)
    assert np.array_equal(out[0], out[1]), "tensor-core moments differ from the SIMT kernel: max |diff| = %g" % np.abs(out[0] - out[1]).max()
    Yf = Y.reshape(-1, T, order="F").astype(np.int64)
    rng = np.random.default_rng(1)
    for _ in range(200):
        r, c = rng.integers(0, d1), rng.integers(0, d2)
        dc = rng.integers(0, 2 * rr + 1)
        dr = rng.integers(0 if dc == 0 else -2 * rr, 2 * rr + 1)
        r2, c2 = r + dr, c + dc
        if not (0 <= r2 < d1 and 0 <= c2 < d2):
            continue
        ref = int(np.dot(Yf[r + c * d1], Yf[r2 + c2 * d1]))
        assert out[1][r + c * d1, _disp_id(dr, dc, rr)] == float(ref)


def test_tensor_long_video_two_passes(built_lib):
    """T > 16384 frames: the kernel runs two K passes and accumulates (int32 accumulators must not overflow)."""
    Y, out, ND = _moments(40, 8, 17000, 2, seed=9)
    assert np.array_equal(out[0], out[1])


@pytest.mark.parametrize("shape", [(64, 48, 700, 9), (100, 37, 1300, 18)])
def test_cta_pair_kernel_equals_simt(built_lib, shape, monkeypatch):
    """The cta_group::2 variant (two SMs share one M = 256 MMA, each staging half of the neighbour operand; opt-in with
    CNMFE_TC_MODE=pair, read per call) is bit-exact too -- incl. an odd number of columns, where the pair's second tile is
    partly outside the block."""
    monkeypatch.setenv("CNMFE_TC_MODE", "pair")
    d1, d2, T, rr = shape
    _, out, _ = _moments(d1, d2, T, rr, seed=6)
    assert np.array_equal(out[0], out[1]), "CTA-pair moments differ from the SIMT kernel: max |diff| = %g" % np.abs(out[0] - out[1]).max()


def test_cta_pair_kernel_two_passes(built_lib, monkeypatch):
    monkeypatch.setenv("CNMFE_TC_MODE", "pair")
    _, out, _ = _moments(40, 8, 17000, 2, seed=10)
    assert np.array_equal(out[0], out[1])
