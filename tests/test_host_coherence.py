"""CPU-only: the host mirror's coherence tracking (cnmf_e_b200/sources2d.py: state is re-sent to the device only when the host
object changed) exercised against a recording stand-in for the library -- no device, no arithmetic."""
from unittest import mock
import numpy as np
import scipy.sparse as sp


class _RecordingLib:
    """Every entry point returns 0 and is counted."""
    def __init__(self):
        self.calls = {}

    def __getattr__(self, name):
        if not name.startswith("cnmfe_"):
            raise AttributeError(name)

        def f(*a):
            self.calls[name] = self.calls.get(name, 0) + 1
            return 0
        return f


def _make(lib, d1=12, d2=10, T=50):
    from cnmf_e_b200 import sources2d, _lib
    with mock.patch.object(_lib, "lib", return_value=lib):
        return sources2d.Sources2D(d1, d2, T, (d1, d2), ring_radius=3)


def test_state_is_resent_only_when_it_changed():
    lib = _RecordingLib()
    obj = _make(lib)
    d, T = obj.d1 * obj.d2, obj.T
    rng = np.random.default_rng(0)
    obj.A = sp.random(d, 3, density=0.1, format="csc", random_state=1)
    obj.C = rng.normal(size=(3, T))
    obj.push_neurons()
    assert lib.calls.get("cnmfe_set_neurons") == 1
    n0 = obj.h2d_bytes
    obj.push_neurons()                                  # writable host arrays cannot be trusted to be unchanged: sent again
    assert lib.calls["cnmfe_set_neurons"] == 2 and obj.h2d_bytes > n0
    # frozen (read-only) objects the device is known to hold are not re-sent
    obj.A, obj.C = obj._freeze(obj.A), obj._freeze(obj.C)
    obj.push_neurons()
    assert lib.calls["cnmfe_set_neurons"] == 3
    n1 = obj.h2d_bytes
    obj.push_neurons()
    assert lib.calls["cnmfe_set_neurons"] == 3 and obj.h2d_bytes == n1
    # a new C object (A unchanged): one more call
    obj.C = obj._freeze(rng.normal(size=(3, T)))
    obj.push_neurons()
    assert lib.calls["cnmfe_set_neurons"] == 4
    # shape mismatch is caught on the host
    obj.C = np.zeros((2, T))
    try:
        obj.push_neurons()
        raise SystemExit("expected an assertion")
    except AssertionError:
        pass


def test_trace_range_and_owner_partition_are_consistent():
    from cnmf_e_b200.sources2d import patch_owners, trace_range
    for npatch, ws in [(16, 8), (4, 2)]:
        o = patch_owners(npatch, ws)
        assert sorted(set(o.tolist())) == list(range(ws))
    assert [trace_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]


def test_ring_weights_resent_unless_read_only():
    lib = _RecordingLib()
    obj = _make(lib)
    dp = obj.d1 * obj.d2
    W, b0 = np.zeros((dp, max(obj.nnb, 1))), np.zeros(dp)
    obj.W[0], obj.b0[0] = W, b0
    obj.push_ring(); obj.push_ring()
    assert lib.calls["cnmfe_set_ring"] == 2             # writable arrays: never assumed unchanged
    W.setflags(write=False); b0.setflags(write=False)
    obj.push_ring(); obj.push_ring()
    assert lib.calls["cnmfe_set_ring"] == 3             # read-only and identical: sent once more, then trusted
    obj.b0[0] = np.zeros(dp)
    obj.push_ring()
    assert lib.calls["cnmfe_set_ring"] == 4
