"""Golden vectors on REAL reference data (a crop of demos/data_1p.tif, see tests/golden/make_golden.py).
CPU: the oracle reproduces the committed vectors.  GPU: the CUDA path (through the C ABI) matches them."""
import os
import numpy as np
import pytest
import scipy.sparse as sp

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "data_1p_crop.npz")


def _load():
    g = np.load(GOLD)
    W = sp.csr_matrix((g["W_data"], g["W_indices"], g["W_indptr"]), shape=(1600, 1600))
    return g, W


def _close(a, b, tol):
    scale = max(1.0, float(np.abs(b).max()))
    assert float(np.abs(np.asarray(a) - np.asarray(b)).max()) <= tol * scale


def test_oracle_reproduces_golden():
    from oracle import cnmfe as OC
    g, W = _load()
    o = OC.OracleSources2D(g["Y"], (40, 40), ring_radius=9, options=dict(spatial_algorithm="hals_thresh"))
    o.A, o.C = sp.csc_matrix(g["A0"]), g["C0"].copy()
    o.P["sn"] = g["sn"]
    o.update_background_parallel()
    _close(sp.csr_matrix(o.W[(0, 0)]).toarray(), W.toarray(), 1e-9)
    o.update_spatial_parallel(IND=sp.csc_matrix(g["IND"]))
    _close(o.A.toarray(), g["A1"], 1e-9)
    o.update_temporal_parallel()
    _close(o.C, g["C"], 1e-8)
    assert np.array_equal(o.S > 0, g["S"] > 0)


@pytest.mark.gpu
def test_cuda_path_matches_golden(built_lib):
    from cnmf_e_b200.sources2d import Sources2D
    g, W = _load()
    n = Sources2D(40, 40, 600, (40, 40), ring_radius=9, options=dict(spatial_algorithm="hals_thresh"))
    n.load_video(g["Y"])
    n.A, n.C = sp.csc_matrix(g["A0"]), g["C0"].copy()
    n.P["sn"] = g["sn"]
    n.update_background_parallel(True)
    _close(n.ring_as_sparse(0).toarray(), W.toarray(), 1e-7)
    _close(n.b0[0], g["b0"], 1e-9)
    n.update_spatial_parallel(True, IND=sp.csc_matrix(g["IND"]))
    _close(n.A.toarray(), g["A1"], 1e-7)
    n.update_temporal_parallel(True)
    _close(n.C_raw, g["C_raw"], 1e-7)
    _close(n.C, g["C"], 1e-7)
    assert np.array_equal(n.S > 0, g["S"] > 0), "spike support differs from the golden vectors"
    _close(n.P["kernel_pars"][:, 0], g["kernel_pars"], 1e-7)
    n.close()
