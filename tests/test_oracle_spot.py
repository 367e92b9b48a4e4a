"""The full-size spot checks (oracle/spot.py) applied to the oracle's OWN small-size results: row-wise re-derivation from the
raw video must agree with the whole-patch oracle update (checks the spot-check machinery bench.py uses at BASELINE sizes)."""
import numpy as np
import scipy.sparse as sp

import gpu_cases as GC


def _slots_from_sparse(view, W):
    Wd = sp.csr_matrix(W)
    nnb = view.r_shift.size
    out = np.zeros((view.nr * view.nc, nnb))
    for m in range(view.nr * view.nc):
        slots, fov, blk = view.ring_of(m)
        out[m, slots] = np.asarray(Wd[m, blk].todense()).ravel()
    return out


def test_spot_checks_agree_with_whole_patch_oracle():
    from oracle import cnmfe as OC, spot
    D = GC.synthetic("ring18")
    d1, d2, T = D["Y"].shape
    rr = 9
    o = OC.OracleSources2D(D["Y"], (d1, d2), ring_radius=rr, options=dict(spatial_algorithm="nnls"))
    o.A, o.C = D["A0"].copy(), D["C0"].copy()
    o.P["sn"] = np.full((d1, d2), 10.0)
    o.update_background_parallel()                 # first run: uniform W
    rs, cs = OC.get_nhood(rr)
    Yf = D["Y"].reshape(-1, T, order="F")
    mp = (0, 0)
    view = spot.PatchView(d1, d2, o.patch_pos[mp], o.block_pos[mp], rs, cs, lambda idx: Yf[idx].astype(np.float64))
    W_old = _slots_from_sparse(view, o.W[mp])
    o.C = o.C * 1.05                               # a changed state, so that the refit moves the active rows
    A_bg, C_bg = o.A.copy(), o.C.copy()
    o.update_background_parallel()                 # steady state: only active pixels refitted
    W_new = _slots_from_sparse(view, o.W[mp])
    rng = np.random.default_rng(0)
    pixels = rng.choice(d1 * d2, 25, replace=False)
    pmax = int((W_old > 0).sum(axis=1).max())
    r = spot.ring_rows(view, pixels, A_bg, C_bg, W_old, W_new, o.b0[mp], pmax)
    assert r["ok"] and r["max_abs_err_W"] <= 1e-9 * r["W_scale"], r
    assert 0 < r["n_refit"] <= 25
    o.update_spatial_parallel(IND=D["IND"])
    maskpix = np.nonzero(np.asarray(sp.csr_matrix(D["IND"]).sum(axis=1)).ravel() > 0)[0]
    pixels = rng.choice(maskpix, 25, replace=False)
    r = spot.spatial_rows(view, pixels, None, None, W_new, o.b0[mp], C_bg, D["IND"], o.A)     # single patch: empty halo
    assert r["ok"] and r["support_equal"] and r["n_nonzero"] > 0, r
    o.update_temporal_parallel()
    # deconv: feed the oracle's own merged C_raw (= C_raw + b is not kept; rebuild it from the definition)
    Craw_in = o.C_raw + 0.0
    o2 = OC.OracleSources2D(D["Y"], (d1, d2), ring_radius=rr)
    o2.C_raw = Craw_in.copy()
    o2.options = o.options
    C2 = o2.deconvTemporal()
    r = spot.deconv_traces(Craw_in, C2, o2.S, o2.C_raw, o2.P["kernel_pars"], o.options["deconv_options"], processes=1)
    assert r["ok"] and r["traces_with_different_spike_support"] == 0, r
