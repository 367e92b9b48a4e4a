"""
oracle/spot.py -- TEST INFRASTRUCTURE ONLY (parity oracle; the product never imports this).

Spot checks of FULL-SIZE results against the float64 oracle: at BASELINE sizes the oracle cannot run a whole update (a
512 x 512 x 10000 ring fit is ~1e14 flop on the CPU), but every output row depends on few inputs, so individual rows can be
re-derived from the raw video exactly as the reference does:

  ring_rows      W(m, :) of sampled pixels   = fit_ring_model.m:92-108 on the 121 video rows of the pixel's ring
  spatial_rows   A(m, :) of sampled pixels   = update_spatial_parallel.m:157-166 (BG subtraction) + nnls_spatial.m / HALS
  deconv_traces  C, S of ALL traces          = deconvTemporal.m:62-84 on the merged C_raw the CUDA path fed its final pass

Used by bench.py AFTER the timed region (reported in config.full_size_checks) and by tests/.
"""
import numpy as np
import scipy.sparse as sp

from . import cnmfe as OC
from . import oasis as O


class PatchView:
    """Geometry of one patch/block + access to rows of its video.  rows_fn(idx) -> (len(idx), T) array of the block pixels
    idx (0-based, r + c*nrb in block coordinates)."""

    def __init__(self, d1, d2, patch_pos, block_pos, r_shift, c_shift, rows_fn):
        self.d1, self.d2 = int(d1), int(d2)
        self.p = [int(x) for x in patch_pos]      # 1-based inclusive r0 r1 c0 c1
        self.b = [int(x) for x in block_pos]
        self.nr, self.nc = self.p[1] - self.p[0] + 1, self.p[3] - self.p[2] + 1
        self.nrb, self.ncb = self.b[1] - self.b[0] + 1, self.b[3] - self.b[2] + 1
        self.r_shift, self.c_shift = np.asarray(r_shift), np.asarray(c_shift)
        self.rows_fn = rows_fn

    def patch_pixel_rc(self, m):
        """FOV (r, c), 0-based, of patch pixel m (MATLAB order inside the patch)."""
        return m % self.nr + self.p[0] - 1, m // self.nr + self.p[2] - 1

    def ring_of(self, m):
        """(slots, fov_linear, block_idx) of the in-FOV ring neighbours of patch pixel m."""
        r, c = self.patch_pixel_rc(m)
        r2, c2 = r + self.r_shift, c + self.c_shift
        ok = (r2 >= 0) & (r2 < self.d1) & (c2 >= 0) & (c2 < self.d2)
        slots = np.nonzero(ok)[0]
        fov = r2[ok] + c2[ok] * self.d1
        blk = (c2[ok] - (self.b[2] - 1)) * self.nrb + (r2[ok] - (self.b[0] - 1))
        return slots, fov, blk

    def centre_of(self, m):
        r, c = self.patch_pixel_rc(m)
        return r + c * self.d1, (c - (self.b[2] - 1)) * self.nrb + (r - (self.b[0] - 1))


def _sub_problem(view, m, A_csr, C, W_row_slots):
    """The rows of the block a single patch pixel's ring regression / BG subtraction touches, ordered by block index (the order
    logical indexing gives in fit_ring_model.m:97-99)."""
    slots, fov, blk = view.ring_of(m)
    fov_c, blk_c = view.centre_of(m)
    all_blk = np.concatenate([blk, [blk_c]])
    all_fov = np.concatenate([fov, [fov_c]])
    order = np.argsort(all_blk, kind="stable")
    pos_of = np.empty_like(order)
    pos_of[order] = np.arange(order.size)
    Y = view.rows_fn(all_blk[order].astype(np.int32))
    A_d = C_d = None
    if A_csr is not None and A_csr.shape[1] > 0:
        Asub = A_csr[all_fov[order], :]
        ksel = np.nonzero(np.asarray(abs(Asub).sum(axis=0)).ravel() > 0)[0]
        A_d = np.asarray(Asub[:, ksel].todense()) if ksel.size else None
        C_d = np.asarray(C)[ksel] if ksel.size else None
    n = order.size
    W = sp.csr_matrix((np.asarray(W_row_slots)[slots], (np.zeros(slots.size, dtype=int), pos_of[:slots.size])), shape=(1, n))
    ind_patch = np.zeros(n, dtype=bool)
    ind_patch[pos_of[-1]] = True
    return Y, A_d, C_d, W, ind_patch, slots, pos_of


def ring_rows(view, pixels, A, C, W_old_slots, W_new_slots, b0_new, pmax, bg_acceleration=True, thresh_outlier=np.nan, sn=None):
    """Re-fit the ring weights of the patch pixels `pixels` with oracle.fit_ring_model (fit_ring_model.m:1-128) from the raw
    video rows and compare with the CUDA result.  A, C: the neurons the BG update saw (d x K csc, K x T).  W_*_slots:
    (d_patch, nnb) slot-form weights before / after the update.  Returns max abs errors relative to max|W|."""
    A_csr = sp.csr_matrix(A)
    err_w = err_b0 = 0.0
    n_active = 0
    scale = max(1.0, float(np.abs(W_new_slots[pixels]).max()))
    for m in pixels:
        Y, A_d, C_d, W_old, ind_patch, slots, pos_of = _sub_problem(view, m, A_csr, C, W_old_slots[m])
        sn_m = None if sn is None else np.array([sn[m]])
        W_fit, b0 = OC.fit_ring_model(Y, A_d, C_d, W_old, thresh_outlier, sn_m, ind_patch, bg_acceleration, pmax=pmax)
        W_fit = np.asarray(W_fit.todense()).ravel()
        n_active += int(np.any(W_fit != np.asarray(W_old.todense()).ravel()))
        got = np.asarray(W_new_slots[m])[slots]
        err_w = max(err_w, float(np.abs(got - W_fit[pos_of[:slots.size]]).max()) if slots.size else 0.0)
        err_b0 = max(err_b0, abs(float(b0[0]) - float(b0_new[m])))
    return dict(n_pixels=int(len(pixels)), n_refit=n_active, max_abs_err_W=err_w, W_scale=scale,
                max_abs_err_b0=err_b0, ok=bool(err_w <= 1e-7 * scale and err_b0 <= 1e-7 * max(1.0, float(np.abs(b0_new).max()))))


def spatial_rows(view, pixels, A_prev, C_prev, W_slots, b0, C, IND, A_new, method="nnls", sn=None, maxN=20):
    """Rows of the spatial update for the patch pixels `pixels`: BG-subtract the pixel's video row as
    update_spatial_parallel.m:157-166 does (Ysig = Y - W (Y - A_prev C_prev) - (b0 - W mean(...))), then solve the row with the
    oracle's nnls_spatial (nnls_spatial.m:14-109) / HALS variants over the neurons whose search mask covers the pixel.
    A_prev / C_prev: the neurons the reference SELECTS for the subtraction -- update_spatial_parallel.m:82-98 keeps only the
    neurons touching the block's HALO (the patch was overwritten with 2 before `mask(:)==1`), so with a single patch
    (block == patch) nothing is subtracted: pass None.  A_new: d x K csc result of the CUDA path."""
    Ap_csr = None if A_prev is None else sp.csr_matrix(A_prev)
    IND_csr = sp.csr_matrix(IND).astype(bool)
    A_new_csr = sp.csr_matrix(A_new)
    C = np.asarray(C, dtype=np.float64)
    K = C.shape[0]
    Ysig = np.zeros((len(pixels), C.shape[1]))
    mask = np.zeros((len(pixels), K), dtype=bool)
    got = np.zeros((len(pixels), K))
    for i, m in enumerate(pixels):
        Y, A_d, C_d, W, ind_patch, slots, pos_of = _sub_problem(view, m, Ap_csr, C_prev, W_slots[m])
        Ysig[i] = OC.bg_subtract_ring(Y, A_d, C_d, W, np.array([b0[m]]), ind_patch)[0]
        fov_c, _ = view.centre_of(m)
        mask[i, IND_csr[fov_c].indices] = True
        got[i] = np.asarray(A_new_csr[fov_c].todense()).ravel()
    if method == "nnls":
        ref = OC.nnls_spatial(Ysig, None, C, mask, maxN)
    elif method == "hals":
        raise ValueError("HALS couples the rows through A only via V: use nnls for row-wise spot checks")
    else:
        raise ValueError(method)
    scale = max(1.0, float(np.abs(ref).max()))
    err = float(np.abs(got - ref).max()) if len(pixels) else 0.0
    return dict(n_pixels=int(len(pixels)), n_nonzero=int((ref != 0).sum()), max_abs_err_A=err, A_scale=scale,
                support_equal=bool(np.array_equal(got != 0, ref != 0)), ok=bool(err <= 1e-7 * scale))


def _deconv_one(args):
    ck_raw, opts = args
    if np.any(np.isnan(ck_raw)):
        T = ck_raw.size
        return np.zeros(T), np.zeros(T), 0.0, np.zeros(1), 0.0
    sn = O.GetSn(ck_raw)
    ck, sk, topt = O.deconvolveCa(ck_raw, opts, sn=sn)
    if np.sum(np.abs(ck)) == 0:
        ck = ck_raw
    return ck, sk, float(topt["b"]), np.atleast_1d(topt["pars"]).ravel(), float(sn)


def deconv_traces(Craw_in, C, S, C_raw_out, kernel_pars, deconv_options, processes=None):
    """deconvTemporal.m:62-84 on every trace of the merged C_raw the CUDA path deconvolved, compared with its C, S, C_raw - b."""
    import multiprocessing as mp
    Craw_in = np.asarray(Craw_in)
    K = Craw_in.shape[0]
    jobs = [(Craw_in[k], deconv_options) for k in range(K)]
    if processes is None:
        processes = min(16, mp.cpu_count())
    if processes > 1 and K > 8:
        with mp.get_context("fork").Pool(processes) as pool:
            res = pool.map(_deconv_one, jobs, chunksize=max(1, K // (4 * processes)))
    else:
        res = [_deconv_one(j) for j in jobs]
    C, S, C_raw_out = np.asarray(C), np.asarray(S), np.asarray(C_raw_out)
    bad_support, err_c, err_s, err_raw, err_g = 0, 0.0, 0.0, 0.0, 0.0
    n_spikes = n_sub_smin = 0
    for k, (ck, sk, b, pars, sn) in enumerate(res):
        scale = max(1.0, float(np.abs(ck).max()))
        bad_support += int(not np.array_equal(S[k] > 0, sk > 0))
        err_c = max(err_c, float(np.abs(C[k] - ck).max()) / scale)
        err_s = max(err_s, float(np.abs(S[k] - sk).max()) / scale)
        err_raw = max(err_raw, float(np.abs(C_raw_out[k] - (Craw_in[k] - b)).max()) / scale)
        if kernel_pars is not None:
            err_g = max(err_g, abs(float(np.asarray(kernel_pars[k]).ravel()[0]) - float(pars[0])))
        n_spikes += int((sk > 0).sum())
        smin = abs(float(deconv_options.get("smin", 0.0))) * sn if float(deconv_options.get("smin", 0.0)) < 0 else float(deconv_options.get("smin", 0.0))
        n_sub_smin += int(((sk > 0) & (sk < smin * (1 - 1e-12))).sum())
    return dict(n_traces=K, traces_with_different_spike_support=bad_support, max_rel_err_C=err_c, max_rel_err_S=err_s,
                max_rel_err_C_raw=err_raw, max_abs_err_g=err_g, oracle_n_spikes=n_spikes,
                oracle_spikes_below_smin=n_sub_smin,
                ok=bool(bad_support == 0 and err_c <= 1e-7 and err_s <= 1e-7 and err_raw <= 1e-7))
