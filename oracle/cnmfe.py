"""
oracle/cnmfe.py -- TEST INFRASTRUCTURE ONLY (parity oracle; the product never imports this).

Float64 NumPy/SciPy restatement of the CNMF-E alternating-update hot path, following (file:line under
/root/reference/ca_source_extraction unless noted):

  get_nhood               endoscope/get_nhood.m:1-24
  ring_W_init             @Sources2D/initComponents_parallel.m:213-236 (== update_background_parallel.m:70-118)
  patch_geometry          endoscope/distribute_data.m:55-77,163-173
  fit_ring_model          endoscope/fit_ring_model.m:1-128
  bg_subtract_ring        @Sources2D/update_spatial_parallel.m:157-166 (== update_temporal_parallel.m:144-153)
  HALS_spatial            utilities/HALS_spatial.m:17-44
  HALS_spatial_thresh     utilities/HALS_spatial_thresh.m:17-53
  nnls_spatial (+nnls)    endoscope/nnls_spatial.m:14-109
  lars_spatial (+nnls)    utilities/lars_spatial.m:15-151  (including the thresh(m) indexing bug at :55)
  HALS_temporal           utilities/HALS_temporal.m:20-119
  OracleSources2D.update_background_parallel   @Sources2D/update_background_parallel.m:1-334 (ring, bg_ssub=1)
  OracleSources2D.update_spatial_parallel      @Sources2D/update_spatial_parallel.m:1-366
  OracleSources2D.update_temporal_parallel     @Sources2D/update_temporal_parallel.m:1-313
  OracleSources2D.deconvTemporal               @Sources2D/deconvTemporal.m:1-106

All pixel indexing is MATLAB column-major (r fastest) but 0-based.  Video `Y` is (d1, d2, T).
PARITY UNPINNED against real MATLAB (no golden vectors in the reference); pinned by the
known-answer tests in tests/test_oracle_cnmfe.py.  Out-of-scope hooks:
determine_search_location / post_process_spatial are injected by the caller (defaults: given mask / identity).
"""
import numpy as np
import scipy.sparse as sp

from . import oasis as O


# ----------------------------------------------------------------------------- geometry
def get_nhood(radius, k=None):
    """get_nhood.m:7-24.  Returns (r_shift, c_shift) in MATLAB find() order (column-major)."""
    sub = np.arange(-radius, radius + 1)
    cind, rind = np.meshgrid(sub, sub)
    R = np.sqrt(cind ** 2 + rind ** 2)
    kern = (R >= radius) & (R < radius + 1)
    cc, rr = np.nonzero(kern.T)           # column-major order: iterate columns then rows
    r_shift = rr - radius
    c_shift = cc - radius
    if k is None or k > r_shift.size:
        return r_shift, c_shift
    temp = np.arctan2(r_shift, c_shift)
    ids = np.argsort(temp, kind="stable")
    ind = np.round(np.linspace(1, ids.size, k)).astype(int) - 1
    return r_shift[ids[ind]], c_shift[ids[ind]]


def patch_geometry(d1, d2, patch_dims, w_overlap):
    """distribute_data.m:39,55-77,163-173.  Returns patch_pos, block_pos as (nr,nc,4) int arrays, 1-based inclusive
    [r0,r1,c0,c1] exactly as the reference stores them."""
    min_w = 2 * w_overlap + 3

    def idx(dn, pd, force_last):
        npatch = int(np.round(dn / pd))          # MATLAB round: half away from zero; inputs are positive
        if dn / pd - np.floor(dn / pd) == 0.5:
            npatch = int(np.floor(dn / pd) + 1)
        if npatch <= 1:
            return np.array([1, dn])
        pi = np.ceil(np.linspace(1, dn, npatch + 1)).astype(int)
        if force_last:
            pi[-1] = dn
        if pi[1] - pi[0] < min_w:
            pi = np.arange(1, dn + 1, min_w)
            pi[-1] = dn
        return pi

    pr = idx(d1, patch_dims[0], True)
    pc = idx(d2, patch_dims[1], False)
    nr_p, nc_p = len(pr) - 1, len(pc) - 1
    patch_pos = np.zeros((nr_p, nc_p, 4), dtype=int)
    block_pos = np.zeros((nr_p, nc_p, 4), dtype=int)
    for m in range(nr_p):
        for n in range(nc_p):
            patch_pos[m, n] = [pr[m], pr[m + 1] - (m != nr_p - 1), pc[n], pc[n + 1] - (n != nc_p - 1)]
            block_pos[m, n] = [max(1, pr[m] - w_overlap - 1), min(d1, pr[m + 1] + w_overlap),
                               max(1, pc[n] - w_overlap - 1), min(d2, pc[n + 1] + w_overlap)]
    return patch_pos, block_pos


def ind_patch_mask(tmp_patch, tmp_block):
    """Logical (nr_block, nc_block) mask of the patch inside its block (update_*_parallel.m)."""
    nrb = tmp_block[1] - tmp_block[0] + 1
    ncb = tmp_block[3] - tmp_block[2] + 1
    m = np.zeros((nrb, ncb), dtype=bool)
    m[tmp_patch[0] - tmp_block[0]:tmp_patch[1] - tmp_block[0] + 1,
      tmp_patch[2] - tmp_block[2]:tmp_patch[3] - tmp_block[2] + 1] = True
    return m


def ring_W_init(tmp_patch, tmp_block, d1, d2, ring_radius, num_neighbors=None):
    """initComponents_parallel.m:213-236 (bg_ssub==1): uniform row-normalised ring weights, CSR d_p x d_blk."""
    r_shift, c_shift = get_nhood(int(np.ceil(ring_radius)), num_neighbors)
    nr = tmp_patch[1] - tmp_patch[0] + 1
    nc = tmp_patch[3] - tmp_patch[2] + 1
    nrb = tmp_block[1] - tmp_block[0] + 1
    ncb = tmp_block[3] - tmp_block[2] + 1
    csub, rsub = np.meshgrid(np.arange(tmp_patch[2], tmp_patch[3] + 1), np.arange(tmp_patch[0], tmp_patch[1] + 1))
    csub = csub.T.reshape(-1, 1)   # column-major flatten
    rsub = rsub.T.reshape(-1, 1)
    ii = np.repeat(np.arange(csub.size).reshape(-1, 1), r_shift.size, axis=1)
    cs = csub + c_shift.reshape(1, -1)
    rs = rsub + r_shift.reshape(1, -1)
    ind = (cs >= 1) & (cs <= d2) & (rs >= 1) & (rs <= d1)
    jj = (cs - tmp_block[2]) * nrb + (rs - tmp_block[0] + 1) - 1
    temp = sp.csr_matrix((np.ones(ind.sum()), (ii[ind], jj[ind])), shape=(nr * nc, nrb * ncb))
    rowsum = np.asarray(temp.sum(axis=1)).ravel()
    return sp.diags(1.0 / rowsum) @ temp


# ----------------------------------------------------------------------------- background
def is_first_run(W_old):
    """length(unique(W_old(1,:)))==2 (fit_ring_model.m:25, update_background_parallel.m:143)."""
    return len(np.unique(np.asarray(W_old[0, :].todense()).ravel())) == 2


def fit_ring_model(Y, A, C, W_old, thresh_outlier, sn=None, ind_patch=None, with_projection=True, pmax=None):
    """fit_ring_model.m:11-128.  Y: (d_blk,T) any dtype; A: (d_blk,K); C: (K,T); W_old: sparse (d_p,d_blk).
    Returns W (CSR, same pattern as W_old) and b0 (d_p,).
    pmax (not in the reference): overrides max_m nnz(W_old(m,:) > 0) of :61 -- used by spot checks that hand over a few rows
    of a large patch and must subsample frames exactly as the full call did."""
    Y = np.asarray(Y)
    d, T = Y.shape
    if A is None or np.size(A) == 0:
        A = np.ones((d, 1))
        C = np.zeros((1, T))
    A = np.asarray(A.todense()) if sp.issparse(A) else np.asarray(A, dtype=np.float64)
    C = np.asarray(C, dtype=np.float64)
    W_old = sp.csr_matrix(W_old)
    if is_first_run(W_old):
        ind_active = np.ones(W_old.shape[0], dtype=bool)
    else:
        ind_active = np.asarray(abs(W_old) @ A.sum(axis=1)).ravel() > 0
    if ind_patch is None:
        ind_patch = np.ones(d, dtype=bool)
    ind_patch = np.asarray(ind_patch).ravel(order="F").astype(bool)
    Ymean = Y.mean(axis=1, dtype=np.float64)
    Cmean = C.mean(axis=1)
    b0 = Ymean[ind_patch] - A[ind_patch, :] @ Cmean
    Yc = Y.astype(np.float64) - Ymean[:, None]
    Cc = C - Cmean[:, None]
    Bf = Yc - A @ Cc
    use_outlier = not (thresh_outlier is None or np.isnan(thresh_outlier))
    if use_outlier:
        Bf_old = W_old @ Bf
        tmp_Bf = Bf[ind_patch, :]
        ind_outlier = tmp_Bf > (Bf_old + thresh_outlier * np.asarray(sn).reshape(-1, 1))
        tmp_Bf[ind_outlier] = Bf_old[ind_outlier]
        Bf[ind_patch, :] = tmp_Bf
    if pmax is None:
        pmax = int(np.max(np.asarray((W_old > 0).sum(axis=1)).ravel()))
    nmax = pmax * 100
    if use_outlier and nmax < T:
        temp = ind_outlier.sum(axis=0)
        ind_frames = temp <= O.quantile(temp, nmax / T)
        nmax = int(np.count_nonzero(ind_frames))
        Bf = Bf[:, ind_frames]
    ind_pixels = np.nonzero(ind_patch)[0]
    dpx = ind_pixels.size
    T = Bf.shape[1]
    if with_projection:
        nk = min(int(np.round(T / 1)), nmax)
        k = int(np.floor(T / nk))
        if k != 1:
            Bf = Bf[:, ::k]
    vec_ones = np.ones((1, Bf.shape[1]))
    indptr, indices = W_old.indptr, W_old.indices
    data = W_old.data
    Wd = W_old.copy()
    for m in range(dpx):
        if not ind_active[m]:
            continue
        idx = ind_pixels[m]
        sl = slice(indptr[m], indptr[m + 1])
        ring = indices[sl][data[sl] != 0]
        order = np.argsort(ring, kind="stable")        # logical row indexing selects in ascending column order
        ring_sorted = ring[order]
        y = Bf[idx, :]
        X = np.vstack([Bf[ring_sorted, :], vec_ones])
        XX = X @ X.T
        Xy = X @ y
        w = np.linalg.solve(XX + np.eye(XX.shape[0]) * np.trace(XX) * 1e-5, Xy)
        vals = w[:-1] + 1e-100
        pos = np.nonzero(data[sl] != 0)[0][order]
        Wd.data[indptr[m] + pos] = vals
    return Wd, b0


def bg_subtract_ring(Yblk, A_prev, C_prev, W, b0, ind_patch):
    """update_spatial_parallel.m:157-166 (bg_ssub==1):
    Ysig = Y(patch,:) - W*(Y - A_prev*C_prev) - (b0 - W*mean(Y - A_prev*C_prev, 2))."""
    ind_patch = np.asarray(ind_patch).ravel(order="F").astype(bool)
    Yd = np.asarray(Yblk, dtype=np.float64)
    if A_prev is not None and np.size(C_prev) > 0:
        Ap = np.asarray(A_prev.todense()) if sp.issparse(A_prev) else np.asarray(A_prev)
        tmp_Y = Yd - Ap @ np.asarray(C_prev, dtype=np.float64)
    else:
        tmp_Y = Yd
    W = sp.csr_matrix(W)
    return (Yd[ind_patch, :] - W @ tmp_Y) - (np.asarray(b0).reshape(-1, 1) - (W @ tmp_Y.mean(axis=1)).reshape(-1, 1))


# ----------------------------------------------------------------------------- spatial solvers
def _moments(Y, C):
    T = C.shape[1]
    Cmean = C.mean(axis=1)
    Ymean = Y.mean(axis=1)
    U = Y @ C.T - T * np.outer(Ymean, Cmean)
    V = C @ C.T - T * np.outer(Cmean, Cmean)
    return U, V


def HALS_spatial(Y, A, C, active_pixel=None, maxIter=1):
    """HALS_spatial.m:17-44."""
    Y = np.asarray(Y, dtype=np.float64)
    A = np.array(A.todense() if sp.issparse(A) else A, dtype=np.float64)
    C = np.asarray(C, dtype=np.float64)
    mask = np.ones(A.shape, dtype=bool) if active_pixel is None or np.size(active_pixel) == 0 else \
        np.asarray(active_pixel.todense() if sp.issparse(active_pixel) else active_pixel).astype(bool)
    A[~mask] = 0
    K = A.shape[1]
    U, V = _moments(Y, C)
    cc = np.diag(V)
    for _ in range(maxIter):
        for k in range(K):
            if cc[k] == 0:
                continue
            ind = mask[:, k]
            ak = np.maximum(0, A[ind, k] + (U[ind, k] - A[ind, :] @ V[:, k]) / cc[k])
            A[ind, k] = ak
    return A


def HALS_spatial_thresh(Y, A, C, active_pixel=None, maxIter=1, sn=None):
    """HALS_spatial_thresh.m:17-53."""
    Y = np.asarray(Y, dtype=np.float64)
    A = np.array(A.todense() if sp.issparse(A) else A, dtype=np.float64)
    C = np.asarray(C, dtype=np.float64)
    mask = np.ones(A.shape, dtype=bool) if active_pixel is None or np.size(active_pixel) == 0 else \
        np.asarray(active_pixel.todense() if sp.issparse(active_pixel) else active_pixel).astype(bool)
    if sn is None:
        sn = O.GetSn(Y)
    sn = np.asarray(sn, dtype=np.float64).ravel(order="F")
    A[~mask] = 0
    K = A.shape[1]
    U, V = _moments(Y, C)
    cc = np.diag(V)
    with np.errstate(divide="ignore"):
        cc_thr = 3.0 / np.sqrt(cc)
    for _ in range(maxIter):
        for k in range(K):
            if cc[k] == 0:
                continue
            ind = mask[:, k]
            if ind.sum() == 0:
                A[:, k] = 0
                continue
            ak = A[ind, k] + (U[ind, k] - A[ind, :] @ V[:, k]) / cc[k]
            ak[ak < sn[ind] * cc_thr[k]] = 0
            A[ind, k] = ak
    return A


def nnls_active_set(A, b, s=None, tol=1e-9, maxIter=None, thresh=None):
    """nnls sub-function: nnls_spatial.m:41-109 (thresh=None) / lars_spatial.m:62-151 (thresh given)."""
    A = np.asarray(A, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64).ravel()
    p = A.shape[1]
    s = np.zeros(p) if s is None else np.array(s, dtype=np.float64)
    if maxIter is None:
        maxIter = p
    if (s > 0).sum() > maxIter:
        s = np.zeros(p)
    for _ in range(int(maxIter)):
        l = b - A @ s
        Pset = s > 0
        if np.max(l) < tol:
            break
        if thresh is not None:
            if s @ A @ s - 2 * s @ b <= thresh:
                break
        temp = int(np.argmax(l))
        Pset[temp] = True
        if Pset.sum() > maxIter:
            break
        mu = None
        while np.any(Pset):
            try:
                mu = np.linalg.solve(A[np.ix_(Pset, Pset)], b[Pset])
            except np.linalg.LinAlgError:
                mu = np.linalg.solve(A[np.ix_(Pset, Pset)] + tol * np.eye(Pset.sum()), b[Pset])
            if np.all(mu > tol):
                break
            sp_ = s[Pset]
            temp2 = sp_ / (sp_ - mu)
            temp2 = temp2[~(mu > tol)]
            a = np.min(temp2)
            s[Pset] = sp_ + a * (mu - sp_)
            Pset[s < tol] = False
        s[Pset] = mu
    return s


def nnls_spatial(Y, A, C, active_pixel=None, maxN=5):
    """nnls_spatial.m:14-38."""
    Y = np.asarray(Y, dtype=np.float64)
    C = np.asarray(C, dtype=np.float64)
    d, K = Y.shape[0], C.shape[0]
    mask = np.ones((d, K), dtype=bool) if active_pixel is None or np.size(active_pixel) == 0 else \
        np.asarray(active_pixel.todense() if sp.issparse(active_pixel) else active_pixel).astype(bool)
    Yc = Y - Y.mean(axis=1, keepdims=True)
    Cc = C - C.mean(axis=1, keepdims=True)
    CC = Cc @ Cc.T
    YC = Cc @ Yc.T
    ind_fit = np.nonzero(mask.sum(axis=1) > 1e-9)[0]
    Aout = np.zeros((d, K))
    for m in ind_fit:
        ind = mask[m, :]
        Aout[m, ind] = nnls_active_set(CC[np.ix_(ind, ind)], YC[ind, m], None, 1e-4, maxN)
    return Aout


def lars_spatial(Y, A, C, active_pixel=None, sn=None):
    """lars_spatial.m:15-58; reproduces the reference's `thresh(m)` (loop counter, not pixel index) at :55."""
    Y = np.asarray(Y, dtype=np.float64)
    C = np.asarray(C, dtype=np.float64)
    d, T = Y.shape
    K = C.shape[0]
    if K == 0:
        return np.zeros((d, 0))
    if sn is None:
        sn = O.GetSn(Y)
    sn = np.asarray(sn, dtype=np.float64).ravel(order="F")
    mask = np.ones((d, K), dtype=bool) if active_pixel is None or np.size(active_pixel) == 0 else \
        np.asarray(active_pixel.todense() if sp.issparse(active_pixel) else active_pixel).astype(bool)
    Yc = Y - Y.mean(axis=1, keepdims=True)
    Cc = C - C.mean(axis=1, keepdims=True)
    CC = Cc @ Cc.T
    YC = Cc @ Yc.T
    ind_fit = np.nonzero(mask.sum(axis=1) > 1e-9)[0]
    Aout = np.zeros((d, K))
    thresh = sn ** 2 * T - np.sum(Yc ** 2, axis=1)
    for m, px in enumerate(ind_fit):
        ind = mask[px, :]
        Aout[px, ind] = nnls_active_set(CC[np.ix_(ind, ind)], YC[ind, px], None, 1e-9, None, thresh[m])
    return Aout


# ----------------------------------------------------------------------------- temporal
def HALS_temporal(Y, A, C, maxIter=1, deconv_options=None):
    """HALS_temporal.m:20-119.  Returns C, C_raw, results_deconv(dict or None), S."""
    Y = np.asarray(Y, dtype=np.float64)
    A = np.asarray(A.todense() if sp.issparse(A) else A, dtype=np.float64)
    C = np.array(C, dtype=np.float64)
    K = A.shape[1]
    T = Y.shape[1]
    deconv_flag = bool(deconv_options)
    C_raw = np.zeros((K, T))
    U = A.T @ Y
    V = A.T @ A
    aa = np.diag(V)
    ind_update = np.nonzero(aa > 0)[0]
    S = np.zeros((K, T))
    sn = np.zeros(K)
    kernel_pars = [None] * K
    for miter in range(1, maxIter + 1):
        for k in ind_update:
            ck_raw = C[k, :] + (U[k, :] - V[k, :] @ C) / aa[k]
            if not deconv_flag:
                ck_raw = ck_raw - np.min(ck_raw)
                C[k, :] = ck_raw
                C_raw[k, :] = ck_raw
            else:
                b = np.mean(ck_raw[ck_raw < np.median(ck_raw)])
                sn_psd = O.GetSn(ck_raw)
                ck_raw = ck_raw - b
                sn[k] = sn_psd
                ck, sk, topt = O.deconvolveCa(ck_raw, deconv_options, maxIter=20, sn=sn_psd, pars=kernel_pars[k])
                kernel_pars[k] = np.atleast_1d(topt["pars"]).ravel()
                ck_raw = ck_raw - topt["b"]
                if np.sum(np.abs(ck)) == 0:
                    ck = ck_raw
                C[k, :] = ck
                if miter == maxIter:
                    S[k, :] = sk
                    C_raw[k, :] = ck_raw
    res = dict(sn=sn, kernel_pars=kernel_pars) if deconv_flag else None
    return C, C_raw, res, S


def estimate_noise(Y, patch_dims, w_overlap, frame_range=None):
    """Sources2D.estimate_noise (@Sources2D/Sources2D.m:328-379), method 'psd': GetSn per pixel, evaluated block by block on
    the block grid of distribute_data.m:91-98 (block_idx = sorted unique clamped patch borders -1-w / +w), INCLUDING the
    reference's assembly quirk: row / column `end-1` of every non-last block is deleted (:369-374) although consecutive blocks
    share their border row, so the kept border row appears twice and the row before it never does.  Y: (d1, d2, T)."""
    d1, d2, T = Y.shape
    if frame_range is None:
        frame_range = (1, min(T, 3000))
    patch_pos, _ = patch_geometry(d1, d2, patch_dims, w_overlap)
    pr = sorted(set(int(x) for x in patch_pos[:, 0, 0]) | {int(patch_pos[-1, 0, 1])})
    pc = sorted(set(int(x) for x in patch_pos[0, :, 2]) | {int(patch_pos[0, -1, 3])})
    # patch_idx_r as distribute_data.m:56-67 built it: starts of the patches and the last row
    def block_idx(pidx, dn):
        b = []
        for x in pidx:
            b += [x - 1 - w_overlap, x + w_overlap]
        b = [min(max(v, 1), dn) for v in b]
        return sorted(set(b))
    bir, bic = block_idx(pr, d1), block_idx(pc, d2)
    nrb, ncb = len(bir) - 1, len(bic) - 1
    cols = []
    for n in range(ncb):
        rows = []
        for m in range(nrb):
            r0, r1, c0, c1 = bir[m], bir[m + 1], bic[n], bic[n + 1]
            Yp = Y[r0 - 1:r1, c0 - 1:c1, frame_range[0] - 1:frame_range[1]].astype(np.float64)
            tmp = O.GetSn(Yp.reshape(-1, Yp.shape[2], order="F")).reshape(r1 - r0 + 1, c1 - c0 + 1, order="F")
            if m != nrb - 1:
                tmp = np.delete(tmp, -2, axis=0)
            if n != ncb - 1:
                tmp = np.delete(tmp, -2, axis=1)
            rows.append(tmp)
        cols.append(np.vstack(rows))
    return np.hstack(cols)


# ----------------------------------------------------------------------------- Sources2D restatement
class OracleSources2D:
    """The slice of Sources2D state the hot path touches (Sources2D.m:10-57) + the three _parallel updates.
    `Y` (d1,d2,T) plays the role of the RAM-mapped mat_data (get_patch_data.m)."""

    def __init__(self, Y, patch_dims, ring_radius=18, options=None):
        self.Y = Y
        self.d1, self.d2, self.T = Y.shape
        self.options = dict(ring_radius=ring_radius, bg_ssub=1, background_model="ring", bg_acceleration=True,
                            num_neighbors=None, thresh_outlier=np.nan, spatial_algorithm="hals", maxIter=5,
                            deconv_flag=True, deconv_options=dict(type="ar1", method="foopsi", smin=-5,
                                                                  optimize_pars=True, optimize_b=True, max_tau=100))
        if options:
            self.options.update(options)
        self.patch_pos, self.block_pos = patch_geometry(self.d1, self.d2, patch_dims, ring_radius)
        self.nr_patch, self.nc_patch = self.patch_pos.shape[:2]
        d = self.d1 * self.d2
        self.A = sp.csc_matrix((d, 0))
        self.C = np.zeros((0, self.T))
        self.C_raw = np.zeros((0, self.T))
        self.S = np.zeros((0, self.T))
        self.A_prev = self.A.copy()
        self.C_prev = self.C.copy()
        self.W, self.b0 = {}, {}
        for mp in self.patches():
            self.W[mp] = ring_W_init(self.patch_pos[mp], self.block_pos[mp], self.d1, self.d2, ring_radius,
                                     self.options["num_neighbors"])
            pp = self.patch_pos[mp]
            self.b0[mp] = np.zeros((pp[1] - pp[0] + 1) * (pp[3] - pp[2] + 1))
        self.P = dict(sn=np.ones((self.d1, self.d2)), kernel_pars=None, neuron_sn=None,
                      Ymean=Y.mean(axis=2, dtype=np.float64))
        self.b0_new = np.zeros((self.d1, self.d2))

    # MATLAB linear patch index order is column-major over (nr_patch, nc_patch)
    def patches(self):
        return [(m, n) for n in range(self.nc_patch) for m in range(self.nr_patch)]

    def _block_mask(self, pos):
        mask = np.zeros((self.d1, self.d2), dtype=bool)
        mask[pos[0] - 1:pos[1], pos[2] - 1:pos[3]] = True
        return mask.ravel(order="F")

    def _get_block(self, pos):
        Yb = self.Y[pos[0] - 1:pos[1], pos[2] - 1:pos[3], :]
        return Yb.reshape(-1, self.T, order="F")

    def reconstruct_b0(self):
        out = np.zeros((self.d1, self.d2))
        for mp in self.patches():
            pp = self.patch_pos[mp]
            out[pp[0] - 1:pp[1], pp[2] - 1:pp[3]] = self.b0[mp].reshape(pp[1] - pp[0] + 1, pp[3] - pp[2] + 1, order="F")
        return out

    def update_background_parallel(self, use_parallel=True):
        o = self.options
        Acsr = sp.csr_matrix(self.A)
        flag_first = is_first_run(self.W[self.patches()[0]])
        for mp in self.patches():
            tb, tp = self.block_pos[mp], self.patch_pos[mp]
            bm = self._block_mask(tb)
            ind = np.asarray(Acsr[bm, :].sum(axis=0)).ravel() > 0
            A_block = Acsr[bm, :][:, ind]
            C_block = self.C[ind, :]
            if A_block.shape[1] == 0 and not flag_first:
                continue
            ind_patch = ind_patch_mask(tp, tb)
            sn_patch = self.P["sn"].ravel(order="F")[bm][ind_patch.ravel(order="F")]
            Ypatch = self._get_block(tb)
            self.W[mp], self.b0[mp] = fit_ring_model(Ypatch, A_block if A_block.shape[1] else None, C_block,
                                                     self.W[mp], o["thresh_outlier"], sn_patch, ind_patch,
                                                     o["bg_acceleration"])
        self.b0_new = self.reconstruct_b0()
        self.A_prev = self.A.copy()
        self.C_prev = self.C.copy()

    def _ysig(self, mp, phase):
        """BG-subtracted patch data.  NB (reference quirk, replicated): the spatial update selects the A_prev
        neurons with `mask(:)==1` AFTER the patch was overwritten with 2 (update_spatial_parallel.m:82-98), i.e.
        only neurons touching the HALO (block minus patch); the temporal update uses the whole block
        (update_temporal_parallel.m:80-92)."""
        tb, tp = self.block_pos[mp], self.patch_pos[mp]
        bm = self._block_mask(tb)
        Apcsr = sp.csr_matrix(self.A_prev)
        sel = bm & ~self._block_mask(tp) if phase == "spatial" else bm
        ind = np.asarray(Apcsr[sel, :].sum(axis=0)).ravel() > 0
        A_prev = Apcsr[bm, :][:, ind]
        C_prev = self.C_prev[ind, :]
        ind_patch = ind_patch_mask(tp, tb)
        return bg_subtract_ring(self._get_block(tb), A_prev, C_prev, self.W[mp], self.b0[mp], ind_patch)

    def update_spatial_parallel(self, use_parallel=True, update_sn=False, IND=None, post_process=None):
        """IND: (d,K) boolean search mask (determine_search_location output, update_spatial_parallel.m:66)."""
        o = self.options
        method = o["spatial_algorithm"]
        d, K = self.A.shape
        IND = sp.csr_matrix(IND).astype(bool)
        Acsr = sp.csr_matrix(self.A)
        A_new = np.zeros((d, K))
        sn_img = self.P["sn"].copy()
        for mp in self.patches():
            tb, tp = self.block_pos[mp], self.patch_pos[mp]
            bm = self._block_mask(tb)
            pm = self._block_mask(tp)
            ind = np.nonzero(np.asarray(IND[pm, :].sum(axis=0)).ravel() > 0)[0]
            if ind.size == 0 and not update_sn:
                continue
            ind_patch = ind_patch_mask(tp, tb).ravel(order="F")
            Ysig = self._ysig(mp, "spatial")
            sn_patch = self.P["sn"].ravel(order="F")[pm]
            if update_sn:
                sn_patch = O.GetSn(Ysig)
                tmp = sn_img.ravel(order="F")
                tmp[pm] = sn_patch
                sn_img = tmp.reshape(self.d1, self.d2, order="F")
            if ind.size == 0:
                continue
            A_patch = np.asarray(Acsr[bm, :][:, ind].todense())[ind_patch, :]
            C_patch = self.C[ind, :]
            IND_patch = np.asarray(IND[pm, :][:, ind].todense())
            if method == "hals":
                temp = HALS_spatial(Ysig, A_patch, C_patch, IND_patch, 3)
            elif method == "hals_thresh":
                temp = HALS_spatial_thresh(Ysig, A_patch, C_patch, IND_patch, 3, sn_patch)
            elif method == "lars":
                temp = lars_spatial(Ysig, A_patch, C_patch, IND_patch, sn_patch)
            else:
                temp = nnls_spatial(Ysig, A_patch, C_patch, IND_patch, 20)
            rows = np.nonzero(pm)[0]
            A_new[np.ix_(rows, ind)] = temp
        if update_sn:
            self.P["sn"] = sn_img
        if post_process is not None:
            A_new = post_process(A_new)
        self.A = sp.csc_matrix(A_new)
        self.b0_new = self.P["Ymean"] - (self.A @ self.C.mean(axis=1)).reshape(self.d1, self.d2, order="F")

    def update_temporal_parallel(self, use_parallel=True, use_c_hat=True):
        o = self.options
        K = self.C.shape[0]
        Acsr = sp.csr_matrix(self.A)
        C_new = np.zeros((K, self.T))
        aa = np.zeros(K)
        dopt = o["deconv_options"] if o["deconv_flag"] else None
        for mp in self.patches():
            tb, tp = self.block_pos[mp], self.patch_pos[mp]
            bm = self._block_mask(tb)
            ind = np.nonzero(np.asarray(Acsr[bm, :].sum(axis=0)).ravel() > 0)[0]
            if ind.size == 0:
                continue
            ind_patch = ind_patch_mask(tp, tb).ravel(order="F")
            A_patch = np.asarray(Acsr[bm, :][:, ind].todense())[ind_patch, :]
            C_patch = self.C[ind, :]
            Ysig = self._ysig(mp, "temporal")
            if not use_c_hat:
                # fast_temporal (update_temporal_parallel.m:314-337)
                with np.errstate(divide="ignore", invalid="ignore"):
                    tmpA = A_patch / A_patch.max(axis=0, keepdims=True)
                tmp_A = A_patch * (tmpA >= 0.5)
                aa_p = np.sum(tmp_A ** 2, axis=0)
                den = np.where(aa_p == 0, np.inf, aa_p)
                C_raw_p = (tmp_A.T @ Ysig) / den[:, None]
            else:
                _, C_raw_p, _, _ = HALS_temporal(Ysig, A_patch, C_patch, o["maxIter"], dopt)
                aa_p = np.sum(A_patch ** 2, axis=0)
            C_new[ind, :] += C_raw_p * aa_p[:, None]
            aa[ind] += aa_p
        aa[aa == 0] = 1
        self.C_raw = C_new / aa[:, None]
        if o["deconv_flag"]:
            self.C = self.deconvTemporal()
        else:
            self.C_raw = self.C_raw - self.C_raw.min(axis=1, keepdims=True)
            self.C = self.C_raw.copy()
        self.b0_new = self.P["Ymean"] - (self.A @ self.C.mean(axis=1)).reshape(self.d1, self.d2, order="F")

    def reconstruct_background(self, frame_range=None):
        """Ybg = reconstruct_background(obj, frame_range) (Sources2D.m:1247-1356), ring model, bg_ssub = 1:
        Ybg(patch) = W * (Y - b0_ - A_prev*C_prev) + b0_new(patch), b0_ = reconstruct_b0() on the block."""
        f0, f1 = (1, self.T) if frame_range is None else (int(frame_range[0]), int(frame_range[1]))
        Tf = f1 - f0 + 1
        b0_ = self.reconstruct_b0().ravel(order="F")
        b0n = np.asarray(self.b0_new).ravel(order="F")
        Apcsr = sp.csr_matrix(self.A_prev)
        Ybg = np.zeros((self.d1 * self.d2, Tf))
        for mp in self.patches():
            tb, tp = self.block_pos[mp], self.patch_pos[mp]
            bm, pm = self._block_mask(tb), self._block_mask(tp)
            ind = np.asarray(Apcsr[bm, :].sum(axis=0)).ravel() > 0
            A_patch = Apcsr[bm, :][:, ind]
            C_patch = self.C_prev[ind, f0 - 1:f1]
            Yp = self._get_block(tb)[:, f0 - 1:f1].astype(np.float64) - b0_[bm][:, None]
            Bf = sp.csr_matrix(self.W[mp]) @ (Yp - A_patch @ C_patch)
            Ybg[pm, :] = Bf + b0n[pm][:, None]
        return Ybg.reshape(self.d1, self.d2, Tf, order="F")

    def compute_RSS(self, frame_range=None):
        """[RSS_total, RSS] = compute_RSS(obj, frame_range) (Sources2D.m:1358-1510), ring model, bg_ssub = 1:
        per patch sum((Y(patch) - A*C) - (W*(Y - b0_ - A_prev*C_prev) + b0_new(patch))).^2."""
        f0, f1 = (1, self.T) if frame_range is None else (int(frame_range[0]), int(frame_range[1]))
        Ybg = self.reconstruct_background((f0, f1)).reshape(self.d1 * self.d2, -1, order="F")
        Acsr = sp.csr_matrix(self.A)
        RSS = {}
        for mp in self.patches():
            tb, tp = self.block_pos[mp], self.patch_pos[mp]
            bm, pm = self._block_mask(tb), self._block_mask(tp)
            ind = np.asarray(Acsr[bm, :].sum(axis=0)).ravel() > 0
            YmAC = self._get_block(tp)[:, f0 - 1:f1].astype(np.float64) - Acsr[pm, :][:, ind] @ self.C[ind, f0 - 1:f1]
            RSS[mp] = float(np.sum((YmAC - Ybg[pm, :]) ** 2))
        total = float(sum(RSS.values()))
        self.P["RSS"] = total
        return total, RSS

    def deconvTemporal(self):
        """deconvTemporal.m:62-105 (serial branch)."""
        K, T = self.C_raw.shape
        C_ = np.zeros((K, T))
        S_ = np.zeros((K, T))
        Craw = np.zeros((K, T))
        kp = [None] * K
        sn = np.zeros(K)
        for k in range(K):
            ck_raw = self.C_raw[k]
            if np.any(np.isnan(ck_raw)):
                continue
            sn[k] = O.GetSn(ck_raw)
            ck, sk, topt = O.deconvolveCa(ck_raw, self.options["deconv_options"], sn=sn[k])
            if np.sum(np.abs(ck)) == 0:
                ck = ck_raw
            C_[k], S_[k] = ck, sk
            kp[k] = np.atleast_1d(topt["pars"]).ravel()
            Craw[k] = ck_raw - topt["b"]
        self.C, self.C_raw, self.S = C_, Craw, S_
        self.P["kernel_pars"] = kp
        self.P["neuron_sn"] = sn
        return C_
