"""
oracle/ssub.py -- TEST INFRASTRUCTURE ONLY.  Ring background model with spatial down-sampling (bg_ssub > 1), the
setting of demos/demo_large_data_1p.m:30 (bg_ssub = 2).

  imresize / contributions   MathWorks `imresize` (Image Processing Toolbox, closed source here): restated from its
                             published algorithm -- separable, per-dimension weight tables
                             u = x/scale + 0.5(1-1/scale); left = floor(u - w/2); P = ceil(w)+2 taps; weights = h(u-idx)
                             normalised to 1; mirror ("symmetric") extension of the indices; for scale < 1 with
                             antialiasing h(x) = scale*k(scale*x), w = w/scale; cubic kernel a = -0.5; 'nearest' = box
                             kernel without antialiasing; output size ceil(in*scale); dimensions resized in order of
                             increasing scale.
  OracleSources2DSsub        update_background_parallel.m:70-118,220-227 (W on the ceil(block/ssub) grid, fit on the
                             NEAREST-down-sampled residual, b0 = mean residual on the patch) and
                             update_spatial_parallel.m:167-177 == update_temporal_parallel.m:154-163 (BG reconstruction:
                             bicubic+antialias down, W, bicubic up).

PARITY UNPINNED against MATLAB (imresize restated from memory of the documented algorithm; no golden vectors).
"""
import numpy as np
import scipy.sparse as sp

from . import cnmfe as OC


def _cubic(x):
    ax = np.abs(x)
    ax2, ax3 = ax * ax, ax * ax * ax
    return (1.5 * ax3 - 2.5 * ax2 + 1) * (ax <= 1) + (-0.5 * ax3 + 2.5 * ax2 - 4 * ax + 2) * ((1 < ax) & (ax <= 2))


def _box(x):
    return ((-0.5 <= x) & (x < 0.5)).astype(np.float64)


def contributions(in_len, out_len, scale, method="bicubic", antialias=True):
    """Weight table of one dimension: returns (weights (out_len,P), indices (out_len,P) 0-based)."""
    if method == "nearest":
        kernel, width, antialias = _box, 1.0, False
    else:
        kernel, width = _cubic, 4.0
    if scale < 1 and antialias:
        h = lambda x: scale * kernel(scale * x)
        width = width / scale
    else:
        h = kernel
    x = np.arange(1, out_len + 1, dtype=np.float64)[:, None]
    u = x / scale + 0.5 * (1 - 1 / scale)
    left = np.floor(u - width / 2)
    P = int(np.ceil(width)) + 2
    indices = left + np.arange(P)[None, :]
    weights = h(u - indices)
    weights = weights / weights.sum(axis=1, keepdims=True)
    aux = np.concatenate([np.arange(1, in_len + 1), np.arange(in_len, 0, -1)])
    indices = aux[np.mod(indices.astype(np.int64) - 1, aux.size)]
    keep = np.any(weights != 0, axis=0)
    return weights[:, keep], indices[:, keep] - 1


def resize_matrix(in_len, out_len, scale, method="bicubic"):
    """The same table as a dense (out_len, in_len) matrix (duplicate indices from the mirror extension add up)."""
    w, idx = contributions(in_len, out_len, scale, method)
    M = np.zeros((out_len, in_len))
    for i in range(out_len):
        np.add.at(M[i], idx[i], w[i])
    return M


def imresize(X, scale=None, out_size=None, method="bicubic"):
    """imresize(X, scale) or imresize(X, [m n]) on the first two dimensions of X."""
    X = np.asarray(X, dtype=np.float64)
    n1, n2 = X.shape[:2]
    if out_size is None:
        m1, m2 = int(np.ceil(n1 * scale)), int(np.ceil(n2 * scale))
        s1 = s2 = float(scale)
    else:
        m1, m2 = out_size
        s1, s2 = m1 / n1, m2 / n2
    R1 = resize_matrix(n1, m1, s1, method)
    R2 = resize_matrix(n2, m2, s2, method)
    order = [0, 1] if s1 <= s2 else [1, 0]
    out = X
    for dim in order:
        if dim == 0:
            out = np.tensordot(R1, out, axes=(1, 0))
        else:
            out = np.moveaxis(np.tensordot(R2, out, axes=(1, 1)), 0, 1)
    return out


def ring_W_init_ssub(nrb, ncb, ring_radius, ssub, num_neighbors=None):
    """initComponents_parallel.m:237-253: uniform ring on the ceil(block/ssub) grid, neighbours inside that grid."""
    d1s, d2s = int(np.ceil(nrb / ssub)), int(np.ceil(ncb / ssub))
    r_shift, c_shift = OC.get_nhood(int(np.ceil(ring_radius / ssub)), num_neighbors)
    csub, rsub = np.meshgrid(np.arange(1, d2s + 1), np.arange(1, d1s + 1))
    csub = csub.T.reshape(-1, 1)
    rsub = rsub.T.reshape(-1, 1)
    ii = np.repeat(np.arange(csub.size).reshape(-1, 1), r_shift.size, axis=1)
    cs = csub + c_shift.reshape(1, -1)
    rs = rsub + r_shift.reshape(1, -1)
    ind = (cs >= 1) & (cs <= d2s) & (rs >= 1) & (rs <= d1s)
    jj = (cs - 1) * d1s + rs - 1
    temp = sp.csr_matrix((np.ones(ind.sum()), (ii[ind], jj[ind])), shape=(d1s * d2s, d1s * d2s))
    rowsum = np.asarray(temp.sum(axis=1)).ravel()
    return sp.diags(1.0 / rowsum) @ temp


class OracleSources2DSsub(OC.OracleSources2D):
    def __init__(self, Y, patch_dims, ring_radius=18, bg_ssub=2, options=None):
        super().__init__(Y, patch_dims, ring_radius, options)
        self.options["bg_ssub"] = bg_ssub
        self.ssub = bg_ssub
        for mp in self.patches():
            tb = self.block_pos[mp]
            self.W[mp] = ring_W_init_ssub(tb[1] - tb[0] + 1, tb[3] - tb[2] + 1, ring_radius, bg_ssub,
                                          self.options["num_neighbors"])

    def update_background_parallel(self, use_parallel=True):
        o = self.options
        Acsr = sp.csr_matrix(self.A)
        flag_first = OC.is_first_run(self.W[self.patches()[0]])
        for mp in self.patches():
            tb, tp = self.block_pos[mp], self.patch_pos[mp]
            bm = self._block_mask(tb)
            ind = np.asarray(Acsr[bm, :].sum(axis=0)).ravel() > 0
            A_block = Acsr[bm, :][:, ind]
            C_block = self.C[ind, :]
            if A_block.shape[1] == 0 and not flag_first:
                continue
            nrb, ncb = tb[1] - tb[0] + 1, tb[3] - tb[2] + 1
            ind_patch = OC.ind_patch_mask(tp, tb)
            Yb = self._get_block(tb).astype(np.float64)
            temp = (Yb - (A_block @ C_block if A_block.shape[1] else 0.0)).reshape(nrb, ncb, self.T, order="F")
            tmp_b0 = temp.mean(axis=2)
            self.b0[mp] = tmp_b0.ravel(order="F")[ind_patch.ravel(order="F")]
            Yds = imresize(temp, 1.0 / self.ssub, method="nearest")
            Yds = Yds.reshape(-1, self.T, order="F")
            self.W[mp], _ = OC.fit_ring_model(Yds, None, None, self.W[mp], o["thresh_outlier"], None, None,
                                              o["bg_acceleration"])
        self.b0_new = self.reconstruct_b0()
        self.A_prev = self.A.copy()
        self.C_prev = self.C.copy()

    def _ysig(self, mp, phase):
        tb, tp = self.block_pos[mp], self.patch_pos[mp]
        bm = self._block_mask(tb)
        Apcsr = sp.csr_matrix(self.A_prev)
        sel = bm & ~self._block_mask(tp) if phase == "spatial" else bm
        ind = np.asarray(Apcsr[sel, :].sum(axis=0)).ravel() > 0
        A_prev = Apcsr[bm, :][:, ind]
        C_prev = self.C_prev[ind, :]
        ind_patch = OC.ind_patch_mask(tp, tb).ravel(order="F")
        nrb, ncb = tb[1] - tb[0] + 1, tb[3] - tb[2] + 1
        Yd = self._get_block(tb).astype(np.float64)
        tmp_Y = Yd - (A_prev @ C_prev if A_prev.shape[1] else 0.0)
        temp = (tmp_Y - tmp_Y.mean(axis=1, keepdims=True)).reshape(nrb, ncb, self.T, order="F")
        temp = imresize(temp, 1.0 / self.ssub)
        d1s, d2s = temp.shape[:2]
        Bf = (sp.csr_matrix(self.W[mp]) @ temp.reshape(-1, self.T, order="F")).reshape(d1s, d2s, self.T, order="F")
        Bf = imresize(Bf, out_size=(nrb, ncb)).reshape(-1, self.T, order="F")
        return Yd[ind_patch] - Bf[ind_patch] - np.asarray(self.b0[mp]).reshape(-1, 1)
