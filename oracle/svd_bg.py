"""
oracle/svd_bg.py -- TEST INFRASTRUCTURE ONLY.  SVD background model of the 2p path (demo_large_data_2p.m:46 default).

  fit_svd_model          ca_source_extraction/endoscope/fit_svd_model.m:1-42 (+ svdsecon.m:18-40: top-nb eigen-pairs of
                         X*X' or X'*X via eigs; restated with a dense SVD -- the pair (b*f) is sign invariant, the
                         individual factors are not, so parity is stated on b*f and on |f|)
  OracleSources2DSVD     the svd branches of update_background_parallel.m:237-243, update_spatial_parallel.m:183-188,
                         update_temporal_parallel.m:169-174 (no halo: tmp_block = patch_pos, :128-132 / :117-121)

Reference quirk replicated: fit_svd_model.m:29 tests the misspelt `thresho_outlier`, so the outlier clamp never runs.
PARITY UNPINNED against MATLAB (no golden vectors; `eigs` is closed source).
"""
import numpy as np
import scipy.sparse as sp

from . import cnmfe as OC


def fit_svd_model(Y, nb, A, C, ind_patch=None):
    Y = np.asarray(Y)
    d, T = Y.shape
    if A is None or np.size(A) == 0:
        A = np.ones((d, 1))
        C = np.zeros((1, T))
    A = np.asarray(A.todense()) if sp.issparse(A) else np.asarray(A, dtype=np.float64)
    C = np.asarray(C, dtype=np.float64)
    if ind_patch is None:
        ind_patch = np.ones(d, dtype=bool)
    ind_patch = np.asarray(ind_patch).ravel(order="F").astype(bool)
    Ymean = Y.mean(axis=1, dtype=np.float64)
    Cmean = C.mean(axis=1)
    Yc = Y.astype(np.float64) - Ymean[:, None]
    Cc = C - Cmean[:, None]
    if nb == 0:
        return np.zeros((ind_patch.sum(), 0)), np.zeros((0, T)), Ymean[ind_patch] - A[ind_patch] @ Cmean
    Bf = Yc - A @ Cc
    X = Bf - Bf.mean(axis=1, keepdims=True)
    u, s, vt = np.linalg.svd(X, full_matrices=False)
    u, s, vt = u[:, :nb], s[:nb], vt[:nb]
    b = u[ind_patch, :] * s[None, :]
    f = vt
    b0 = Ymean[ind_patch] - A[ind_patch, :] @ Cmean - b @ f.mean(axis=1)
    return b, f, b0


class OracleSources2DSVD(OC.OracleSources2D):
    def __init__(self, Y, patch_dims, ring_radius=18, nb=1, options=None):
        super().__init__(Y, patch_dims, ring_radius, options)
        self.options["background_model"] = "svd"
        self.nb = nb
        self.b, self.f = {}, {}
        for mp in self.patches():
            pp = self.patch_pos[mp]
            self.b[mp] = np.zeros(((pp[1] - pp[0] + 1) * (pp[3] - pp[2] + 1), nb))
            self.f[mp] = np.zeros((nb, self.T))

    def update_background_parallel(self, use_parallel=True):
        Acsr = sp.csr_matrix(self.A)
        flag_first = (np.mean(self.b[self.patches()[0]]) == 0)      # mean2(b{1})==0, update_background_parallel.m:145
        for mp in self.patches():
            tb, tp = self.block_pos[mp], self.patch_pos[mp]
            bm = self._block_mask(tb)
            ind = np.asarray(Acsr[bm, :].sum(axis=0)).ravel() > 0
            A_block = Acsr[bm, :][:, ind]
            if A_block.shape[1] == 0 and not flag_first:
                continue
            ind_patch = OC.ind_patch_mask(tp, tb)
            self.b[mp], self.f[mp], self.b0[mp] = fit_svd_model(self._get_block(tb), self.nb,
                                                               A_block if A_block.shape[1] else None, self.C[ind],
                                                               ind_patch)
        self.A_prev = self.A.copy()
        self.C_prev = self.C.copy()

    def _ysig(self, mp, phase):
        tp = self.patch_pos[mp]
        Yp = self._get_block(tp).astype(np.float64)
        return Yp - (self.b[mp] @ self.f[mp] + self.b0[mp][:, None])

    # the spatial / temporal drivers of the base class use block_pos for the neuron selection and the data; the svd
    # model uses the patch itself for both (no halo)
    def update_spatial_parallel(self, use_parallel=True, update_sn=False, IND=None, post_process=None):
        saved = self.block_pos
        self.block_pos = self.patch_pos
        try:
            super().update_spatial_parallel(use_parallel, update_sn, IND, post_process)
        finally:
            self.block_pos = saved

    def update_temporal_parallel(self, use_parallel=True, use_c_hat=True):
        saved = self.block_pos
        self.block_pos = self.patch_pos
        try:
            super().update_temporal_parallel(use_parallel, use_c_hat)
        finally:
            self.block_pos = saved
