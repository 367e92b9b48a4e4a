"""
oracle/nmf_bg.py -- TEST INFRASTRUCTURE ONLY.  nmf background model (demo_large_data_2p.m with background_model = 'nmf').

  nnmf_als               MathWorks Statistics and Machine Learning Toolbox `nnmf(a, k)` -- CLOSED SOURCE, not under
                         /root/reference (SURVEY.md 8c lists it; no version pinned: README.md:18-24 only names the toolbox).
                         Restated from its published algorithm (doc page "nnmf", default options): 'als' algorithm,
                         1 replicate, w0 = rand(n,k), h0 = rand(k,m), MaxIter 100, TolFun 1e-4, TolX 1e-4; per iteration
                         h = max(0, w0\\a), w = max(0, a/h); dnorm = sqrt(sum(sum((a-w*h).^2))/numel(a));
                         delta = max(max|w-w0|/(sqrt(eps)+max|w0|), max|h-h0|/(sqrt(eps)+max|h0|)); after the first
                         iteration stop when delta <= TolX, or dnorm0-dnorm <= TolFun*max(1,dnorm0), or MaxIter; then the rows
                         of h are scaled to unit length (w scaled inversely) and the components sorted by descending sum(w.^2).
                         The random start is replaced by a fixed hash sequence shared with the CUDA path (hash_uniform).
  fit_nmf_model          ca_source_extraction/endoscope/fit_nmf_model.m:1-23 (B = Y - A*C, outlier clamp :11-15 -- a no-op for
                         the default thresh_outlier = NaN --, [b, f] = nnmf(B, nb), b = b(ind_patch, :))
  OracleSources2DNMF     the nmf branches of update_background_parallel.m:231-236, update_spatial_parallel.m:179-182,
                         update_temporal_parallel.m:165-168

PARITY UNPINNED against MATLAB: nnmf draws its start from MATLAB's global random stream, so the reference result is itself
random at the level of the stopping tolerance (1e-4); parity GPU <-> oracle is stated for the shared fixed start.
"""
import numpy as np
import scipy.sparse as sp

from . import cnmfe as OC
from .svd_bg import OracleSources2DSVD


def hash_uniform(idx):
    """murmur3 finaliser of a uint32 index -> (0, 1); identical to nmf_hash_uniform in csrc/kernels_svd.cuh."""
    x = np.asarray(idx, dtype=np.uint64) & 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x85EBCA6B) & 0xFFFFFFFF
    x ^= x >> 13
    x = (x * 0xC2B2AE35) & 0xFFFFFFFF
    x ^= x >> 16
    return ((x >> 8).astype(np.float64) + 0.5) / 16777216.0


def nnmf_als(a, k, maxiter=100, tolfun=1e-4, tolx=1e-4):
    a = np.asarray(a, dtype=np.float64)
    n, m = a.shape
    w0 = hash_uniform(np.arange(n * k, dtype=np.uint64)).reshape(k, n).T.copy()                   # column-major fill of rand(n,k)
    h0 = hash_uniform(0x9E3779B9 + np.arange(k * m, dtype=np.uint64)).reshape(m, k).T.copy()      # column-major fill of rand(k,m)
    nm = a.size
    sqrteps = np.sqrt(np.finfo(np.float64).eps)
    dnorm0 = 0.0
    for it in range(1, maxiter + 1):
        h = np.maximum(0.0, np.linalg.lstsq(w0, a, rcond=None)[0])
        w = np.maximum(0.0, np.linalg.lstsq(h.T, a.T, rcond=None)[0].T)
        d = a - w @ h
        dnorm = np.sqrt(np.sum(d * d) / nm)
        dw = np.max(np.abs(w - w0) / (sqrteps + np.max(np.abs(w0))))
        dh = np.max(np.abs(h - h0) / (sqrteps + np.max(np.abs(h0))))
        delta = max(dw, dh)
        if it > 1:
            if delta <= tolx:
                break
            elif dnorm0 - dnorm <= tolfun * max(1.0, dnorm0):
                break
            elif it == maxiter:
                break
        dnorm0, w0, h0 = dnorm, w, h
    hlen = np.sqrt(np.sum(h * h, axis=1))
    hlen[hlen == 0] = 1.0
    w = w * hlen[None, :]
    h = h / hlen[:, None]
    idx = np.argsort(-np.sum(w * w, axis=0), kind="stable")
    return w[:, idx], h[idx, :], it


def fit_nmf_model(Y, nb, A, C, ind_patch=None):
    Y = np.asarray(Y)
    d, T = Y.shape
    B = Y.astype(np.float64)
    if A is not None and np.size(A) > 0:
        A = np.asarray(A.todense()) if sp.issparse(A) else np.asarray(A, dtype=np.float64)
        B = B - A @ np.asarray(C, dtype=np.float64)
    if ind_patch is None:
        ind_patch = np.ones(d, dtype=bool)
    ind_patch = np.asarray(ind_patch).ravel(order="F").astype(bool)
    b, f, it = nnmf_als(B, nb)
    return b[ind_patch, :], f, it


class OracleSources2DNMF(OracleSources2DSVD):
    def __init__(self, Y, patch_dims, ring_radius=18, nb=1, options=None):
        super().__init__(Y, patch_dims, ring_radius, nb, options)
        self.options["background_model"] = "nmf"
        self.iters = {}

    def update_background_parallel(self, use_parallel=True):
        Acsr = sp.csr_matrix(self.A)
        flag_first = (np.mean(self.b[self.patches()[0]]) == 0)
        for mp in self.patches():
            tb, tp = self.block_pos[mp], self.patch_pos[mp]
            bm = self._block_mask(tb)
            ind = np.asarray(Acsr[bm, :].sum(axis=0)).ravel() > 0
            A_block = Acsr[bm, :][:, ind]
            if A_block.shape[1] == 0 and not flag_first:
                continue
            self.b[mp], self.f[mp], self.iters[mp] = fit_nmf_model(self._get_block(tb), self.nb,
                                                                   A_block if A_block.shape[1] else None, self.C[ind],
                                                                   OC.ind_patch_mask(tp, tb))
        self.A_prev = self.A.copy()
        self.C_prev = self.C.copy()

    def _ysig(self, mp, phase):
        return self._get_block(self.patch_pos[mp]).astype(np.float64) - self.b[mp] @ self.f[mp]
