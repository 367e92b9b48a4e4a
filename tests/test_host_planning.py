"""CPU-only: the host planning of the update calls (csrc/ctx.cu build_local: block / patch-local CSR + CSC views of A, A_prev
and the search mask, neuron selection rules of update_*_parallel.m) against a direct NumPy restatement, through the
cnmfe_debug_local_view test hook (no device work)."""
import ctypes
import numpy as np
import pytest
import scipy.sparse as sp

SEL_SUM_BLOCK, SEL_SUM_HALO, SEL_ANY_PATCH = 0, 1, 2
ROWS_BLOCK, ROWS_BLOCK_PATCHONLY, ROWS_PATCH = 0, 1, 2


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _local_view(lib, d1, d2, patch, block, M, sel, rows, vals=None):
    from cnmf_e_b200 import _lib as L
    M = sp.csc_matrix(M); M.sort_indices()
    K, nnz = M.shape[1], M.nnz
    jc, ir, pr = M.indptr.astype(np.int64), M.indices.astype(np.int64), M.data.astype(np.float64)
    vj = vi = vp = None
    if vals is not None:
        V = sp.csc_matrix(vals, dtype=np.float64); V.sort_indices()
        vj, vi, vp = V.indptr.astype(np.int64), V.indices.astype(np.int64), V.data.astype(np.float64)
    nrb, ncb = block[1] - block[0] + 1, block[3] - block[2] + 1
    nr, nc = patch[1] - patch[0] + 1, patch[3] - patch[2] + 1
    nrows = nr * nc if rows == ROWS_PATCH else nrb * ncb
    nl, nk = ctypes.c_int(0), ctypes.c_int(0)
    ids = np.zeros(max(K, 1), np.int32); ptr = np.zeros(nrows + 1, np.int32)
    col = np.zeros(max(nnz, 1), np.int32); val = np.zeros(max(nnz, 1)); src = np.zeros(max(nnz, 1), np.int64)
    cptr = np.zeros(K + 1, np.int32); crow = np.zeros(max(nnz, 1), np.int32); cval = np.zeros(max(nnz, 1))
    bbox = np.zeros(4 * max(K, 1), np.int32)
    pp, bp = np.array(patch, np.int32), np.array(block, np.int32)
    L.check(lib.cnmfe_debug_local_view(d1, d2, _p(pp), _p(bp), K, _p(jc), _p(ir), _p(pr), sel, rows, _p(vj), _p(vi), _p(vp),
                                       ctypes.byref(nl), ctypes.byref(nk), _p(ids), _p(ptr), _p(col), _p(val), _p(src),
                                       _p(cptr), _p(crow), _p(cval), _p(bbox)))
    n, k = nl.value, nk.value
    return dict(ids=ids[:n], ptr=ptr, col=col[:k], val=val[:k], src=src[:k], cptr=cptr[:n + 1], crow=crow[:k], cval=cval[:k],
                bbox=bbox[:4 * n].reshape(n, 4), nrows=nrows)


def _reference(d1, d2, patch, block, M, sel, rows, vals=None):
    """Direct restatement: dense masks, no cleverness (1-based inclusive positions as in distribute_data.m:163-173)."""
    M = sp.csc_matrix(M); M.sort_indices()
    Md = M.toarray()
    pat = (M != 0).toarray() | np.zeros(M.shape, bool)
    stored = np.zeros(M.shape, bool)
    stored[M.indices, np.repeat(np.arange(M.shape[1]), np.diff(M.indptr))] = True      # explicit entries, zeros included
    Vd = Md if vals is None else sp.csc_matrix(vals).toarray()
    r = np.tile(np.arange(d1), d2); c = np.repeat(np.arange(d2), d1)
    inb = (r >= block[0] - 1) & (r <= block[1] - 1) & (c >= block[2] - 1) & (c <= block[3] - 1)
    inp = (r >= patch[0] - 1) & (r <= patch[1] - 1) & (c >= patch[2] - 1) & (c <= patch[3] - 1)
    nrb = block[1] - block[0] + 1
    nr = patch[1] - patch[0] + 1
    idx_block = (c - (block[2] - 1)) * nrb + (r - (block[0] - 1))
    idx_patch = (c - (patch[2] - 1)) * nr + (r - (patch[0] - 1))
    ids, cols = [], []
    for k in range(M.shape[1]):
        if sel == SEL_SUM_BLOCK:
            take = Md[inb, k].sum() > 0
        elif sel == SEL_SUM_HALO:
            take = Md[inb & ~inp, k].sum() > 0
        else:
            take = bool(stored[inp, k].any())
        if not take:
            continue
        keep = stored[:, k] & (inp if rows in (ROWS_PATCH, ROWS_BLOCK_PATCHONLY) else inb)
        px = np.nonzero(keep)[0]
        li = idx_patch[px] if rows == ROWS_PATCH else idx_block[px]
        ids.append(k)
        cols.append((li, Vd[px, k], px))
    return ids, cols


@pytest.mark.parametrize("geom", [
    # d1, d2, patch (r0 r1 c0 c1), block
    (40, 36, (1, 40, 1, 36), (1, 40, 1, 36)),
    (48, 60, (1, 24, 31, 60), (1, 31, 24, 60)),
    (48, 60, (25, 48, 1, 30), (18, 48, 1, 37)),
])
def test_local_view_matches_restatement(built_lib, geom):
    d1, d2, patch, block = geom
    rng = np.random.default_rng(d1 * 7 + patch[2])
    K = 25
    A = sp.random(d1 * d2, K, density=0.0, format="lil")
    rr, cc = np.meshgrid(np.arange(d1), np.arange(d2), indexing="ij")
    for k in range(K):
        r0, c0 = rng.uniform(0, d1), rng.uniform(0, d2)
        m = ((rr - r0) ** 2 + (cc - c0) ** 2 <= rng.uniform(2, 5) ** 2)
        A[(rr[m] + cc[m] * d1), k] = rng.uniform(0.1, 2.0, m.sum())
    A = sp.csc_matrix(A)
    A = sp.hstack([A, sp.csc_matrix((d1 * d2, 2))], format="csc")                      # two empty neurons
    IND = sp.csc_matrix((A != 0).astype(float))
    IND = sp.csc_matrix(sp.vstack([IND[1:], IND[:1]]) + IND)                            # a mask that is not the support of A
    IND.data[:] = 1.0
    cases = [(A, SEL_SUM_BLOCK, ROWS_BLOCK, None), (A, SEL_SUM_BLOCK, ROWS_BLOCK_PATCHONLY, None),
             (A, SEL_SUM_HALO, ROWS_BLOCK, None), (IND, SEL_ANY_PATCH, ROWS_PATCH, A)]
    for M, sel, rows, vals in cases:
        got = _local_view(built_lib, d1, d2, patch, block, M, sel, rows, vals)
        ids, cols = _reference(d1, d2, patch, block, M, sel, rows, vals)
        assert list(got["ids"]) == ids
        Ms = sp.csc_matrix(M); Ms.sort_indices()
        nrw = (patch[1] - patch[0] + 1) if rows == ROWS_PATCH else (block[1] - block[0] + 1)
        for j, (li, v, px) in enumerate(cols):
            a, b = got["cptr"][j], got["cptr"][j + 1]
            assert np.array_equal(got["crow"][a:b], li)
            assert np.array_equal(got["cval"][a:b], v)
            r, c = li % nrw, li // nrw
            if rows == ROWS_PATCH:
                r, c = r + patch[0] - block[0], c + patch[2] - block[2]
            exp = [r.min(), r.max(), c.min(), c.max()] if li.size else [0, -1, 0, -1]
            assert list(got["bbox"][j]) == exp
        # CSR by pixel: same entries, ascending local neuron id inside each row; entry_src points back into the CSC
        dense = np.zeros((got["nrows"], max(len(ids), 1)))
        seen = np.zeros_like(dense, bool)
        for j, (li, v, px) in enumerate(cols):
            dense[li, j] = v; seen[li, j] = True
        for i in range(got["nrows"]):
            a, b = got["ptr"][i], got["ptr"][i + 1]
            cs = got["col"][a:b]
            assert np.all(np.diff(cs) > 0)
            assert np.array_equal(np.nonzero(seen[i])[0], cs)
            assert np.array_equal(got["val"][a:b], dense[i, cs])
            for e, j in zip(got["src"][a:b], cs):
                assert Ms.indices[e] in cols[j][2] and np.searchsorted(Ms.indptr, e, side="right") - 1 == ids[j]
