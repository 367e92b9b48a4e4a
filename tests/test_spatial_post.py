"""CPU-only: the library's host-side brackets of the spatial update (csrc/host_spatial.cu: connectivity_constraint /
post_process_spatial, determine_search_location 'ellipse') against the NumPy/SciPy restatement in oracle/spatial_post.py,
plus known answers.  These entry points do no device work, so they run without a GPU."""
import numpy as np
import scipy.sparse as sp


class _Host:
    """The two wrapper methods of Sources2D without creating a device context."""
    def __init__(self, lib, d1, d2):
        self._lib, self.d1, self.d2, self.A = lib, d1, d2, None


def _wrap(lib, d1, d2):
    from cnmf_e_b200.sources2d import Sources2D
    h = _Host(lib, d1, d2)
    h.post = lambda A: Sources2D.post_process_spatial(h, A)
    h.search = lambda A, **kw: Sources2D.determine_search_location(h, A, **kw)
    return h


def _footprints(rng, d1, d2, K):
    cols = []
    rr, cc = np.meshgrid(np.arange(d1), np.arange(d2), indexing="ij")
    for k in range(K):
        r0, c0 = rng.uniform(-2, d1 + 2), rng.uniform(-2, d2 + 2)          # some centres at / beyond the border
        sr, sc, th = rng.uniform(1.5, 4.0), rng.uniform(1.5, 4.0), rng.uniform(0, np.pi)
        x, y = (rr - r0) * np.cos(th) + (cc - c0) * np.sin(th), -(rr - r0) * np.sin(th) + (cc - c0) * np.cos(th)
        a = np.exp(-0.5 * (x / sr) ** 2 - 0.5 * (y / sc) ** 2) * rng.uniform(5, 20)
        a[a < 0.05 * a.max()] = 0
        # a detached blob and isolated specks that the constraint must remove
        r1, c1 = int(rng.integers(0, d1)), int(rng.integers(0, d2))
        a[max(0, r1 - 3):r1 + 4, max(0, c1 - 3):c1 + 4] += 0.3 * a.max()
        for _ in range(6):
            a[int(rng.integers(0, d1)), int(rng.integers(0, d2))] += 0.5 * a.max()
        cols.append(sp.csc_matrix(a.reshape(-1, 1, order="F")))
    return sp.hstack(cols, format="csc")


def test_connectivity_constraint_matches_oracle(built_lib):
    from oracle import spatial_post as OP
    rng = np.random.default_rng(11)
    for d1, d2, K in [(40, 36, 12), (25, 61, 8)]:
        A = _footprints(rng, d1, d2, K)
        h = _wrap(built_lib, d1, d2)
        got = h.post(A)
        ref = OP.post_process_spatial(A, d1, d2)
        assert got.shape == ref.shape
        assert np.array_equal(got.toarray(), ref.toarray())
        assert got.nnz < A.nnz                                             # something was removed


def test_connectivity_constraint_known_answers(built_lib):
    d1 = d2 = 20
    h = _wrap(built_lib, d1, d2)
    img = np.zeros((d1, d2))
    img[3:10, 4:11] = 2.0          # 7 x 7 plateau: survives the 5 x 5 opening
    img[5, 6] = 5.0                # the maximum, inside the plateau
    img[15, 15] = 4.0              # isolated pixel: removed
    img[12:15, 1:4] = 3.0          # 3 x 3 block: does not survive the opening, not in the mask, hence removed
    A = sp.csc_matrix(img.reshape(-1, 1, order="F"))
    out = h.post(A).toarray().reshape(d1, d2, order="F")
    exp = np.zeros((d1, d2)); exp[3:10, 4:11] = 2.0; exp[5, 6] = 5.0
    assert np.array_equal(out, exp)
    # maximum on a speck that the opening removes: l(ind_max) = 0 -> every labelled component goes, the rest stays (quirk)
    img2 = np.zeros((d1, d2)); img2[3:10, 4:11] = 2.0; img2[15, 15] = 9.0
    out2 = h.post(sp.csc_matrix(img2.reshape(-1, 1, order="F"))).toarray().reshape(d1, d2, order="F")
    exp2 = np.zeros((d1, d2)); exp2[15, 15] = 9.0
    assert np.array_equal(out2, exp2)
    # empty column and K = 0
    assert h.post(sp.csc_matrix((d1 * d2, 1))).nnz == 0
    assert h.post(sp.csc_matrix((d1 * d2, 0))).shape == (d1 * d2, 0)


def test_search_location_ellipse_matches_oracle(built_lib):
    from oracle import spatial_post as OP
    rng = np.random.default_rng(5)
    for d1, d2, K in [(48, 40, 10), (30, 70, 7)]:
        A = _footprints(rng, d1, d2, K)
        A = sp.hstack([A, sp.csc_matrix((d1 * d2, 1))], format="csc")      # plus an empty component
        h = _wrap(built_lib, d1, d2)
        for kw in (dict(), dict(min_size=2.0, max_size=5.0, dist=2.5)):
            got = h.search(A, **kw)
            ref = OP.determine_search_location(A, d1, d2, **kw)
            assert got.shape == ref.shape
            diff = (got.astype(np.int8) - ref.astype(np.int8))
            assert diff.nnz == 0, "%d mask pixels differ" % diff.nnz
            assert np.all(np.diff(got.indptr) > 0)


def test_search_location_isotropic_known_answer(built_lib):
    d1 = d2 = 41
    h = _wrap(built_lib, d1, d2)
    img = np.zeros((d1, d2)); img[20, 20] = 1.0                            # a single pixel: zero variance -> min_size circle
    IND = h.search(sp.csc_matrix(img.reshape(-1, 1, order="F")))
    m = IND.toarray().reshape(d1, d2, order="F")
    rr, cc = np.meshgrid(np.arange(d1), np.arange(d2), indexing="ij")
    assert np.array_equal(m, np.sqrt((rr - 20.0) ** 2 + (cc - 20.0) ** 2) / 3.0 <= 3.0)


def test_graph_conn_comp_matches_scipy(built_lib):
    """cnmfe_graph_conn_comp = the reference's graph_conn_comp_mex: labels 1..c ordered by each component's smallest node."""
    import ctypes
    from scipy.sparse.csgraph import connected_components
    from cnmf_e_b200 import _lib as L
    rng = np.random.default_rng(3)
    for n, dens in [(1, 0.0), (12, 0.08), (60, 0.03), (200, 0.004)]:
        M = sp.random(n, n, density=dens, random_state=int(rng.integers(1 << 30)), format="csc")
        S = sp.csc_matrix(((M + M.T) != 0).astype(float))           # symmetric flag matrix, as the merge routines build it
        S.sort_indices()
        jc, ir = S.indptr.astype(np.int64), S.indices.astype(np.int64)
        lab = np.zeros(n, np.uint32)
        nc = ctypes.c_int(0)
        L.check(built_lib.cnmfe_graph_conn_comp(n, jc.ctypes.data_as(ctypes.c_void_p), ir.ctypes.data_as(ctypes.c_void_p),
                                                lab.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nc)))
        c_ref, l_ref = connected_components(S, directed=False)
        assert nc.value == c_ref
        # same partition, and our numbering follows the smallest node of each component
        first = {}
        for i, l in enumerate(lab):
            first.setdefault(int(l), i)
        assert sorted(first) == list(range(1, c_ref + 1))
        assert [first[l] for l in range(1, c_ref + 1)] == sorted(first.values())
        for i in range(n):
            for j in range(i + 1, n):
                assert (lab[i] == lab[j]) == (l_ref[i] == l_ref[j])
    # a non-symmetric matrix that reaches an earlier component is an error, as in the reference
    A = sp.csc_matrix(np.array([[0, 0, 0], [0, 0, 0], [1, 0, 0.]]).T)   # column 2 lists row 0: node 2 -> node 0
    A.sort_indices()
    jc, ir = A.indptr.astype(np.int64), A.indices.astype(np.int64)
    lab = np.zeros(3, np.uint32); nc = ctypes.c_int(0)
    rc = built_lib.cnmfe_graph_conn_comp(3, jc.ctypes.data_as(ctypes.c_void_p), ir.ctypes.data_as(ctypes.c_void_p),
                                         lab.ctypes.data_as(ctypes.c_void_p), ctypes.byref(nc))
    assert rc != 0 and b"mixed labeling" in built_lib.cnmfe_last_error()


def test_circular_constraints_matches_oracle(built_lib):
    from oracle import spatial_post as OP
    from cnmf_e_b200.sources2d import Sources2D
    rng = np.random.default_rng(23)
    for d1, d2, K in [(40, 36, 10), (25, 61, 6)]:
        A = _footprints(rng, d1, d2, K)
        A = sp.hstack([A, sp.csc_matrix((d1 * d2, 1))], format="csc")          # an empty neuron
        line = np.zeros((d1, d2)); line[7, 3:9] = 1.0                           # a single row: left unchanged
        A = sp.hstack([A, sp.csc_matrix(line.reshape(-1, 1, order="F"))], format="csc")
        h = _Host(built_lib, d1, d2)
        for connected in (False, True):
            got = Sources2D.post_process_spatial(h, A, connected=connected, circular=True)
            ref = OP.post_process_spatial(A, d1, d2, connected=connected, circular=True)
            assert np.array_equal(got.toarray(), ref.toarray())
        assert np.array_equal(Sources2D.post_process_spatial(h, A, connected=False, circular=True)[:, -1].toarray().ravel(),
                              line.ravel(order="F"))


def test_search_location_dilate_matches_oracle(built_lib):
    from oracle import spatial_post as OP
    from cnmf_e_b200.sources2d import Sources2D
    rng = np.random.default_rng(41)
    for d1, d2, K in [(44, 38, 9), (26, 58, 6)]:
        A = _footprints(rng, d1, d2, K)
        A = sp.hstack([A, sp.csc_matrix((d1 * d2, 1)), A[:, :1]], format="csc")     # an empty neuron; last column = "nb" column
        h = _Host(built_lib, d1, d2)
        for kw in (dict(), dict(nrgthr=0.99, nb=0, bSiz=4), dict(nb=2, bSiz=2)):
            got = Sources2D.determine_search_location(h, A, method="dilate", **kw)
            ref = OP.determine_search_location_dilate(A, d1, d2, **kw)
            assert got.shape == ref.shape
            diff = got.astype(np.int8) - ref.astype(np.int8)
            assert diff.nnz == 0, "%d mask pixels differ (%s)" % (diff.nnz, kw)
        # the thresholded mask contains the bulk of each real footprint
        IND = Sources2D.determine_search_location(h, A, method="dilate")
        for k in range(K):
            a = A[:, k].toarray().ravel()
            assert IND[:, k].toarray().ravel()[np.argmax(a)]
