"""TEST INFRASTRUCTURE ONLY -- float64 NumPy/SciPy restatement of the two routines that bracket the spatial update
(PARITY UNPINNED: the reference holds no fixtures for them and MATLAB is not available here):

  connectivity_constraint   /root/reference/ca_source_extraction/endoscope/connectivity_constraint.m:1-20
                            (called by @Sources2D/post_process_spatial.m:19-32 when spatial_constraints.connected)
  determine_search_location /root/reference/ca_source_extraction/utilities/determine_search_location.m:57-92 ('ellipse')
                            with com() from utilities/com.m

MathWorks toolbox behaviour restated from its documentation: imopen = imdilate(imerode(.)) with a flat 'square' structuring
element; pixels outside the image never win (erosion pads +Inf, dilation -Inf); bwlabel(., 4) = 4-connected components.
"""
import numpy as np
import scipy.ndimage as ndi
import scipy.sparse as sp


def connectivity_constraint(img, thr=0.01, sz=5):
    """img: (d1, d2) float array.  Returns the constrained copy (connectivity_constraint.m:12-20)."""
    img = np.array(img, dtype=np.float64)
    ind_max = int(np.argmax(img.ravel(order="F")))                      # [~, ind_max] = max(img(:))
    ero = ndi.grey_erosion(img, size=(sz, sz), mode="constant", cval=np.inf)
    ai_open = ndi.grey_dilation(ero, size=(sz, sz), mode="constant", cval=-np.inf)
    temp = ai_open > img.max() * thr
    lab, _ = ndi.label(temp, structure=[[0, 1, 0], [1, 1, 1], [0, 1, 0]])
    l_max = lab.ravel(order="F")[ind_max]
    out = img.copy()
    out[lab != l_max] = 0.0
    return out


def circular_constraints(img):
    """endoscope/circular_constraints.m:1-54 (show_imgs = false)."""
    img = np.array(img, dtype=np.float64)
    rr, cc = np.nonzero(img)
    if rr.size == 0:
        return img
    rmin, rmax, cmin, cmax = rr.min(), rr.max(), cc.min(), cc.max()
    if (rmax - rmin < 1) or (cmax - cmin < 1):
        return img
    sub = img[rmin:rmax + 1, cmin:cmax + 1].copy()
    nr, nc = sub.shape
    ind_max = int(np.argmax(sub.ravel(order="F")))
    vmax = sub.ravel(order="F")[ind_max]
    y0, x0 = ind_max % nr, ind_max // nr
    x, y = np.meshgrid(np.arange(nc), np.arange(nr))
    fy, fx = np.gradient(sub)                                             # MATLAB: [fx, fy] = gradient(img)
    ind = ((fx * (x0 - x) + fy * (y0 - y)) < 0) & (sub < vmax / 3)
    sub[ind] = 0
    lab, _ = ndi.label(sub != 0, structure=[[0, 1, 0], [1, 1, 1], [0, 1, 0]])
    keep = ndi.binary_dilation(lab == lab[y0, x0], structure=np.ones((3, 3), bool))
    sub[~keep] = 0
    sub = ndi.median_filter(sub, size=3, mode="constant", cval=0.0)      # medfilt2: zero padding
    img[rmin:rmax + 1, cmin:cmax + 1] = sub
    return img


def post_process_spatial(A, d1, d2, connected=True, circular=False):
    """A: (d1*d2, K) sparse.  post_process_spatial.m:19-32."""
    A = sp.csc_matrix(A, dtype=np.float64)
    cols = []
    for k in range(A.shape[1]):
        ai = np.asarray(A[:, k].todense()).reshape(d1, d2, order="F")
        if connected:
            ai = connectivity_constraint(ai)
        if circular:
            ai = circular_constraints(ai)
        cols.append(sp.csc_matrix(ai.reshape(-1, 1, order="F")))
    return sp.hstack(cols, format="csc") if cols else sp.csc_matrix((d1 * d2, 0))


def determine_search_location(A, d1, d2, min_size=3.0, max_size=8.0, dist=3.0):
    """'ellipse' method (determine_search_location.m:57-92).  Returns a (d, K) boolean csc matrix."""
    A = sp.csc_matrix(A, dtype=np.float64).copy().tolil()
    d, K = A.shape
    empty = np.asarray(A.sum(axis=0)).ravel() == 0
    for k in np.nonzero(empty)[0]:
        A[0, k] = 1.0                                                     # :51-55
    A = A.tocsc()
    x = np.tile(np.arange(1, d1 + 1), d2).astype(np.float64)              # Coor.x
    y = np.repeat(np.arange(1, d2 + 1), d1).astype(np.float64)            # Coor.y
    cols = []
    for k in range(K):
        a = np.asarray(A[:, k].todense()).ravel()
        s = a.sum()
        cm = np.array([x @ a / s, y @ a / s])                             # com.m:26
        cm[cm < 0] = 0
        cm[0] = min(cm[0], d1); cm[1] = min(cm[1], d2)
        cor = np.stack([x - cm[0], y - cm[1]], axis=1)
        Vr = (cor.T * a) @ cor / s
        D, V = np.linalg.eigh(Vr)                                         # ascending, like MATLAB eig of a symmetric matrix
        d11 = min(max_size ** 2, max(min_size ** 2, D[0]))
        d22 = min(max_size ** 2, max(min_size ** 2, D[1]))
        ind = np.sqrt((cor @ V[:, 0]) ** 2 / d11 + (cor @ V[:, 1]) ** 2 / d22) <= dist
        cols.append(sp.csc_matrix(ind.reshape(-1, 1)))
    return sp.hstack(cols, format="csc").astype(bool) if cols else sp.csc_matrix((d, 0), dtype=bool)


def threshold_components(A, d1, d2, nrgthr=0.9999, nb=1):
    """utilities/threshold_components.m:1-62 (2-D): 3x3 median, energy threshold, 3x3 closing, largest-energy 8-connected
    component; the last nb columns are copied unchanged (:24)."""
    A = sp.csc_matrix(A, dtype=np.float64)
    d, K = A.shape
    cols = []
    for i in range(K):
        a = np.asarray(A[:, i].todense()).ravel()
        if i >= K - nb:
            cols.append(sp.csc_matrix(a.reshape(-1, 1)))
            continue
        At = ndi.median_filter(a.reshape(d1, d2, order="F"), size=3, mode="constant", cval=0.0).ravel(order="F")
        e = At ** 2
        ind = np.argsort(e, kind="stable")                                  # [temp, ind] = sort(A_temp.^2, 'ascend')
        temp = np.cumsum(e[ind])
        hit = np.nonzero(temp > (1 - nrgthr) * temp[-1])[0]
        BW = np.zeros(d, bool)
        if hit.size:
            BW[ind[hit[0]:]] = True
        BW = BW.reshape(d1, d2, order="F")
        sq = np.ones((3, 3), bool)
        dil = ndi.binary_dilation(BW, structure=sq, border_value=0)
        BW = ndi.binary_erosion(dil, structure=sq, border_value=1)          # imclose: outside pixels never decide
        # label in column-major discovery order (scipy scans row-major): label the transpose
        Lt, num = ndi.label(BW.T, structure=np.ones((3, 3), bool))
        Lab = Lt.T
        out = np.zeros(d)
        if num > 0:
            nrg = np.array([np.sum(e[(Lab == l).ravel(order="F")]) for l in range(1, num + 1)])
            ff = (Lab == (int(np.argmax(nrg)) + 1)).ravel(order="F")
            out[ff] = At[ff]
        cols.append(sp.csc_matrix(out.reshape(-1, 1)))
    return sp.hstack(cols, format="csc") if cols else sp.csc_matrix((d, 0))


def determine_search_location_dilate(A, d1, d2, nrgthr=0.9999, nb=1, bSiz=3):
    """'dilate' method (determine_search_location.m:93-99) with expandCore = strel('disk', bSiz, 0)."""
    A = sp.csc_matrix(A, dtype=np.float64).copy().tolil()
    d, K = A.shape
    empty = np.asarray(A.sum(axis=0)).ravel() == 0
    for k in np.nonzero(empty)[0]:
        A[0, k] = 1.0
    Ath = threshold_components(A.tocsc(), d1, d2, nrgthr, nb)
    rr, cc = np.meshgrid(np.arange(-bSiz, bSiz + 1), np.arange(-bSiz, bSiz + 1), indexing="ij")
    disk = (rr ** 2 + cc ** 2) <= bSiz ** 2
    cols = []
    for i in range(K):
        a = np.asarray(Ath[:, i].todense()).reshape(d1, d2, order="F")
        dil = ndi.grey_dilation(a, footprint=disk, mode="constant", cval=-np.inf)
        cols.append(sp.csc_matrix((dil > 0).reshape(-1, 1, order="F")))
    return sp.hstack(cols, format="csc").astype(bool) if cols else sp.csc_matrix((d, 0), dtype=bool)
