"""
oracle/gen.py -- TEST INFRASTRUCTURE ONLY.  Deterministic synthetic 1p data + initial state (SURVEY.md §8d recipe):
neurons = truncated isotropic Gaussians (sigma=gSig, radius gSiz/2), AR(1) traces with Bernoulli spikes
(recursion as OASIS_matlab/functions/gen_data.m:32-39), smooth 1p background b0 + sum_i G_i(x) f_i(t), iid noise,
quantised to uint16.  Seeds: numpy default_rng(seed).
"""
import numpy as np
import scipy.sparse as sp
from scipy.ndimage import gaussian_filter


def place_centres(rng, d1, d2, K, min_dist=8.0, margin=8):
    pts = []
    tries = 0
    while len(pts) < K and tries < 200000:
        tries += 1
        r = rng.uniform(margin, d1 - 1 - margin)
        c = rng.uniform(margin, d2 - 1 - margin)
        if all((r - a) ** 2 + (c - b) ** 2 >= min_dist ** 2 for a, b in pts[-400:]) and \
                all((r - a) ** 2 + (c - b) ** 2 >= min_dist ** 2 for a, b in pts):
            pts.append((r, c))
    return np.array(pts)


def footprints(d1, d2, centres, gSig=3.0, radius=6.5, amp=None):
    """Sparse (d, K) CSC, column-major pixel index r + c*d1."""
    rows, cols, vals = [], [], []
    R = int(np.ceil(radius))
    for k, (r0, c0) in enumerate(centres):
        ri, ci = int(round(r0)), int(round(c0))
        for dc in range(-R, R + 1):
            for dr in range(-R, R + 1):
                r, c = ri + dr, ci + dc
                if 0 <= r < d1 and 0 <= c < d2:
                    d2_ = (r - r0) ** 2 + (c - c0) ** 2
                    if d2_ <= radius ** 2:
                        rows.append(r + c * d1)
                        cols.append(k)
                        vals.append((1.0 if amp is None else amp[k]) * np.exp(-d2_ / (2 * gSig ** 2)))
    return sp.csc_matrix((vals, (rows, cols)), shape=(d1 * d2, len(centres)))


def disk_mask(d1, d2, centres, radius):
    rows, cols = [], []
    R = int(np.ceil(radius))
    for k, (r0, c0) in enumerate(centres):
        ri, ci = int(round(r0)), int(round(c0))
        for dc in range(-R, R + 1):
            for dr in range(-R, R + 1):
                r, c = ri + dr, ci + dc
                if 0 <= r < d1 and 0 <= c < d2 and (r - r0) ** 2 + (c - c0) ** 2 <= radius ** 2:
                    rows.append(r + c * d1)
                    cols.append(k)
    return sp.csc_matrix((np.ones(len(rows), dtype=bool), (rows, cols)), shape=(d1 * d2, len(centres)))


def ar_traces(rng, K, T, g=(0.95,), rate=0.5 / 30):
    S = (rng.random((K, T)) < rate).astype(np.float64)
    C = S.copy()
    g = np.atleast_1d(g)
    p = g.size
    for t in range(p, T):
        for j in range(p):
            C[:, t] += g[j] * C[:, t - 1 - j]
    return C, S


def lowpass_unit(rng, n, T, width=150.0):
    f = gaussian_filter(rng.standard_normal((n, T + 6 * int(width))), sigma=(0, width), mode="wrap")
    f = f[:, 3 * int(width):3 * int(width) + T]
    f -= f.mean(axis=1, keepdims=True)
    f /= f.std(axis=1, keepdims=True)
    return f


def make_synthetic(d1, d2, T, K, seed, sn=10.0, nblob=8, bg_amp=100.0, chunk=500, b0_level=2000.0,
                   sigma_b0=40.0, sigma_blob=60.0):
    """Returns dict with Y (d1,d2,T) uint16, A_true (csc), C_true, S_true, centres, A0 (csc), C0, IND (csc bool)."""
    rng = np.random.default_rng(seed)
    centres = place_centres(rng, d1, d2, K)
    K = len(centres)
    amp = rng.uniform(0.5, 1.5, K) * 20.0
    A = footprints(d1, d2, centres, amp=amp)
    C, S = ar_traces(rng, K, T)
    sc = min(d1, d2) / 512.0
    b0 = b0_level + gaussian_filter(rng.normal(0, 300.0, (d1, d2)), sigma_b0 * max(sc, 0.1)) * 10.0
    gy, gx = np.mgrid[0:d1, 0:d2]
    blobs = np.zeros((nblob, d1 * d2))
    for i in range(nblob):
        r0, c0 = rng.uniform(0, d1), rng.uniform(0, d2)
        blobs[i] = np.exp(-((gy - r0) ** 2 + (gx - c0) ** 2) / (2 * (sigma_blob * max(sc, 0.15)) ** 2)).ravel(order="F")
    f = lowpass_unit(rng, nblob, T) * bg_amp
    Y = np.empty((d1, d2, T), dtype=np.uint16)
    b0v = b0.ravel(order="F")
    for t0 in range(0, T, chunk):
        t1 = min(T, t0 + chunk)
        X = (A @ C[:, t0:t1]) + b0v[:, None] + blobs.T @ f[:, t0:t1] + rng.normal(0, sn, (d1 * d2, t1 - t0))
        Y[:, :, t0:t1] = np.clip(np.rint(X), 0, 65535).astype(np.uint16).reshape(d1, d2, t1 - t0, order="F")
    # initial state
    A0d = A.toarray() * (1 + 0.2 * rng.standard_normal((d1 * d2, K)))
    A0d[A0d < 0] = 0
    for k in range(K):
        A0d[:, k] = gaussian_filter(A0d[:, k].reshape(d1, d2, order="F"), 1.0).ravel(order="F")
    A0d[A0d < 1e-3 * A0d.max()] = 0
    C0 = C + rng.normal(0, 0.2, C.shape)
    IND = disk_mask(d1, d2, centres, 8.5)
    return dict(Y=Y, A_true=A, C_true=C, S_true=S, centres=centres, A0=sp.csc_matrix(A0d), C0=C0, IND=IND, sn=sn)
