"""Regenerates tests/golden/data_1p_c1.npz + data_1p_c1_video.npy.xz (run in the development container, where /root/reference
exists): BASELINE.json configs[0] -- the reference's own demo movie demos/data_1p.tif at its full 128 x 128 field of view, first
1000 frames ("128x128x1000, ~50 neurons" in BASELINE.json; the file holds 2000), ring radius 18 (the demos' value), one patch.

  input  : the uint16 frames as they are in the TIFF (xz-compressed .npy: 15 MB)
  state  : a fixed initial (A0, C0, IND, sn): 50 Gaussian seeds at the strongest local peak-to-noise maxima
  golden : outputs of the float64 oracle for  update_background (first run: all 16384 pixels, 121 x 121 systems)
           -> update_spatial (hals_thresh, demo_large_data_1p.m:32) -> update_temporal   with bg_ssub = 1, and the same chain with
           bg_ssub = 2 (demo_large_data_1p.m:30).  Of the ring weights (16384 x 120 doubles) only 256 sampled rows are kept.

The reference itself (MATLAB) cannot run here: these vectors pin the ORACLE and the CUDA path to each other on the reference's
real data; they are not MATLAB outputs (PARITY UNPINNED against MATLAB)."""
import lzma
import os
import sys
import time

import numpy as np
import scipy.sparse as sp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

T_C1, K_C1, RR = 1000, 50, 18


def load_video(path=None):
    """(128, 128, 1000) uint16 from the committed fixture."""
    path = path or os.path.join(HERE, "data_1p_c1_video.npy.xz")
    with lzma.open(path, "rb") as f:
        return np.load(f)


def sample_rows(d, n=256, seed=7):
    return np.sort(np.random.default_rng(seed).choice(d, n, replace=False))


def main():
    from PIL import Image
    from make_golden import build_state
    from oracle import cnmfe as OC
    from oracle.ssub import OracleSources2DSsub
    im = Image.open("/root/reference/demos/data_1p.tif")
    frames = []
    for i in range(T_C1):
        im.seek(i)
        frames.append(np.array(im))
    Y = np.stack(frames, axis=2).astype(np.uint16)
    with lzma.open(os.path.join(HERE, "data_1p_c1_video.npy.xz"), "wb", preset=6) as f:
        np.save(f, Y)
    d1, d2, T = Y.shape
    A0, C0, IND, sn = build_state(Y, K=K_C1, rr=RR)
    rows = sample_rows(d1 * d2)
    out = dict(A0_data=A0.data, A0_indices=A0.indices, A0_indptr=A0.indptr, C0=C0, sn=sn, rows=rows,
               IND_indices=sp.csc_matrix(IND).indices, IND_indptr=sp.csc_matrix(IND).indptr)
    for tag, mk in (("s1", lambda: OC.OracleSources2D(Y, (d1, d2), ring_radius=RR, options=dict(spatial_algorithm="hals_thresh"))),
                    ("s2", lambda: OracleSources2DSsub(Y, (d1, d2), ring_radius=RR, bg_ssub=2, options=dict(spatial_algorithm="hals_thresh")))):
        t0 = time.time()
        o = mk()
        o.A, o.C = A0.copy(), C0.copy()
        o.P["sn"] = sn
        o.update_background_parallel()
        W = sp.csr_matrix(o.W[(0, 0)])
        if tag == "s1":
            Wr = W[rows]
            out["W_rows_data"], out["W_rows_indices"], out["W_rows_indptr"] = Wr.data, Wr.indices, Wr.indptr
        else:
            out["W2_data"], out["W2_indices"], out["W2_indptr"] = W.data, W.indices, W.indptr       # coarse grid: 4096 x 4096, 56 per row
        out["b0_" + tag] = o.b0[(0, 0)].copy()
        o.update_spatial_parallel(IND=IND)
        A1 = sp.csc_matrix(o.A)
        out["A1_%s_data" % tag], out["A1_%s_indices" % tag], out["A1_%s_indptr" % tag] = A1.data, A1.indices, A1.indptr
        o.update_temporal_parallel()
        out["C_" + tag], out["C_raw_" + tag] = o.C, o.C_raw
        Ss = sp.csr_matrix(o.S)
        out["S_%s_data" % tag], out["S_%s_indices" % tag], out["S_%s_indptr" % tag] = Ss.data, Ss.indices, Ss.indptr
        out["g_" + tag] = np.array([p[0] for p in o.P["kernel_pars"]])
        out["neuron_sn_" + tag] = o.P["neuron_sn"]
        print(tag, "oracle chain %.1f s; spikes %d, nnz(A) %d" % (time.time() - t0, Ss.nnz, A1.nnz))
    path = os.path.join(HERE, "data_1p_c1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
