"""Regenerates tests/golden/data_1p_crop.npz (run in the development container, where /root/reference exists):

  input  : a 40 x 40 x 600 crop of the reference's own demo movie demos/data_1p.tif (uint16, 128x128x2000)
  state  : a fixed, seeded initial (A0, C0, IND, sn) built from the crop (5 Gaussian seeds at the brightest
           fluctuating spots, ring radius 9)
  golden : outputs of the float64 oracle (oracle/cnmfe.py) for background -> spatial(hals_thresh) -> temporal

The reference itself (MATLAB) cannot run here, so these vectors pin the ORACLE and the CUDA path to each other on real
CNMF-E data; they are not MATLAB outputs (PARITY UNPINNED against MATLAB)."""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def build_state(Y, K=5, rr=9):
    from scipy.ndimage import gaussian_filter, maximum_filter
    from oracle import gen, oasis as O
    d1, d2, T = Y.shape
    Yf = Y.astype(np.float64)
    hp = Yf - gaussian_filter(Yf, (4, 4, 0))
    pnr = hp.max(axis=2) / (hp.std(axis=2) + 1e-9)
    loc = (pnr == maximum_filter(pnr, size=9))
    loc[:6, :] = loc[-6:, :] = False
    loc[:, :6] = loc[:, -6:] = False
    cand = np.argwhere(loc)
    cand = cand[np.argsort(-pnr[loc])][:K]
    centres = cand.astype(np.float64)
    A0 = gen.footprints(d1, d2, centres, amp=np.ones(len(centres)))
    Yr = Yf.reshape(-1, T, order="F")
    C0 = np.zeros((len(centres), T))
    for k in range(len(centres)):
        a = A0[:, k].toarray().ravel()
        tr = a @ (hp.reshape(-1, T, order="F")) / (a @ a)
        C0[k] = tr - np.median(tr)
    IND = gen.disk_mask(d1, d2, centres, 7.5)
    sn = O.GetSn(Yr).reshape(d1, d2, order="F")
    return A0, C0, IND, sn


def main():
    from PIL import Image
    from oracle import cnmfe as OC
    im = Image.open("/root/reference/demos/data_1p.tif")
    frames = []
    for i in range(600):
        im.seek(i)
        frames.append(np.array(im)[44:84, 60:100])
    Y = np.stack(frames, axis=2).astype(np.uint16)
    A0, C0, IND, sn = build_state(Y)
    o = OC.OracleSources2D(Y, (40, 40), ring_radius=9, options=dict(spatial_algorithm="hals_thresh"))
    o.A, o.C = A0.copy(), C0.copy()
    o.P["sn"] = sn
    o.update_background_parallel()
    W = sp.csr_matrix(o.W[(0, 0)])
    b0 = o.b0[(0, 0)].copy()
    o.update_spatial_parallel(IND=IND)
    A1 = o.A.toarray()
    o.update_temporal_parallel()
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data_1p_crop.npz")
    np.savez_compressed(out, Y=Y, A0=A0.toarray(), C0=C0, IND=IND.toarray(), sn=sn, W_data=W.data, W_indices=W.indices,
                        W_indptr=W.indptr, b0=b0, A1=A1, C=o.C, C_raw=o.C_raw, S=o.S,
                        kernel_pars=np.array([p[0] for p in o.P["kernel_pars"]]), neuron_sn=o.P["neuron_sn"])
    print("wrote", out, os.path.getsize(out), "bytes; spikes per neuron:", (o.S > 0).sum(1))


if __name__ == "__main__":
    main()
