#!/usr/bin/env python
"""bench.py -- one "step" = one CNMF-E update iteration (update_background_parallel + update_spatial_parallel +
update_temporal_parallel, the triple at demos/demo_large_data_1p.m:199-201) over a resident synthetic 1p video.

Workload at N=1: BASELINE.json configs[1]: 512x512x10000 uint16, 300 neurons, ring background (radius 18, 120
neighbours, bg_ssub=1), single patch, nnls spatial, foopsi/ar1 deconvolution with the demo's options.
N>1 (weak scaling): the FOV grows to 512 x 512N, one 512x512 patch (+19 px halo) and 300 neurons per GPU;
one all-reduce of the K x T merge buffers per step (update_temporal_parallel.m:269-280) + the A-row exchange.

Prints ONE JSON line (rank 0).  `--impl reference` times the float64 CPU restatement of the reference algorithm
(oracle/, NumPy/SciPy, all host cores) on a bounded sample -- MATLAB is not installed on these boxes.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20260925 + 2
D1 = 512
D2_PER_GPU = 512
T_FULL = 10000
K_PER_GPU = 300
RING = 18


# ----------------------------------------------------------------------------------------------- synthetic data
def _place_centres(rng, d1, d2, K, min_dist=8.0, margin=8):
    pts = np.zeros((0, 2))
    while len(pts) < K:
        cand = np.stack([rng.uniform(margin, d1 - 1 - margin, 4 * K), rng.uniform(margin, d2 - 1 - margin, 4 * K)], 1)
        for p in cand:
            if len(pts) >= K:
                break
            if len(pts) == 0 or np.min(np.sum((pts - p) ** 2, axis=1)) >= min_dist ** 2:
                pts = np.vstack([pts, p])
    return pts


def _footprints(d1, d2, centres, amp, gSig=3.0, radius=6.5):
    import scipy.sparse as sp
    rows, cols, vals = [], [], []
    R = int(np.ceil(radius))
    dr, dc = np.meshgrid(np.arange(-R, R + 1), np.arange(-R, R + 1), indexing="ij")
    for k, (r0, c0) in enumerate(centres):
        r = int(round(r0)) + dr.ravel()
        c = int(round(c0)) + dc.ravel()
        d2_ = (r - r0) ** 2 + (c - c0) ** 2
        ok = (r >= 0) & (r < d1) & (c >= 0) & (c < d2) & (d2_ <= radius ** 2)
        rows.append(r[ok] + c[ok] * d1)
        cols.append(np.full(ok.sum(), k))
        vals.append(amp[k] * np.exp(-d2_[ok] / (2 * gSig ** 2)))
    return sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(d1 * d2, len(centres)))


def _disk_mask(d1, d2, centres, radius):
    import scipy.sparse as sp
    rows, cols = [], []
    R = int(np.ceil(radius))
    dr, dc = np.meshgrid(np.arange(-R, R + 1), np.arange(-R, R + 1), indexing="ij")
    for k, (r0, c0) in enumerate(centres):
        r = int(round(r0)) + dr.ravel()
        c = int(round(c0)) + dc.ravel()
        ok = (r >= 0) & (r < d1) & (c >= 0) & (c < d2) & ((r - r0) ** 2 + (c - c0) ** 2 <= radius ** 2)
        rows.append(r[ok] + c[ok] * d1)
        cols.append(np.full(ok.sum(), k))
    return sp.csc_matrix((np.ones(sum(len(x) for x in rows), dtype=bool), (np.concatenate(rows), np.concatenate(cols))),
                         shape=(d1 * d2, len(centres)))


def make_problem(d1, d2, T, K, seed, kind="1p"):
    """Global (all-rank) description: neurons, traces, background parameters.  SURVEY.md §8d recipe.
    kind "2p" (configs[3]): rank-1 background b (x) f, b = 100 (1 + 0.5 blur_30(N(0,1)) / max|.|), f = 1 + 0.1 lowpass, noise 1."""
    from scipy.ndimage import gaussian_filter
    rng = np.random.default_rng(seed)
    centres = _place_centres(rng, d1, d2, K)
    amp = rng.uniform(0.5, 1.5, K) * 20.0
    A = _footprints(d1, d2, centres, amp)
    S = (rng.random((K, T)) < 0.5 / 30).astype(np.float64)
    C = S.copy()
    for t in range(1, T):
        C[:, t] += 0.95 * C[:, t - 1]
    b0 = 2000.0 + gaussian_filter(rng.normal(0, 300.0, (d1, d2)), 40.0) * 10.0
    nblob = 8 * max(1, d2 // 512)
    bc = np.stack([rng.uniform(0, d1, nblob), rng.uniform(0, d2, nblob)], 1)
    width = 150.0
    f = gaussian_filter(rng.standard_normal((nblob, T + 6 * int(width))), sigma=(0, width), mode="wrap")
    f = f[:, 3 * int(width):3 * int(width) + T]
    f -= f.mean(axis=1, keepdims=True)
    f /= f.std(axis=1, keepdims=True)
    f *= 100.0
    A0 = A.copy()
    A0.data = np.maximum(A0.data * (1 + 0.2 * rng.standard_normal(A0.data.size)), 0.0)
    C0 = C + rng.normal(0, 0.2, C.shape)
    IND = _disk_mask(d1, d2, centres, 8.5)
    out = dict(A=A, C=C, S=S, b0=b0, blob_centres=bc, f=f, A0=A0, C0=C0, IND=IND, centres=centres, sn=10.0)
    if kind == "2p":
        fld = gaussian_filter(rng.standard_normal((d1, d2)), 30.0)
        out["bgb"] = 100.0 * (1.0 + 0.5 * fld / np.abs(fld).max())
        lp = gaussian_filter(rng.standard_normal(T + 900), 150.0)[450:450 + T]
        out["bgf"] = 1.0 + 0.1 * lp / np.abs(lp).max()
        out["b0"] = np.zeros((d1, d2)); out["blob_centres"] = bc[:0]; out["f"] = f[:0]; out["sn"] = 1.0
    return out


def build_block_on_gpu(prob, block, d1, T, seed, device, sn=10.0, chunk=500):
    """uint16 frame-major block [T][nrb*ncb] on the GPU (torch), deterministic per (seed, FOV column)."""
    import torch
    r0, r1, c0, c1 = [int(x) for x in block]       # 1-based inclusive
    nrb, ncb = r1 - r0 + 1, c1 - c0 + 1
    dev = torch.device("cuda", device)
    rows = torch.arange(r0 - 1, r1, device=dev, dtype=torch.float32)
    out = torch.empty((T, ncb, nrb), dtype=torch.int16, device=dev)   # uint16 bit patterns
    A = prob["A"].tocsr()
    pix = (np.arange(c0 - 1, c1)[:, None] * d1 + np.arange(r0 - 1, r1)[None, :]).ravel()   # block order: c major, r fastest
    Ab = A[pix, :]
    keep = np.nonzero(np.asarray(Ab.sum(axis=0)).ravel() > 0)[0]
    Abk = Ab[:, keep].tocoo()
    A_t = torch.sparse_coo_tensor(np.vstack([Abk.row, Abk.col]), Abk.data.astype(np.float32),
                                  size=(nrb * ncb, len(keep)), device=dev).coalesce()
    C_t = torch.from_numpy(prob["C"][keep].astype(np.float32)).to(dev)
    f_t = torch.from_numpy(prob["f"].astype(np.float32)).to(dev)
    b0_t = torch.from_numpy(prob["b0"][r0 - 1:r1, c0 - 1:c1].T.astype(np.float32).copy()).to(dev)   # (ncb, nrb)
    cols = torch.arange(c0 - 1, c1, device=dev, dtype=torch.float32)
    blobs = []
    for (br, bcc) in prob["blob_centres"]:
        g = torch.exp(-((rows[None, :] - br) ** 2 + (cols[:, None] - bcc) ** 2) / (2 * 60.0 ** 2))
        blobs.append(g.reshape(-1))
    blobs = torch.stack(blobs, 1) if blobs else None          # (ncb*nrb, nblob)
    bgb = bgf = None
    if "bgb" in prob:
        bgb = torch.from_numpy(prob["bgb"][r0 - 1:r1, c0 - 1:c1].T.astype(np.float32).copy()).to(dev).reshape(-1, 1)
        bgf = torch.from_numpy(prob["bgf"].astype(np.float32)).to(dev)
    sn = float(prob.get("sn", sn))
    gen = torch.Generator(device=dev)
    for t0 in range(0, T, chunk):
        t1 = min(T, t0 + chunk)
        X = torch.sparse.mm(A_t, C_t[:, t0:t1]) + b0_t.reshape(-1, 1)
        if blobs is not None:
            X = X + blobs @ f_t[:, t0:t1]
        if bgb is not None:
            X = X + bgb * bgf[None, t0:t1]
        X = X.reshape(ncb, nrb, t1 - t0)
        for j in range(ncb):
            gen.manual_seed(seed * 1000003 + (c0 - 1 + j) * 4099 + t0)
            X[j] += torch.randn((nrb, t1 - t0), generator=gen, device=dev) * sn
        X = torch.clamp(torch.round(X), 0, 65535)
        out[t0:t1] = X.permute(2, 0, 1).to(torch.int32).to(torch.int16)   # wraps: same bits as uint16
    return out.reshape(T, ncb * nrb)


# ----------------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region.  NVML in-process (a sample every 20 ms: the default timed region
    is about a second, shorter than an `nvidia-smi` start-up); `nvidia-smi -lms` is the fallback, started ahead of the region
    (`prepare`) with only the lines that arrive inside it kept."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.index, self.samples, self.proc, self.h, self.nv = index, [], None, None, None
        self.on = False
        self.alive = True
        self.source = None

    def prepare(self):
        """Before the warm-up: open NVML (or launch nvidia-smi) so that sampling costs nothing when the region starts."""
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = None
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
                h = nv.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                h = nv.nvmlDeviceGetHandleByIndex(self.index)
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            self.nv, self.h, self.source = nv, h, "nvml"
            threading.Thread(target=self._poll, daemon=True).start()
            return
        except Exception:
            self.nv = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nv
        bits = (("hw_slowdown", getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8)),
                ("hw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40)),
                ("sw_thermal_slowdown", getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20)),
                ("sw_power_cap", getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)))
        try:
            mx = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
        except Exception:
            mx = float("nan")
        while self.alive:
            if self.on:
                try:
                    sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                    try:
                        r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                    except Exception:
                        r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    self.samples.append((sm, mx, [n for n, b in bits if r & b]))
                except Exception:
                    pass
            time.sleep(0.02)

    def _read(self):
        for line in self.proc.stdout:
            if not self.on:
                continue
            p = [x.strip() for x in line.strip().split(",")]
            if len(p) < 7:
                continue
            try:
                sm, mx = float(p[0]), float(p[1])
            except ValueError:
                continue
            names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
            self.samples.append((sm, mx, [n for n, v in zip(names, p[3:7]) if v.lower().startswith("active")]))

    def start(self):
        if self.source is None:
            self.prepare()
        self.on = True

    def stop(self):
        self.on = False
        self.alive = False
        if self.proc:
            self.proc.terminate()
        sm = [s[0] for s in self.samples]
        mx = [s[1] for s in self.samples if s[1] == s[1]]
        reasons = set()
        for s in self.samples:
            reasons.update(s[2])
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm), source=self.source)


# ----------------------------------------------------------------------------------------------- CPU reference arm
SAMPLE = dict(d1=96, d2=96, T=1500, K=12, ring=RING)   # ~5-12 s of CPU work per iteration on 8-16 host cores


def cpu_reference_step(state=None):
    """One update iteration of the float64 oracle on the bounded sample.  Returns (seconds, pixel*frames)."""
    from oracle import gen, cnmfe as OC
    if state is None:
        D = gen.make_synthetic(SAMPLE["d1"], SAMPLE["d2"], SAMPLE["T"], SAMPLE["K"], seed=SEED, nblob=4)
        o = OC.OracleSources2D(D["Y"], (SAMPLE["d1"], SAMPLE["d2"]), ring_radius=SAMPLE["ring"],
                               options=dict(spatial_algorithm="nnls"))
        o.A, o.C = D["A0"].copy(), D["C0"].copy()
        o.P["sn"] = np.full((SAMPLE["d1"], SAMPLE["d2"]), 10.0)
        state = dict(o=o, IND=D["IND"])
    o = state["o"]
    t = time.perf_counter()
    o.update_background_parallel()
    o.update_spatial_parallel(IND=state["IND"])
    o.update_temporal_parallel()
    dt = time.perf_counter() - t
    return dt, SAMPLE["d1"] * SAMPLE["d2"] * SAMPLE["T"], state


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count()
    state = None
    for _ in range(args.warmup):
        _, _, state = cpu_reference_step(state)
    t0 = time.perf_counter()
    units = 0
    for _ in range(args.steps):
        dt, u, state = cpu_reference_step(state)
        units += u
    el = time.perf_counter() - t0
    val = units / el
    sample = "%dx%dx%d uint16, K=%d, ring r=%d (120 nbrs), single patch, nnls + foopsi/ar1; per-pixel-frame cost is size-independent" % (
        SAMPLE["d1"], SAMPLE["d2"], SAMPLE["T"], SAMPLE["K"], SAMPLE["ring"])
    line = dict(impl="reference", metric="pixels*frames/sec per spatial+temporal+BG update iter", value=val,
                unit="pixel*frames/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * el / max(args.steps, 1), higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64", data="synthetic",
                config=dict(workload="CNMF-E update iteration (BG ring fit + nnls spatial + HALS temporal/OASIS), bounded sample of configs[1]",
                            sample=sample),
                cpu_baseline=dict(value=val, unit="pixel*frames/s", cores=cores, kind="port", sample=sample,
                                  note="float64 NumPy/SciPy restatement of the reference algorithm (oracle/); MATLAB is not installed"),
                e2e=dict(value=val, unit="pixel*frames/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------- full-size checks
def full_size_invariants(A, IND, C, S, kernel_pars, neuron_sn, smin_mult=5.0):
    """Size-independent properties of one update iteration, evaluated on the FULL-size host results after the timed
    region (never inside it; nothing here runs on the GPU):
      spatial  A >= 0 and supp(A) inside the search mask (update_spatial_parallel.m:321-335, nnls);
      temporal oasisAR1 pool algebra of the final deconvTemporal (oasisAR1.m:101-109): S >= 0, S_t = C_t - g C_{t-1} where
               S_t > 0, C_t = g C_{t-1} elsewhere (t >= 1); and the smin rule in the form the reference enforces it: the
               forward test oasisAR1.m:64-65 compares v'/w' with (v/w) g^l + smin WITHOUT clipping v/w at 0 (only the
               back-track test :82-83 has max(0, .)), so a pool that follows a NEGATIVE pool (c clipped to 0, :105) may
               start with a spike below smin.  Exact invariant: S_t >= smin wherever S_t > 0 and C_{t-1} > 0.
    Returns max violations (relative to the trace scale) -- reported, not asserted."""
    import scipy.sparse as sp
    out = {}
    try:
        A = sp.csc_matrix(A)
        out["A_min"] = float(A.data.min()) if A.nnz else 0.0
        outside = A.copy(); outside.data[:] = 1.0
        ind = sp.csc_matrix(IND).astype(np.float64)
        out["A_nnz_outside_search_mask"] = int((outside - outside.multiply(ind)).count_nonzero())
        g = np.asarray(kernel_pars, dtype=np.float64).reshape(len(neuron_sn), -1)[:, 0]
        C, S = np.asarray(C), np.asarray(S)
        scale = np.maximum(np.abs(C).max(axis=1), 1e-300)
        pred = g[:, None] * C[:, :-1]
        resid = C[:, 1:] - pred - S[:, 1:]
        out["S_min"] = float(S.min()) if S.size else 0.0
        out["max_rel_AR1_residual"] = float((np.abs(resid) / scale[:, None]).max()) if resid.size else 0.0
        spikes = S[:, 1:] > 0
        smin = smin_mult * np.asarray(neuron_sn, dtype=np.float64)
        after_positive = spikes & (C[:, :-1] > 0)
        viol = np.where(after_positive, smin[:, None] - S[:, 1:], 0.0)
        out["max_rel_spike_below_smin_after_positive_pool"] = float(max(0.0, (viol / scale[:, None]).max())) if viol.size else 0.0
        out["n_spikes"] = int(spikes.sum())
        out["n_spikes_below_smin_after_clipped_pool"] = int((spikes & ~after_positive & (S[:, 1:] < smin[:, None])).sum())
        out["ok"] = bool(out["A_min"] >= 0.0 and out["A_nnz_outside_search_mask"] == 0 and out["S_min"] >= 0.0
                         and out["max_rel_AR1_residual"] < 1e-9 and out["max_rel_spike_below_smin_after_positive_pool"] < 1e-9)
    except Exception as e:   # a reporting aid must never take the bench line down
        out["error"] = repr(e)
    return out


def full_size_oracle_checks(obj, IND, n_pix=200, seed=1):
    """One more update iteration through the host-buffer API after the timed region, each stage spot-checked against the float64
    oracle re-deriving individual output rows from the raw video (oracle/spot.py): W rows (fit_ring_model.m:92-108), A rows
    (update_spatial_parallel.m:157-166 + nnls_spatial.m) and ALL traces of the final deconvTemporal (deconvTemporal.m:62-84) on the
    merged C_raw the CUDA path fed it.  Single rank, single patch form (configs[1])."""
    import ctypes
    import scipy.sparse as sp
    from cnmf_e_b200 import _lib
    from oracle import spot
    out = {}
    try:
        lib = _lib.lib()
        i = obj.owned_patches()[0]
        T = obj.T

        def rows_fn(idx):
            idx = np.ascontiguousarray(idx, dtype=np.int32)
            buf = np.empty((idx.size, T), dtype=np.uint16)
            _lib.check(lib.cnmfe_debug_video_rows(obj._h, i, idx.size, idx.ctypes.data_as(ctypes.c_void_p), buf.ctypes.data_as(ctypes.c_void_p)))
            return buf.astype(np.float64)

        view = spot.PatchView(obj.d1, obj.d2, obj.patch_of(i), obj.block_of(i), obj.r_shift, obj.c_shift, rows_fn)
        rng = np.random.default_rng(seed)
        W_old = obj.W[i]                       # stays intact: pull_ring alternates between two buffers
        A_bg, C_bg = obj.A, obj.C
        pmax = int((np.asarray(W_old) > 0).sum(axis=1).max())
        obj.update_background_parallel()
        W_new, b0 = obj.W[i], obj.b0[i]
        # sampled pixels: half from the refitted (active) set, half from the search masks
        sumA = np.asarray(abs(sp.csr_matrix(A_bg)).sum(axis=1)).ravel()
        changed = np.nonzero((np.asarray(W_new) != np.asarray(W_old)).any(axis=1))[0]
        INDr = sp.csr_matrix(IND)
        p = obj.patch_of(i)
        rr_, cc_ = np.meshgrid(np.arange(p[0] - 1, p[1]), np.arange(p[2] - 1, p[3]), indexing="ij")
        fov_of_patch_pixel = (rr_ + cc_ * obj.d1).ravel(order="F")
        inmask = np.nonzero(np.asarray(INDr[fov_of_patch_pixel].sum(axis=1)).ravel() > 0)[0]
        pix_a = rng.choice(changed, min(n_pix // 2, changed.size), replace=False) if changed.size else np.zeros(0, dtype=int)
        pix_m = rng.choice(inmask, min(n_pix // 2, inmask.size), replace=False) if inmask.size else np.zeros(0, dtype=int)
        pixels = np.concatenate([pix_a, pix_m]).astype(int)
        out["ring_rows_vs_oracle"] = dict(spot.ring_rows(view, pixels, A_bg, C_bg, W_old, W_new, b0, pmax,
                                                         bool(obj.options["bg_acceleration"])),
                                          n_pixels_refit_by_gpu=int(changed.size), sumA_nonzero_pixels=int((sumA > 0).sum()))
        obj.options["spatial_algorithm"] = "nnls"
        obj.update_spatial_parallel(IND=IND)
        single = obj.npatch == 1
        out["spatial_rows_vs_oracle"] = spot.spatial_rows(view, pix_m, None if single else obj.A_prev, None if single else obj.C_prev,
                                                          W_new, b0, C_bg, IND, obj.A) if single else dict(skipped="multi-patch")
        obj.update_temporal_parallel()
        K = obj.A.shape[1]
        Craw_in = np.empty((K, T))
        _lib.check(lib.cnmfe_get_merged_craw(obj._h, Craw_in.ctypes.data_as(ctypes.c_void_p)))
        out["deconvTemporal_vs_oracle"] = spot.deconv_traces(Craw_in, obj.C, obj.S, obj.C_raw, obj.P["kernel_pars"],
                                                            obj.options["deconv_options"])
        out["ok"] = bool(all(v.get("ok", False) for v in out.values() if isinstance(v, dict) and "skipped" not in v))
    except Exception as e:
        import traceback
        out["error"] = repr(e) + " | " + traceback.format_exc().splitlines()[-1]
        out["ok"] = False
    return out


# ----------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from cnmf_e_b200 import _lib
    from cnmf_e_b200.sources2d import Sources2D
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    lib = _lib.lib()
    if args.workload == "c3":
        # BASELINE.json configs[2]: 1024 x 1024 x 20000, 1000 neurons, patch_dims [256, 256] -> 4 x 4 patches
        # (distribute_data.m:56-67: patch_idx = ceil(linspace(1, 1024, 5))), blocks with 19-px halos, sharded over the ranks
        d1, d2, T, K, patch_dims, seed, scaling = 1024, 1024, args.frames or 20000, 1000, (256, 256), 20260925 + 3, "strong"
        if 16 % world:
            raise SystemExit("--workload c3 shards 16 patches: --gpus must divide 16")
    elif args.workload == "c4":
        # BASELINE.json configs[3]: synthetic 2p 512 x 512 x 50000, 200 neurons, rank-1 nmf background (demo_large_data_2p path), one patch
        d1, d2, T, K, patch_dims, seed, scaling = 512, 512, args.frames or 50000, 200, (512, 512), 20260925 + 4, "weak"
        if world != 1:
            raise SystemExit("--workload c4 is a single-patch, single-GPU configuration")
    else:
        d1, d2, T, K, patch_dims, seed, scaling = D1, D2_PER_GPU * world, args.frames or T_FULL, K_PER_GPU * world, (D1, D2_PER_GPU), SEED, "weak"
    c4 = args.workload == "c4"
    prob = make_problem(d1, d2, T, K, seed, kind="2p" if c4 else "1p")
    opts = dict(spatial_algorithm="nnls", use_tensor_gram=bool(args.tensor), bg_ssub=args.bg_ssub)
    if c4:
        opts.update(background_model="nmf", nb=1, bg_ssub=1)
    obj = Sources2D(d1, d2, T, patch_dims, ring_radius=RING, device=local, rank=rank, world_size=world,
                    options=opts)
    for i in obj.owned_patches():
        blk = build_block_on_gpu(prob, obj.block_of(i), d1, T, seed, local)
        torch.cuda.synchronize()
        obj.load_block_dev(i, blk.data_ptr(), 1)
        del blk
        torch.cuda.empty_cache()
    obj.A, obj.C = prob["A0"].copy(), prob["C0"].copy()
    obj.P["sn"] = np.full((d1, d2), float(prob["sn"]))
    IND = prob["IND"]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        obj.update_background_parallel(sync_host=False)
        obj.update_spatial_parallel(IND=IND, sync_host=False)
        if world > 1:
            obj.exchange_spatial()
        obj.update_temporal_parallel(sync_host=False)

    def step_e2e():
        obj.update_background_parallel()
        obj.update_spatial_parallel(IND=IND)
        obj.update_temporal_parallel()

    # state to the device once; resident steps follow
    obj._push_options(); obj.push_neurons(); obj.push_prev(); obj.push_ring()
    import ctypes
    lib.cnmfe_set_sn(obj._h, np.asfortranarray(obj.P["sn"]).ctypes.data_as(ctypes.c_void_p))
    phases = np.zeros(7)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.prepare()
    for _ in range(args.warmup):
        step_resident()
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = lib.cnmfe_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    lib.cnmfe_timer_begin(obj._h)
    t0 = time.perf_counter()
    gram_ms = 0.0
    call_ms = np.zeros(3)
    for _ in range(args.steps):
        tc = time.perf_counter()
        obj.update_background_parallel(sync_host=False)
        call_ms[0] += 1e3 * (time.perf_counter() - tc)
        p = np.array(obj.phase_ms()); phases += p; gram_ms += p[0]
        tc = time.perf_counter()
        obj.update_spatial_parallel(IND=IND, sync_host=False)
        call_ms[1] += 1e3 * (time.perf_counter() - tc)
        phases += np.array(obj.phase_ms())
        if world > 1:
            obj.exchange_spatial()
        tc = time.perf_counter()
        obj.update_temporal_parallel(sync_host=False)
        call_ms[2] += 1e3 * (time.perf_counter() - tc)
        phases += np.array(obj.phase_ms())
    ms = ctypes.c_float()
    lib.cnmfe_timer_end(obj._h, ctypes.byref(ms))
    barrier()
    wall = time.perf_counter() - t0
    launches = lib.cnmfe_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    t_dev = max(ms.value / 1e3, 1e-9)
    t_use = max(t_dev, wall)   # host planning between launches is part of the step
    tt = torch.tensor([t_use], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    t_max = float(tt.item())
    units = float(d1) * d2 * T * args.steps
    value = units / t_max
    # ---- e2e through the public API with host buffers (H2D of A, C, W, ...; D2H of W, A, C, C_raw, S)
    obj.pull_ring(); obj.pull_spatial(); obj.pull_temporal()
    nE = max(1, min(args.steps, 2))
    # one untimed pass through the host-buffer API: first-use costs (page-locking the pull buffers) are not part of a step
    obj.update_background_parallel(); obj.update_spatial_parallel(IND=IND); obj.update_temporal_parallel()
    if os.environ.get("CNMFE_HOST_PROFILE"):      # diagnostics: where the host-buffer path spends its wall time
        import functools
        acc = {}
        def timed(name, fn):
            @functools.wraps(fn)
            def w(*a, **k):
                t = time.perf_counter()
                r = fn(*a, **k)
                acc[name] = acc.get(name, 0.0) + 1e3 * (time.perf_counter() - t)
                return r
            return w
        for nm in ("push_neurons", "push_prev", "push_ring", "pull_ring", "pull_temporal", "reconstruct_b0", "_push_options",
                   "update_background_parallel", "update_spatial_parallel", "update_temporal_parallel"):
            setattr(obj, nm, timed(nm, getattr(obj, nm)))
        def step_e2e():
            obj.update_background_parallel()
            obj.update_spatial_parallel(IND=IND)
            obj.update_temporal_parallel()
        import atexit
        atexit.register(lambda: sys.stderr.write("[e2e host profile, ms over %d steps] %s\n" % (nE, json.dumps(acc))))
    barrier()
    h2d0, d2h0 = obj.h2d_bytes, obj.d2h_bytes
    te = time.perf_counter()
    for _ in range(nE):
        step_e2e()
    barrier()
    e2e_t = (time.perf_counter() - te) / nE
    # bytes the Sources2D mirror moved inside the e2e region (it re-sends only host state that changed since the last sync);
    # taken here, before the post-timing checks read the lazily mirrored W / C_raw / S
    h2d = (obj.h2d_bytes - h2d0) // nE + d1 * d2 * 8
    d2h = (obj.d2h_bytes - d2h0) // nE
    te_t = torch.tensor([e2e_t], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te_t, op=dist.ReduceOp.MAX)
    e2e_val = float(d1) * d2 * T / float(te_t.item())
    # per-rank wall / phase times (multi-GPU: where the step waits -- the merge collective runs at the pace of the slowest rank)
    per_rank = None
    if world > 1:
        mine = dict(rank=rank, call_wall_ms=[float(x) / args.steps for x in call_ms], phase_ms=[float(x) / args.steps for x in phases],
                    device_ms=1e3 * t_dev / args.steps, wall_ms=1e3 * wall / args.steps)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        per_rank = gathered
    checks = None
    if rank == 0:
        checks = full_size_invariants(obj.A, IND, obj.C, obj.S, obj.P.get("kernel_pars"), obj.P.get("neuron_sn"))
        if world == 1 and not args.no_oracle_checks and not c4:
            checks["oracle"] = full_size_oracle_checks(obj, IND)
            checks["ok"] = bool(checks.get("ok", False) and checks["oracle"].get("ok", False))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the two big kernels of the background update (live CUDA-event times on the launching streams)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    bf16 = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "2x measured bf16 sustained (MEASURED_PEAKS.json)" if peaks else "2x fallback 1.4 PF bf16 sustained"
    ss = max(1, args.bg_ssub)
    rr = -(-RING // ss)
    ND = 2 * rr * (4 * rr + 1) + (2 * rr + 1)
    nblk = [obj.block_of(i) for i in obj.owned_patches()]
    db = float(sum(-(-int(b[1] - b[0] + 1) // ss) * -(-int(b[3] - b[2] + 1) // ss) for b in nblk))
    T_fit = int(lib.cnmfe_last_gram_frames(obj._h)) or T     # frames the ring fit used (fit_ring_model.m:84-90 may keep every k-th)
    int8_ops = 2.0 * 4.0 * ND * db * T_fit       # u16 x u16 = 4 u8 x u8 products, 2 ops per MAC
    gram_s = max(gram_ms / 1e3 / args.steps, 1e-9)
    achieved = int8_ops / gram_s / 1e12
    gram_roof = dict(bound="tensor", kernel="ring_s2_tc_kernel: banded second moments S2 (%s)" % ("tcgen05 INT8" if lib.cnmfe_last_gram_was_tensor(obj._h) else "SIMT u64"),
                     achieved=achieved, peak=2.0 * bf16, unit="TOP/s (int8 dense)", frac=achieved / (2.0 * bf16),
                     traffic=(89.47e9 * (T_fit / 10000.0) * (db / 262144.0) if (lib.cnmfe_last_gram_was_tensor(obj._h) and ss == 1) else None),
                     traffic_source="dram__bytes_read.sum (84.24 GB) + dram__bytes_write.sum (5.23 GB) of one ncu --set full capture (profiles/r1_ncu_full_gram.csv), scaled by frames/10000 and pixels/262144",
                     peak_source=peak_src, ms_per_launch=1e3 * gram_s, algorithmic_ops_per_launch=int8_ops)
    # ---- the other kernels of the step, each against the bound that applies to it (live per-phase CUDA-event times)
    ph = [float(x) / args.steps / 1e3 for x in phases]          # seconds per step
    nnb1 = 121 if ss == 1 else None
    others = {}
    solve_roof = None
    if nnb1 and ph[1] > 0:
        n_act = float(lib.cnmfe_last_active_pixels(obj._h)) or float(sum(int(b[1] - b[0] + 1) * int(b[3] - b[2] + 1) for b in nblk))
        n1 = nnb1 + 1                                              # ring pixels + ones row + right-hand side row of the augmented system
        sol_bytes = n_act * (nnb1 * (nnb1 + 1) / 2 + nnb1) * 8.0   # moment gather + weights out
        sol_flop = n_act * (2.0 * n1 ** 3 / 6.0 + 2.0 * n1 ** 2)   # LDL' of the augmented matrix + back substitution (fp64; mul and add counted)
        fp64_peak = 36.5                                           # TFLOP/s: DMMA m8n8k4 == DFMA rate measured on this pool by scripts/micro/dmma.cu
        solve_roof = dict(bound="tensor", kernel="ring_solve_mma_kernel: per-pixel 122 x 122 block LDL' (fp64 tensor path, DMMA.8x8x4)",
                          achieved=sol_flop / ph[1] / 1e12, peak=fp64_peak, unit="TFLOP/s (fp64)", frac=sol_flop / ph[1] / 1e12 / fp64_peak,
                          traffic=11.07e9 * (n_act / 235617.0),
                          traffic_source="dram__bytes_read.sum (10.80 GB) + dram__bytes_write.sum (0.27 GB) of one ncu --set full capture (profiles/r2_ncu_full_solve_mma.csv), scaled by active pixels/235617",
                          peak_source="fp64 tensor-path rate measured by scripts/micro/dmma.cu on this pool's B200 (36.5 TFLOP/s for DMMA and for DFMA); MEASURED_PEAKS.json carries no fp64 figure",
                          ms_per_launch=1e3 * ph[1], algorithmic_ops_per_launch=sol_flop, algorithmic_bytes_per_launch=sol_bytes,
                          active_pixels=n_act,
                          note="latency bound, not pipe bound: the 16 block pivots (8 x 8 inversions, a chain of dependent reciprocals) serialise each pixel; ncu: DMMA sub-pipe 20.9 % active, barrier stalls 37 % (profiles/r2_ncu_source_solve_mma_top.txt)")
    if ph[2] > 0:
        pb = 3.0 * d1 * d2 * T * 2 / world
        others["projections (proj_mc_tile x2, proj_bt_list)"] = dict(ms=1e3 * ph[2], bound="hbm", algorithmic_bytes=pb, gbs=pb / ph[2] / 1e9,
                                                                    frac=pb / ph[2] / 1e9 / peaks.get("hbm_gbs", 6650.0),
                                                                    note="three streaming passes over the resident uint16 video (SURVEY 8d: 3*d*T*2 B)")
    if ph[4] > 0:
        others["hals_temporal_kernel"] = dict(ms=1e3 * ph[4], bound="dependency chain (5 sweeps x overlapping-neuron chain of exact sequential OASIS fits)",
                                              note="CNMFE_HALS_PROFILE=1 prints the critical chain; see profiles/README_r2.md")
    if c4:
        iters = int(lib.cnmfe_last_nmf_iterations(obj._h))
        nb_bytes = (2.0 * iters + 1.0) * d1 * d2 * T * 2.0         # two streaming passes per ALS iteration + the ||B||^2 pass, u16 video
        gram_roof = dict(bound="hbm", kernel="nmf background fit: matrix-free ALS on the resident video (proj_mc_tile + svd_colsum_partial per iteration)",
                         achieved=nb_bytes / gram_s / 1e9, peak=peaks.get("hbm_gbs", 6650.0), unit="GB/s", frac=nb_bytes / gram_s / 1e9 / peaks.get("hbm_gbs", 6650.0),
                         traffic=None, peak_source="hbm_gbs of MEASURED_PEAKS.json" if peaks else "fallback 6650 GB/s", ms_per_launch=1e3 * gram_s,
                         algorithmic_bytes_per_launch=nb_bytes, als_iterations=iters)
        solve_roof = None
    # `roofline` = the kernel with the largest share of the step; the other one goes under other_kernels
    if solve_roof is not None and ph[1] >= gram_s:
        roofline = solve_roof
        others["ring_s2_tc_kernel"] = gram_roof
    else:
        roofline = gram_roof
        if solve_roof is not None:
            others["ring_solve_mma_kernel"] = solve_roof
    roofline["hbm_iteration"] = dict(algorithmic_bytes=3.0 * d1 * d2 * T * 2 / world, gbs=3.0 * d1 * d2 * T * 2 / world / (t_max / args.steps) / 1e9,
                                     peak_gbs=peaks.get("hbm_gbs", 6650.0))
    roofline["other_kernels"] = others
    # ---- CPU baseline, bounded sample, rank 0
    cores = os.cpu_count()
    cpu = None
    if world == 1 and not args.no_cpu:
        dt, u, st = cpu_reference_step(None)
        dt2, u2, st = cpu_reference_step(st)
        cpu = dict(value=u2 / dt2, unit="pixel*frames/s", cores=cores, kind="port",
                   sample="%dx%dx%d, K=%d, ring r=%d, 1 steady-state iteration (%.1f s)" % (SAMPLE["d1"], SAMPLE["d2"], SAMPLE["T"], SAMPLE["K"], SAMPLE["ring"], dt2))
    line = dict(metric="pixels*frames/sec per spatial+temporal+BG update iter", value=value, unit="pixel*frames/s",
                n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * t_max / args.steps,
                higher_is_better=True, scaling=scaling, vs_baseline=None, dtype="f64 (exact int64 second moments from the u16 video)",
                data="synthetic",
                config=dict(workload=({"c3": "configs[2]", "c4": "configs[3] (2p, rank-1 nmf background instead of the ring)"}.get(args.workload, "configs[1]")) + ": synthetic %dx%dx%d uint16, %d neurons, ring-BG r=18 (120 nbrs, bg_ssub=BGSSUB), %dx%d patches (%d per GPU, 19-px halo blocks), nnls spatial, foopsi/ar1 OASIS (smin=-5, optimize_pars, optimize_b)".replace("BGSSUB", str(args.bg_ssub)) % (d1, d2, T, K, patch_dims[0], patch_dims[1], len(obj.owned_patches())),
                            patches=dict(grid=[int(obj.nr_patch), int(obj.nc_patch)], owned_by_rank0=[int(x) for x in obj.owned_patches()],
                                         gram_tensor=bool(lib.cnmfe_last_gram_was_tensor(obj._h))),
                            l2="inputs (%.1f GB resident video per GPU) larger than L2; no flush needed" % (float(sum(int(b[1] - b[0] + 1) * int(b[3] - b[2] + 1) for b in nblk)) * T * 2 / 1e9),
                            full_size_checks=checks, per_rank=per_rank,
                            seed=SEED, device_ms_per_step=1e3 * t_dev / args.steps, wall_ms_per_step=1e3 * wall / args.steps,
                            call_wall_ms_per_step=dict(zip(["update_background", "update_spatial", "update_temporal"], [float(x) / args.steps for x in call_ms])),
                            phase_ms_per_step=dict(zip(["gram", "ring_solve", "projections", "spatial_solve", "temporal_sweeps", "deconvTemporal", "other"],
                                                       [float(x) / args.steps for x in phases]))),
                clocks=clocks, gpu_launches=int(launches),
                e2e=dict(value=e2e_val, unit="pixel*frames/s", h2d_bytes_per_step=int(h2d), d2h_bytes_per_step=int(d2h),
                         note="Sources2D.update_* on host numpy state: every call syncs the host mirror -- H2D of what changed on the host; D2H of b0, A (values on the search pattern), C, kernel_pars, neuron_sn into host arrays every step; the ring weights W and C_raw / S (state no update reads back from the host) stay on the device until obj.W / obj.C_raw / obj.S are read; the uint16 video stays resident in HBM (loaded once, like the reference's mat_data)"),
                roofline=roofline, cpu_baseline=cpu)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def run_oasis_stress(args):
    """BASELINE.json configs[4]: OASIS-only stress, N traces x T frames AR(2) deconvolution (deconvolveCa(y,'ar2',
    'foopsi', pars, 'smin', -3)), one GPU.  Metric: samples/s.  `value`: traces resident on the device
    (cnmfe_deconvolve_dev), CUDA-event time; `e2e`: through the host entry point cnmfe_deconvolve with host buffers on a
    bounded slice (H2D + D2H inside).  Roofline: 24 B per sample (fp64 in, c and s out) against the HBM peak."""
    import ctypes
    import torch
    from cnmf_e_b200 import _lib
    from cnmf_e_b200.oasis import make_deconv_opts
    lib = _lib.lib()
    N, T = args.oasis_traces, args.oasis_frames
    g = (1.7, -0.712)
    rs = np.random.RandomState(3)
    # gen_data (functions/gen_data.m:30-41) restated: spikes from the MT19937 stream, AR(2) recursion, unit Gaussian noise
    yc = (rs.rand(T, N).T < 0.5 / 30).astype(np.float64)
    for t in range(2, T):
        yc[:, t] += g[0] * yc[:, t - 1] + g[1] * yc[:, t - 2]
    gen = torch.Generator(device="cuda")
    gen.manual_seed(3)
    y = (torch.from_numpy(yc).cuda() + torch.randn((N, T), generator=gen, device="cuda", dtype=torch.float64)).contiguous()
    d, pars, sn = make_deconv_opts(dict(type="ar2", method="foopsi", pars=list(g), smin=-3))
    pars_d = torch.tensor(np.tile(np.array(g), (N, 1)), device="cuda", dtype=torch.float64).contiguous()
    c = torch.empty_like(y); s_ = torch.empty_like(y)
    outs = torch.zeros((N, 6), device="cuda", dtype=torch.float64)
    V = ctypes.c_void_p

    def step():
        _lib.check(lib.cnmfe_deconvolve_dev(V(y.data_ptr()), T, N, ctypes.byref(d), None, V(pars_d.data_ptr()), V(c.data_ptr()),
                                            V(s_.data_ptr()), V(outs.data_ptr()), 0))
    sampler = ClockSampler(0)
    sampler.prepare()
    launches0 = lib.cnmfe_launch_count()
    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = lib.cnmfe_launch_count()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop()
    launches = lib.cnmfe_launch_count() - l0
    # parity at full length on a few traces (post-timing; the oracle is the checker)
    checks = {}
    try:
        from oracle import oasis as O
        k = 4
        ch, sh = c[:k].cpu().numpy(), s_[:k].cpu().numpy()
        yh = y[:k].cpu().numpy()
        bad, err = 0, 0.0
        for n in range(k):
            co, so, _ = O.deconvolveCa(yh[n], dict(type="ar2", method="foopsi", pars=list(g), smin=-3))
            bad += int(not np.array_equal(sh[n] > 0, so > 0))
            err = max(err, float(np.abs(ch[n] - co).max() / max(1.0, np.abs(co).max())))
        checks = dict(traces_checked_vs_oracle=k, spike_support_mismatches=bad, max_rel_err_c=err, ok=bool(bad == 0 and err < 1e-7))
    except Exception as e:
        checks = dict(error=repr(e))
    # e2e: host buffers through cnmfe_deconvolve on a slice of the traces
    ne = min(N, 500)
    Yh = np.ascontiguousarray(y[:ne].cpu().numpy())
    ph = np.tile(np.array(g), (ne, 1))
    ce = np.empty((ne, T)); se = np.empty((ne, T))
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    _lib.check(lib.cnmfe_deconvolve(P(Yh), T, ne, ctypes.byref(d), None, P(ph), P(ce), P(se), None, None, None, None, None, 0))
    t0 = time.perf_counter()
    _lib.check(lib.cnmfe_deconvolve(P(Yh), T, ne, ctypes.byref(d), None, P(ph), P(ce), P(se), None, None, None, None, None, 0))
    e2e_t = time.perf_counter() - t0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    gbs = 24.0 * N * T / (ms / 1e3) / 1e9
    print(json.dumps(dict(metric="OASIS AR2 deconvolution samples/s (configs[4])", value=N * T / (ms / 1e3), unit="samples/s",
                          n_gpus=1, steps=args.steps, warmup=max(3, args.warmup), ms_per_step=ms,
                          higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                          config=dict(workload="configs[4]: OASIS-only stress, %d traces x %d frames, AR2 foopsi (pars given, smin=-3), traces resident on the device" % (N, T),
                                      l2="inputs (%.1f GB) larger than L2" % (N * T * 8 / 1e9), parity=checks),
                          clocks=clocks, gpu_launches=int(launches),
                          e2e=dict(value=ne * T / e2e_t, unit="samples/s", h2d_bytes_per_step=int(Yh.nbytes + ph.nbytes), d2h_bytes_per_step=int(ce.nbytes + se.nbytes),
                                   note="cnmfe_deconvolve with host buffers on %d of the traces (pageable memory, H2D + D2H inside)" % ne),
                          roofline=dict(bound="hbm", kernel="deconv_batch_kernel (one CTA per trace: exact sequential AR2 pool-adjacent-violators)",
                                        achieved=gbs, peak=hbm, unit="GB/s", frac=gbs / hbm, traffic=None,
                                        algorithmic_bytes_per_launch=24.0 * N * T,
                                        note="AR2 PAV is a data-dependent sequential scan per trace (oasisAR2.m:78-140): latency bound, not bandwidth bound"),
                          cpu_baseline=None)))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=None, help="frames (default: 10000 for configs[1], 20000 for --workload c3)")
    ap.add_argument("--tensor", type=int, default=1)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-oracle-checks", action="store_true", help="skip the post-timing oracle spot checks of the full-size results")
    ap.add_argument("--bg-ssub", type=int, default=1, help="options.bg_ssub of the ring model (configs[1] is quoted at 1)")
    ap.add_argument("--workload", default="iteration", choices=["iteration", "c3", "c4", "oasis"])
    ap.add_argument("--oasis-traces", type=int, default=5000)
    ap.add_argument("--oasis-frames", type=int, default=100000)
    args = ap.parse_args()
    if args.workload == "oasis":
        run_oasis_stress(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
