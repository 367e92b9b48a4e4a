"""Host-side mirror of the reference's `Sources2D` handle class for the alternating-update hot path.

Same method names, argument meaning and side effects as
  ca_source_extraction/@Sources2D/update_background_parallel.m:1   (obj.W, obj.b0, obj.b0_new, obj.A_prev, obj.C_prev)
  ca_source_extraction/@Sources2D/update_spatial_parallel.m:1      (obj.A, obj.b0_new)
  ca_source_extraction/@Sources2D/update_temporal_parallel.m:1     (obj.C_raw, obj.C, obj.S, obj.P.kernel_pars, obj.P.neuron_sn)
plus the short aliases BASELINE.json names (update_background / update_spatial / update_temporal).
The video is uploaded once (the analogue of getReady + map_data_to_memory, Sources2D.m:185-194,214-265) and stays
resident on the GPU in its native integer dtype; every update runs in libcnmfe_b200.so (no CPU fallback).

Out of scope hooks, supplied by the caller exactly where the reference calls them:
  IND = determine_search_location(...)   -> `update_spatial_parallel(..., IND=...)` (or `self.search_fn`)
  post_process_spatial(...)              -> `self.post_process_fn` (default identity)
"""
import ctypes
import os
import sys

import numpy as np
import scipy.sparse as sp

from . import _lib as L
from .oasis import make_deconv_opts

_SPATIAL = {"hals": 0, "hals_thresh": 1, "nnls": 2, "lars": 3}


def patch_geometry(d1, d2, patch_dims, w_overlap):
    """patch_pos / block_pos of endoscope/distribute_data.m:39,55-77,163-173 (1-based inclusive [r0 r1 c0 c1])."""
    min_w = 2 * w_overlap + 3
    patch_dims = [max(int(p), min_w) for p in np.atleast_1d(patch_dims)]      # distribute_data.m:46
    if len(patch_dims) == 1:
        patch_dims = patch_dims * 2

    def idx(dn, pd, force_last):
        x = dn / pd
        npatch = int(np.floor(x + 0.5))            # MATLAB round (positive arguments)
        if npatch <= 1:
            return np.array([1, dn])
        pi = np.ceil(np.linspace(1, dn, npatch + 1)).astype(int)
        if force_last:
            pi[-1] = dn
        if pi[1] - pi[0] < min_w:
            pi = np.arange(1, dn + 1, min_w)
            pi[-1] = dn
        return pi

    pr = idx(d1, patch_dims[0], True)
    pc = idx(d2, patch_dims[1], False)
    nr_p, nc_p = len(pr) - 1, len(pc) - 1
    patch_pos = np.zeros((nr_p, nc_p, 4), dtype=np.int32)
    block_pos = np.zeros((nr_p, nc_p, 4), dtype=np.int32)
    for m in range(nr_p):
        for n in range(nc_p):
            patch_pos[m, n] = [pr[m], pr[m + 1] - (m != nr_p - 1), pc[n], pc[n + 1] - (n != nc_p - 1)]
            block_pos[m, n] = [max(1, pr[m] - w_overlap - 1), min(d1, pr[m + 1] + w_overlap),
                               max(1, pc[n] - w_overlap - 1), min(d2, pc[n + 1] + w_overlap)]
    return patch_pos, block_pos


def block_indices(patch_pos, d1, d2, w_overlap):
    """block_idx_r, block_idx_c of distribute_data.m:91-98 (1-based): every patch border x contributes x-1-w and x+w,
    clamped to the FOV, sorted, unique (so even a single patch gives three blocks per dimension)."""
    def one(pidx, dn):
        b = set()
        for x in pidx:
            b.add(min(max(int(x) - 1 - w_overlap, 1), dn))
            b.add(min(max(int(x) + w_overlap, 1), dn))
        return np.array(sorted(b))
    pr = sorted(set(int(x) for x in patch_pos[:, 0, 0]) | {int(patch_pos[-1, 0, 1])})
    pc = sorted(set(int(x) for x in patch_pos[0, :, 2]) | {int(patch_pos[0, -1, 3])})
    return one(pr, d1), one(pc, d2)


def trace_range(K, rank, world_size):
    """Contiguous block [k0, k1) of the K traces handled by `rank` when the final deconvTemporal is split over ranks
    (sizes differ by at most one)."""
    base, rem = divmod(int(K), int(world_size))
    k0 = rank * base + min(rank, rem)
    return k0, k0 + base + (1 if rank < rem else 0)


def patch_owners(npatch, world_size):
    """Blocked patch -> rank assignment (patches in MATLAB linear order): rank r owns a contiguous run."""
    return np.array([(i * world_size) // npatch for i in range(npatch)], dtype=np.int64)


def merge_temporal(num, den):
    """C_raw = sum_p aa_p C_raw,p / sum_p aa_p with 0 -> 1 denominators (update_temporal_parallel.m:269-280)."""
    den = np.where(den == 0, 1.0, den)
    return num * (1.0 / den)[:, None]


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class _PinnedPool:
    """Page-locked host buffers for the large device->host copies (ring weights: 252 MB per 512 x 512 patch; C / C_raw / S:
    K*T*8 each).  A buffer is handed out again only when NOTHING references the array it backed any more -- an array the
    caller (or this class, e.g. obj.C_prev) still holds is never overwritten; the pool then grows by one buffer instead."""

    def __init__(self, lib):
        self._lib = lib
        self._bufs = {}          # (name, shape) -> [base arrays]

    def take(self, name, shape):
        shape = tuple(int(x) for x in shape)
        # buffers of this name with another shape (K changed): drop the ones nobody uses any more
        for key in [k for k in self._bufs if k[0] == name and k[1] != shape]:
            keep = []
            for b in self._bufs[key]:
                if sys.getrefcount(b) > 3:
                    keep.append(b)
                else:
                    self._lib.cnmfe_host_unregister(_ptr(b))
            if keep:
                self._bufs[key] = keep
            else:
                del self._bufs[key]
        lst = self._bufs.setdefault((name, shape), [])
        for b in lst:
            if sys.getrefcount(b) <= 3:       # the list, the loop variable, getrefcount's argument: no view of it is alive
                v = b.view()
                v.setflags(write=True)
                return v
        b = np.empty(shape)
        if b.nbytes >= (1 << 20) and self._lib.cnmfe_host_register(_ptr(b), b.nbytes) == 0:
            lst.append(b)
        return b.view()

    def close(self):
        for lst in self._bufs.values():
            for b in lst:
                self._lib.cnmfe_host_unregister(_ptr(b))
        self._bufs = {}


class _RingWeights(dict):
    """obj.W: patch -> slot-form ring weights.  After a background update the weights stay on the device (252 MB per 512 x 512
    patch) and are fetched the first time they are READ here -- the reference keeps obj.W as state nobody but the next update
    touches (update_background_parallel.m:311-317), so copying them out after every call would only tax the link."""

    def __init__(self, owner):
        super().__init__()
        self._owner = owner

    def _sync(self, k=None):
        o = self._owner
        if o._ring_w_stale and (k is None or k in o._ring_w_stale):
            o._pull_ring_weights(sorted(o._ring_w_stale) if k is None else [k])

    def __getitem__(self, k):
        self._sync(k)
        return dict.__getitem__(self, k)

    def get(self, k, default=None):
        self._sync(k)
        return dict.get(self, k, default)

    def __contains__(self, k):
        return dict.__contains__(self, k) or k in self._owner._ring_w_stale

    def __setitem__(self, k, v):
        self._owner._ring_w_stale.discard(k)
        dict.__setitem__(self, k, v)

    def items(self):
        self._sync()
        return dict.items(self)

    def values(self):
        self._sync()
        return dict.values(self)


class Sources2D:
    def __init__(self, d1, d2, T, patch_dims=None, ring_radius=18, num_neighbors=None, device=0, rank=0,
                 world_size=1, options=None):
        self.d1, self.d2, self.T = int(d1), int(d2), int(T)
        self.device, self.rank, self.world_size = device, rank, world_size
        if patch_dims is None:
            patch_dims = (d1, d2)
        # CNMFSetParms.m:286-308 defaults + demo_large_data_1p.m:11-55 overrides relevant to this path
        self.options = dict(ring_radius=ring_radius, bg_ssub=1, background_model="ring", bg_acceleration=True,
                            num_neighbors=num_neighbors, thresh_outlier=np.nan, spatial_algorithm="hals", maxIter=5,
                            deconv_flag=True, deconv_options=dict(type="ar1", method="foopsi", smin=-5,
                                                                  optimize_pars=True, optimize_b=True, max_tau=100),
                            replicate_spatial_aprev_quirk=True, use_tensor_gram=True, nb=1,
                            # multi-GPU: split the final deconvTemporal over ranks by trace; validated against the
                            # oracle on 2 and 4 GPUs by tests/test_gpu_multi.py
                            shard_deconv=bool(int(os.environ.get("CNMFE_SHARD_DECONV", "1"))))
        if options:
            self.options.update(options)
        self.patch_pos, self.block_pos = patch_geometry(self.d1, self.d2, patch_dims, ring_radius)
        self.nr_patch, self.nc_patch = self.patch_pos.shape[:2]
        self.patches = [(m, n) for n in range(self.nc_patch) for m in range(self.nr_patch)]   # MATLAB linear order
        self.npatch = len(self.patches)
        # patch -> rank (blocked assignment)
        self.owner = patch_owners(self.npatch, world_size)
        owned = (self.owner == rank).astype(np.uint8)
        pp = np.ascontiguousarray(np.stack([self.patch_pos[mp] for mp in self.patches]).astype(np.int32))
        bp = np.ascontiguousarray(np.stack([self.block_pos[mp] for mp in self.patches]).astype(np.int32))
        self._lib = L.lib()
        h = ctypes.c_void_p()
        L.check(self._lib.cnmfe_create(ctypes.byref(h), self.d1, self.d2, self.T, self.npatch, _ptr(pp), _ptr(bp),
                                       _ptr(owned), int(ring_radius), int(num_neighbors or 0), int(device)))
        self._h = h
        L.check(self._lib.cnmfe_set_trace_major(self._h, 1))   # NumPy (K, T) C-order arrays are trace-contiguous
        self._owned = owned.astype(bool)
        nnb = ctypes.c_int()
        L.check(self._lib.cnmfe_ring_offsets(self._h, ctypes.byref(nnb), None, None))
        self.nnb = nnb.value
        rs = np.zeros(self.nnb, dtype=np.int32)
        cs = np.zeros(self.nnb, dtype=np.int32)
        L.check(self._lib.cnmfe_ring_offsets(self._h, ctypes.byref(nnb), _ptr(rs), _ptr(cs)))
        self.r_shift, self.c_shift = rs, cs
        d = self.d1 * self.d2
        self.A = sp.csc_matrix((d, 0))
        self.C = np.zeros((0, self.T))
        self._C_raw = np.zeros((0, self.T))
        self._S = np.zeros((0, self.T))
        self._lazy_temporal = set()      # {"C_raw", "S"}: newer on the device than on the host, fetched when read
        self.A_prev = sp.csc_matrix((d, 0))
        self.C_prev = np.zeros((0, self.T))
        self._ring_w_stale = set()  # patches whose weights on the device are newer than the host copy (fetched on read)
        self.W = _RingWeights(self)
        self.b0 = {}
        self._temporal_partial = set()   # multi-rank: which of (1: C_raw, 2: S, 3: per-trace outputs) still hold only this rank's rows on the device
        self._ring_synced = {}      # patch -> (W, b0) objects whose contents the device holds
        self._dev = {}              # "A" / "C" / "A_prev" / "C_prev" -> host object whose contents the device holds
        self._pool = _PinnedPool(self._lib)   # page-locked buffers of the large pulls (never reused while referenced)
        self.h2d_bytes = 0          # bytes this object has sent to / fetched from the device (state, not the video)
        self.d2h_bytes = 0
        self.b = {}
        self.f = {}
        self.b0_new = np.zeros((self.d1, self.d2))
        self.P = dict(sn=np.ones((self.d1, self.d2)), kernel_pars=None, neuron_sn=None, Ymean=None)
        self.search_fn = None
        self.post_process_fn = None
        self.frame_range = (1, self.T)

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_h", None):
            self._pool.close()
            self._lib.cnmfe_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ data
    def owned_patches(self):
        return [i for i in range(self.npatch) if self._owned[i]]

    def block_of(self, i):
        return self.block_pos[self.patches[i]]

    def patch_of(self, i):
        return self.patch_pos[self.patches[i]]

    def load_video(self, Y):
        """Y: (d1,d2,T) uint8/uint16 array (numpy).  Uploads the blocks of the owned patches (get_patch_data.m with
        with_overlap=true) and keeps them resident."""
        Y = np.asarray(Y)
        codes = {np.dtype(np.uint8): 0, np.dtype(np.uint16): 1, np.dtype(np.float32): 2, np.dtype(np.float64): 3}
        if Y.dtype not in codes:
            raise L.CnmfeError("video dtype %s unsupported: uint8 / uint16, or single / double holding integer counts" % Y.dtype)
        dt = codes[Y.dtype]
        for i in self.owned_patches():
            b = self.block_of(i)
            blk = Y[b[0] - 1:b[1], b[2] - 1:b[3], :]
            # frame-major, pixel index r + c*nrb fastest  == MATLAB memory order of the (nrb, ncb, T) array
            buf = np.ascontiguousarray(np.transpose(blk, (2, 1, 0)))
            L.check(self._lib.cnmfe_upload_block(self._h, i, _ptr(buf), dt))
        self.P["Ymean"] = Y.mean(axis=2, dtype=np.float64)     # obj.P.Ymean (cell2mat'ed); load_block_dev callers set it themselves

    def load_block_dev(self, i, dev_ptr, dtype):
        """Block of patch i already on the device (frame-major); dtype 0 = uint8, 1 = uint16."""
        L.check(self._lib.cnmfe_upload_block_dev(self._h, i, ctypes.c_void_p(dev_ptr), dtype))

    # ------------------------------------------------------------------ option / state marshalling
    def _push_options(self):
        o = L.Options()
        self._lib.cnmfe_options_defaults(ctypes.byref(o))
        opt = self.options
        model = str(opt["background_model"]).lower()
        if model not in ("ring", "svd", "nmf"):
            raise L.CnmfeError("background_model must be 'ring', 'svd' or 'nmf'")
        o.thresh_outlier = float("nan") if opt["thresh_outlier"] is None else float(opt["thresh_outlier"])
        o.spatial_algorithm = _SPATIAL[str(opt["spatial_algorithm"]).lower()] if \
            str(opt["spatial_algorithm"]).lower() in _SPATIAL else 2
        o.maxIter_temporal = int(opt["maxIter"])
        o.deconv_flag = int(bool(opt["deconv_flag"]))
        o.bg_acceleration = int(bool(opt["bg_acceleration"]))
        o.replicate_spatial_aprev_quirk = int(bool(opt["replicate_spatial_aprev_quirk"]))
        o.use_tensor_gram = int(bool(opt["use_tensor_gram"]))
        o.background_model = {"ring": 0, "svd": 1, "nmf": 2}[model]
        o.nb = int(opt.get("nb", 1))
        o.bg_ssub = int(opt.get("bg_ssub", 1)) if model == "ring" else 1
        if opt["deconv_flag"]:
            d, _, _ = make_deconv_opts(opt["deconv_options"] or {})
            o.deconv = d
        L.check(self._lib.cnmfe_set_options(self._h, ctypes.byref(o)))

    @staticmethod
    def _csc(A):
        A = sp.csc_matrix(A)
        A.sort_indices()
        return (np.ascontiguousarray(A.indptr, dtype=np.int64), np.ascontiguousarray(A.indices, dtype=np.int64),
                np.ascontiguousarray(A.data, dtype=np.float64))

    # Host <-> device coherence.  Arrays this class hands out (pull_*) are read-only, so "same object" means "same
    # contents": a part whose host object is the one the device was last synchronised with is not sent again.  Code
    # that must edit in place makes a writable copy and assigns it back (a new object => re-sent), or calls
    # invalidate().
    def invalidate(self):
        self._dev = {}
        self._ring_synced = {}

    @staticmethod
    def _freeze(x):
        if sp.issparse(x):
            for part in (x.data, x.indices, x.indptr):
                part.setflags(write=False)
        else:
            x.setflags(write=False)
        return x

    @staticmethod
    def _frozen(x):
        return (not x.data.flags.writeable) if sp.issparse(x) else (not x.flags.writeable)

    def _mark(self, name, obj):
        """The device now holds the contents of `obj` -- remembered only if obj is read-only NOW (a writable array could be
        modified and frozen later, which would make stale device state look current)."""
        self._dev[name] = obj if self._frozen(obj) else None

    def _current(self, name, obj):
        return self._dev.get(name) is obj and self._frozen(obj)

    def _push_pair(self, setter, nameA, A, nameC, C):
        K = A.shape[1]
        C = np.asarray(C)
        assert C.shape == (K, self.T), "%s must be K x T" % nameC
        sendA, sendC = not self._current(nameA, A), not self._current(nameC, C)
        if not (sendA or sendC):
            return
        jc = ir = pr = Cc = None
        if sendA or K == 0:
            jc, ir, pr = self._csc(A)
        if sendC:
            Cc = np.ascontiguousarray(C, dtype=np.float64)
        L.check(setter(self._h, K, _ptr(jc), _ptr(ir), _ptr(pr), _ptr(Cc)))
        self.h2d_bytes += sum(x.nbytes for x in (jc, ir, pr, Cc) if x is not None)
        self._mark(nameA, A)
        self._mark(nameC, C)

    def set_sn(self, sn=None):
        """obj.P.sn -> device (d1 x d2 noise map used by hals_thresh / lars and the outlier clamp)."""
        if sn is not None:
            self.P["sn"] = np.asarray(sn, dtype=np.float64)
        L.check(self._lib.cnmfe_set_sn(self._h, _ptr(np.asfortranarray(self.P["sn"], dtype=np.float64))))
        self.h2d_bytes += self.P["sn"].size * 8

    def push_neurons(self):
        self._push_pair(self._lib.cnmfe_set_neurons, "A", self.A, "C", self.C)

    def push_prev(self):
        self._push_pair(self._lib.cnmfe_set_prev, "A_prev", self.A_prev, "C_prev", self.C_prev)

    def push_ring(self):
        if str(self.options["background_model"]).lower() in ("svd", "nmf"):
            for i in range(self.npatch):
                b, f, b0 = self.b.get(i), self.f.get(i), self.b0.get(i)
                if b is None and f is None and b0 is None:
                    continue
                bf = None if b is None else np.asfortranarray(b, dtype=np.float64)
                ff = None if f is None else np.asfortranarray(f, dtype=np.float64)
                b0f = None if b0 is None else np.ascontiguousarray(b0, dtype=np.float64)
                L.check(self._lib.cnmfe_set_bf(self._h, i, _ptr(bf), _ptr(ff), _ptr(b0f)))
            return
        for i in range(self.npatch):
            stale = i in self._ring_w_stale          # the device holds newer weights than the host: nothing to send
            W = None if stale else dict.get(self.W, i)
            b0 = self.b0.get(i)
            if W is None and b0 is None:
                continue
            # arrays handed out by pull_ring are read-only: an unchanged identity means the device copy is current
            # (the ring weights are ~250 MB at 512x512; re-sending them on every call would dominate the step)
            s = self._ring_synced.get(i)
            if s is not None and (stale or s[0] is W) and s[1] is b0:
                continue
            Wf = None if W is None else np.ascontiguousarray(W, dtype=np.float64)   # (d_patch, nnb) C-order == nnb x d_patch col-major
            b0f = None if b0 is None else np.ascontiguousarray(b0, dtype=np.float64)
            L.check(self._lib.cnmfe_set_ring(self._h, i, _ptr(Wf), _ptr(b0f)))
            self.h2d_bytes += sum(x.nbytes for x in (Wf, b0f) if x is not None)
            # identity is only evidence of "unchanged" for read-only arrays (a writable one may be edited in place)
            ro = all(x is None or not x.flags.writeable for x in (W, b0))
            self._ring_synced[i] = (W if not stale else (s[0] if s else None), b0) if ro else None

    def pull_ring(self, weights=True):
        """obj.b0 (and, with weights=True, obj.W) of the owned patches from the device.  weights=False marks the weights as
        "newer on the device": they are fetched when obj.W[i] is read."""
        if str(self.options["background_model"]).lower() in ("svd", "nmf"):
            nb = int(self.options.get("nb", 1))
            for i in self.owned_patches():
                p = self.patch_of(i)
                dp = (p[1] - p[0] + 1) * (p[3] - p[2] + 1)
                b = np.zeros((dp, nb), order="F"); f = np.zeros((nb, self.T), order="F"); b0 = np.zeros(dp)
                L.check(self._lib.cnmfe_get_bf(self._h, i, _ptr(b), _ptr(f), _ptr(b0)))
                self.d2h_bytes += b.nbytes + f.nbytes + b0.nbytes
                self.b[i], self.f[i], self.b0[i] = np.ascontiguousarray(b), np.ascontiguousarray(f), b0
            return
        for i in self.owned_patches():
            p = self.patch_of(i)
            dp = (p[1] - p[0] + 1) * (p[3] - p[2] + 1)
            b0 = np.empty(dp)
            L.check(self._lib.cnmfe_get_ring(self._h, i, None, _ptr(b0)))
            self.d2h_bytes += b0.nbytes
            b0.setflags(write=False)
            self.b0[i] = b0
            prev = self._ring_synced.get(i)
            self._ring_synced[i] = (prev[0] if prev else None, b0)
            self._ring_w_stale.add(i)
        if weights:
            self._pull_ring_weights(self.owned_patches())

    def _pull_ring_weights(self, which):
        for i in which:
            p = self.patch_of(i)
            dp = (p[1] - p[0] + 1) * (p[3] - p[2] + 1)
            if int(self.options.get("bg_ssub", 1)) > 1:
                d1s, d2s, nnb, _, _ = self.ssub_dims(i)
                wshape = (d1s * d2s, nnb)
            else:
                wshape = (dp, self.nnb)
            # W lands in a page-locked buffer that is reused only once nobody references the array it backed before
            dict.pop(self.W, i, None)
            W = self._pool.take(("W", i), wshape)
            L.check(self._lib.cnmfe_get_ring(self._h, i, _ptr(W), None))
            self.d2h_bytes += W.nbytes
            W.setflags(write=False)
            dict.__setitem__(self.W, i, W)
            self._ring_w_stale.discard(i)
            prev = self._ring_synced.get(i)
            self._ring_synced[i] = (W, prev[1] if prev else None)

    def ssub_dims(self, i):
        """(d1s, d2s, nnb, r_shift, c_shift) of the coarse ring grid of patch i (bg_ssub > 1)."""
        d1s, d2s, nnb = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        L.check(self._lib.cnmfe_ssub_dims(self._h, i, ctypes.byref(d1s), ctypes.byref(d2s), ctypes.byref(nnb), None, None))
        rs = np.zeros(nnb.value, dtype=np.int32); cs = np.zeros(nnb.value, dtype=np.int32)
        L.check(self._lib.cnmfe_ssub_dims(self._h, i, ctypes.byref(d1s), ctypes.byref(d2s), ctypes.byref(nnb), _ptr(rs), _ptr(cs)))
        return d1s.value, d2s.value, nnb.value, rs, cs

    def ring_as_sparse_ssub(self, i):
        """obj.W{i} for bg_ssub > 1: sparse (d1s*d2s x d1s*d2s) on the coarse grid (initComponents_parallel.m:237-253)."""
        d1s, d2s, nnb, rs, cs = self.ssub_dims(i)
        W = self.W[i]
        rr = np.tile(np.arange(d1s), d2s)
        cc = np.repeat(np.arange(d2s), d1s)
        rows, cols, vals = [], [], []
        for s in range(nnb):
            r2, c2 = rr + rs[s], cc + cs[s]
            ok = (r2 >= 0) & (r2 < d1s) & (c2 >= 0) & (c2 < d2s)
            rows.append(np.nonzero(ok)[0]); cols.append((c2 * d1s + r2)[ok]); vals.append(W[ok, s])
        return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(d1s * d2s, d1s * d2s))

    def ring_as_sparse(self, i):
        """obj.W{i} as the reference stores it: sparse (d_patch x d_block), initComponents_parallel.m:221-236."""
        W = self.W[i]
        p, b = self.patch_of(i), self.block_of(i)
        nr, nc = p[1] - p[0] + 1, p[3] - p[2] + 1
        nrb, ncb = b[1] - b[0] + 1, b[3] - b[2] + 1
        rr = np.tile(np.arange(p[0], p[1] + 1), nc)
        cc = np.repeat(np.arange(p[2], p[3] + 1), nr)
        rows, cols, vals = [], [], []
        for s in range(self.nnb):
            r2, c2 = rr + self.r_shift[s], cc + self.c_shift[s]
            ok = (r2 >= 1) & (r2 <= self.d1) & (c2 >= 1) & (c2 <= self.d2)
            jj = (c2 - b[2]) * nrb + (r2 - b[0])
            rows.append(np.nonzero(ok)[0])
            cols.append(jj[ok])
            vals.append(W[ok, s])
        return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                             shape=(nr * nc, nrb * ncb))

    def reconstruct_b0(self):
        out = np.zeros((self.d1, self.d2))
        for i in self.owned_patches():
            p = self.patch_of(i)
            out[p[0] - 1:p[1], p[2] - 1:p[3]] = self.b0[i].reshape(p[1] - p[0] + 1, p[3] - p[2] + 1, order="F")
        return out

    # ------------------------------------------------------------------ the three updates
    def update_background_parallel(self, use_parallel=True, sync_host=True):
        """update_background_parallel(obj, use_parallel).  `use_parallel` is accepted for signature compatibility; the
        library does its own intra-call concurrency (SURVEY.md §8b threading)."""
        self._push_options()
        if sync_host:
            self.push_neurons()
            self.push_ring()
            to = self.options["thresh_outlier"]
            if to is not None and not np.isnan(to):
                self.set_sn()                      # the outlier clamp compares with thresh_outlier * obj.P.sn (fit_ring_model.m:51)
        L.check(self._lib.cnmfe_update_background(self._h))
        if not sync_host and self._is_ring():
            self._ring_w_stale.update(self.owned_patches())
        if sync_host:
            self.pull_ring(weights=False)          # b0 now, W when somebody reads obj.W[i]
            self.b0_new = self.reconstruct_b0()
            if self.world_size > 1:     # patches are disjoint: the sum over ranks assembles the full map (update_background_parallel.m:315)
                self.b0_new = self._allreduce_sum(self.b0_new.ravel()).reshape(self.d1, self.d2)
            # obj.A_prev = obj.A; obj.C_prev = obj.C (update_background_parallel.m:316-317): read-only state is shared,
            # not copied; the device took the same snapshot
            self.A_prev = self.A if self._frozen(self.A) else self._freeze(self.A.copy())
            self.C_prev = self.C if self._frozen(self.C) else self._freeze(np.array(self.C, copy=True))
            self._mark("A_prev", self.A_prev)
            self._mark("C_prev", self.C_prev)

    def estimate_noise(self, Y=None, frame_range=None, chunk=32768, replicate_block_quirk=True):
        """sn = estimate_noise(obj, frame_range, 'psd') (@Sources2D/Sources2D.m:328-379): per-pixel GetSn of the raw video on
        frames [1, min(T, 3000)] by default (1-based inclusive, as in the reference).  Returns the (d1, d2) map; like the
        reference it does NOT assign obj.P.sn (the caller does, demo_large_data_1p.m / initComponents_parallel.m).
        Y = None (single rank): the rows are read from the video resident on the device (cnmfe_estimate_noise); Y = (d1, d2, T)
        host array: its rows go through the library's GetSn kernel in chunks.
        replicate_block_quirk: the reference assembles the map from the blocks of distribute_data.m:91-98 and deletes row /
        column `end-1` (not `end`) of every non-last block (Sources2D.m:369-374), so the map position just before an interior
        block border holds the value of the border pixel itself; replicated by default, False gives the plain per-pixel map."""
        from . import oasis as G
        d = self.d1 * self.d2
        if Y is None:
            if frame_range is None:
                frame_range = (1, min(self.T, 3000))
            if self.world_size != 1:
                raise L.CnmfeError("estimate_noise from the resident video needs all patches on this rank; pass Y")
            out = np.zeros((self.d1, self.d2), order="F")
            L.check(self._lib.cnmfe_estimate_noise(self._h, int(frame_range[0]), int(frame_range[1]), _ptr(out)))
            sn = np.ascontiguousarray(out)
        else:
            Y = np.asarray(Y)
            T = Y.shape[2]
            if frame_range is None:
                frame_range = (1, min(T, 3000))
            f0, f1 = int(frame_range[0]) - 1, int(frame_range[1])
            snv = np.empty(d)
            Yr = Y.reshape(d, T, order="F")                      # pixel index r + c*d1, MATLAB order
            for p0 in range(0, d, int(chunk)):
                rows = np.ascontiguousarray(Yr[p0:p0 + int(chunk), f0:f1], dtype=np.float64)
                snv[p0:p0 + rows.shape[0]] = G.GetSn(rows, device=self.device)
            sn = snv.reshape(self.d1, self.d2, order="F")
        if replicate_block_quirk:
            br, bc = block_indices(self.patch_pos, self.d1, self.d2, int(self.options["ring_radius"]))
            rmap, cmap = np.arange(self.d1), np.arange(self.d2)
            for b in br[1:-1]:
                rmap[b - 2] = b - 1          # 0-based: position of pixel b-1 shows pixel b
            for b in bc[1:-1]:
                cmap[b - 2] = b - 1
            sn = np.ascontiguousarray(sn[np.ix_(rmap, cmap)])
        return sn

    # ---- diagnostics on the fused BG subtraction ---------------------------------------------------
    def _b0_maps(self):
        b0 = self.reconstruct_b0()
        if self.world_size > 1:
            b0 = self._allreduce_sum(b0.ravel()).reshape(self.d1, self.d2)
        return np.asfortranarray(b0), np.asfortranarray(np.asarray(self.b0_new, dtype=np.float64))

    def compute_RSS(self, frame_range=None):
        """[RSS_total, RSS] = compute_RSS(obj, frame_range) (@Sources2D/Sources2D.m:1358-1510; ring model, bg_ssub = 1).  Uses the
        state on the host (A, C, A_prev, C_prev, W, b0, b0_new), sent to the device if it changed.  Multi-rank: a collective
        call; RSS of the patches this rank owns are summed over the ranks."""
        f0, f1 = (1, self.T) if frame_range is None else (int(frame_range[0]), int(frame_range[1]))
        self._push_options(); self.push_neurons(); self.push_prev(); self.push_ring()
        b0, b0n = self._b0_maps()
        rss = np.zeros(self.npatch)
        L.check(self._lib.cnmfe_compute_rss(self._h, f0, f1, _ptr(b0), _ptr(b0n), _ptr(rss)))
        if self.world_size > 1:
            rss = self._allreduce_sum(rss)
        self.P["RSS"] = float(rss.sum())
        return self.P["RSS"], rss

    def reconstruct_background(self, frame_range=None):
        """Ybg = reconstruct_background(obj, frame_range) (@Sources2D/Sources2D.m:1247-1356; ring model, bg_ssub = 1): (d1, d2,
        nframes) array, filled on the patches this rank owns."""
        f0, f1 = (1, self.T) if frame_range is None else (int(frame_range[0]), int(frame_range[1]))
        self._push_options(); self.push_prev(); self.push_ring()
        b0, b0n = self._b0_maps()
        nf = f1 - f0 + 1
        out = np.zeros((self.d1, self.d2, nf))
        for i in self.owned_patches():
            p = self.patch_of(i)
            nr, nc = p[1] - p[0] + 1, p[3] - p[2] + 1
            buf = np.empty((nr * nc, nf))              # [pixel][frame] (trace-major boundary layout)
            L.check(self._lib.cnmfe_reconstruct_background(self._h, i, f0, f1, _ptr(b0), _ptr(b0n), _ptr(buf)))
            self.d2h_bytes += buf.nbytes
            out[p[0] - 1:p[1], p[2] - 1:p[3], :] = buf.reshape(nr, nc, nf, order="F")
        return out

    # ---- host-side brackets of the spatial update (library C++: csrc/host_spatial.cu) ----------------------------------
    def determine_search_location(self, A=None, min_size=3.0, max_size=8.0, dist=3.0, method="ellipse", nrgthr=0.9999,
                                  nb=1, bSiz=3):
        """IND = determine_search_location(obj.A, method, options) (utilities/determine_search_location.m): 'ellipse' (:57-92,
        the reference's default search_method) or 'dilate' (:93-99: threshold_components + imdilate with a disk of radius
        bSiz), as a (d, K) boolean csc matrix.  Usable as `search_fn`."""
        A = sp.csc_matrix(self.A if A is None else A, dtype=np.float64)
        A.sort_indices()
        K = A.shape[1]
        jc = np.ascontiguousarray(A.indptr, dtype=np.int64)
        ir = np.ascontiguousarray(A.indices, dtype=np.int64)
        pr = np.ascontiguousarray(A.data, dtype=np.float64)
        m = str(method).lower()
        if m not in ("ellipse", "dilate") or (m == "ellipse" and np.isinf(dist)):
            # determine_search_location.m:91 (ellipse with dist == Inf) and the `otherwise` branch (:100-101): IND = true(d, K),
            # all-zero components excluded (:103-105)
            full = np.ones((self.d1 * self.d2, K), dtype=bool)
            full[:, np.asarray(A.sum(axis=0)).ravel() == 0] = False
            return sp.csc_matrix(full)
        if str(method).lower() == "dilate":
            cap = 1
            for k in range(K):
                rows = ir[jc[k]:jc[k + 1]]
                h = w = 1
                if rows.size:
                    r, c = rows % self.d1, rows // self.d1
                    h, w = int(r.max() - r.min() + 1), int(c.max() - c.min() + 1)
                cap += (h + 4 + 2 * int(bSiz)) * (w + 4 + 2 * int(bSiz))
            ojc = np.zeros(K + 1, dtype=np.int64)
            oir = np.zeros(cap, dtype=np.int64)
            L.check(self._lib.cnmfe_search_location_dilate(self.d1, self.d2, K, _ptr(jc), _ptr(ir), _ptr(pr), float(nrgthr), int(nb),
                                                           int(bSiz), _ptr(ojc), _ptr(oir), cap))
            n = int(ojc[K])
            return sp.csc_matrix((np.ones(n, dtype=bool), oir[:n].copy(), ojc), shape=(self.d1 * self.d2, K))
        reach = int(np.ceil(dist * max(max_size, min_size)))
        cap = max(1, K * (2 * reach + 2) ** 2)
        ojc = np.zeros(K + 1, dtype=np.int64)
        oir = np.zeros(cap, dtype=np.int64)
        L.check(self._lib.cnmfe_search_location_ellipse(self.d1, self.d2, K, _ptr(jc), _ptr(ir), _ptr(pr), float(min_size),
                                                        float(max_size), float(dist), _ptr(ojc), _ptr(oir), cap))
        n = int(ojc[K])
        return sp.csc_matrix((np.ones(n, dtype=bool), oir[:n].copy(), ojc), shape=(self.d1 * self.d2, K))

    def post_process_spatial(self, A_new=None, thr=0.01, sz=5, connected=True, circular=False):
        """A_ = post_process_spatial(obj, A_new) (@Sources2D/post_process_spatial.m:19-32); the CNMFSetParms default is
        spatial_constraints = struct('circular', false, 'connected', true).  Usable as `post_process_fn`."""
        A = sp.csc_matrix(self.A if A_new is None else A_new, dtype=np.float64).copy()
        A.sort_indices()
        K = A.shape[1]
        jc = np.ascontiguousarray(A.indptr, dtype=np.int64)
        ir = np.ascontiguousarray(A.indices, dtype=np.int64)
        pr = np.ascontiguousarray(A.data, dtype=np.float64)
        if connected:
            L.check(self._lib.cnmfe_connectivity_constraint(self.d1, self.d2, K, _ptr(jc), _ptr(ir), _ptr(pr), float(thr), int(sz)))
        out = sp.csc_matrix((pr, ir, jc), shape=A.shape)
        out.eliminate_zeros()
        if circular and K > 0:
            out.sort_indices()
            jc = np.ascontiguousarray(out.indptr, dtype=np.int64)
            ir = np.ascontiguousarray(out.indices, dtype=np.int64)
            pr = np.ascontiguousarray(out.data, dtype=np.float64)
            # the result lives on the bounding boxes of the footprints
            cap = 1
            for k in range(K):
                rows = ir[jc[k]:jc[k + 1]]
                if rows.size:
                    r, c = rows % self.d1, rows // self.d1
                    cap += int((r.max() - r.min() + 1) * (c.max() - c.min() + 1))
            ojc = np.zeros(K + 1, dtype=np.int64); oir = np.zeros(cap, dtype=np.int64); opr = np.zeros(cap)
            L.check(self._lib.cnmfe_circular_constraints(self.d1, self.d2, K, _ptr(jc), _ptr(ir), _ptr(pr), _ptr(ojc), _ptr(oir),
                                                         _ptr(opr), cap))
            n = int(ojc[K])
            out = sp.csc_matrix((opr[:n].copy(), oir[:n].copy(), ojc), shape=A.shape)
        return out

    def update_spatial_parallel(self, use_parallel=True, update_sn=False, IND=None, sync_host=True):
        """update_spatial_parallel(obj, use_parallel, update_sn).  IND: (d,K) boolean search mask
        (determine_search_location output, update_spatial_parallel.m:66); defaults to self.search_fn(self) or, without a hook,
        to the library's determine_search_location with options.search_method ('ellipse').  Post-processing
        (post_process_spatial, :341) runs when post_process_fn is set, e.g. `obj.post_process_fn = obj.post_process_spatial`."""
        self._push_options()
        if IND is None:
            # update_spatial_parallel.m:62-66: IND = determine_search_location(obj.A, options.search_method, options)
            if self.search_fn is not None:
                IND = self.search_fn(self)
            else:
                o = self.options
                IND = self.determine_search_location(self.A, method=o.get("search_method", "ellipse"),
                                                     min_size=o.get("min_size", 3.0), max_size=o.get("max_size", 8.0),
                                                     dist=o.get("dist", 3.0), nrgthr=o.get("nrgthr", 0.9999),
                                                     nb=o.get("nb", 1), bSiz=o.get("bSiz", 3))
        INDc = sp.csc_matrix(IND).astype(bool)
        INDc.sort_indices()
        if sync_host:
            self.push_neurons()
            self.push_prev()
            self.push_ring()
            self.set_sn()
        jc = np.ascontiguousarray(INDc.indptr, dtype=np.int64)
        ir = np.ascontiguousarray(INDc.indices, dtype=np.int64)
        self._ind_jc, self._ind_ir = jc, ir
        L.check(self._lib.cnmfe_set_search(self._h, INDc.shape[1], _ptr(jc), _ptr(ir)))
        self.h2d_bytes += jc.nbytes + ir.nbytes
        L.check(self._lib.cnmfe_update_spatial_ex(self._h, int(bool(update_sn))))
        if update_sn:
            snm = np.zeros((self.d1, self.d2), order="F")
            L.check(self._lib.cnmfe_get_sn_map(self._h, _ptr(snm)))
            self.P["sn"] = self._allreduce_owned_pixels(np.ascontiguousarray(snm)) if self.world_size > 1 else np.ascontiguousarray(snm)
        if sync_host:
            vals = np.zeros(ir.size)
            L.check(self._lib.cnmfe_get_spatial(self._h, _ptr(vals)))
            self.d2h_bytes += vals.nbytes
            vals = self._allreduce_sum(vals)
            A_new = sp.csc_matrix((vals, ir.copy(), jc.copy()), shape=INDc.shape)
            A_new.eliminate_zeros()
            if self.post_process_fn is not None:
                A_new = sp.csc_matrix(self.post_process_fn(A_new))
            self.A = self._freeze(A_new)
            if self.P.get("Ymean") is not None and self._is_ring():      # update_spatial_parallel.m:347-351
                self.b0_new = self.P["Ymean"] - (self.A @ self.C.mean(axis=1)).reshape(self.d1, self.d2, order="F")
            if self.world_size > 1 or self.post_process_fn is not None:
                self.push_neurons()
            else:
                self._mark("A", self.A)      # the device kept exactly these values

    def update_temporal_parallel(self, use_parallel=True, use_c_hat=True, sync_host=True):
        """update_temporal_parallel(obj, use_parallel, use_c_hat)."""
        self._push_options()
        L.check(self._lib.cnmfe_set_use_c_hat(self._h, int(bool(use_c_hat))))
        if sync_host:
            self.push_neurons()
            self.push_prev()
            self.push_ring()
        L.check(self._lib.cnmfe_update_temporal_patches(self._h))
        if self.world_size > 1:
            self._allreduce_merge_buffers()
        if self.world_size > 1 and self.options.get("shard_deconv"):
            # the final deconvTemporal is split over the ranks by trace; every rank needs the complete C for the
            # next update (one exchange now), C_raw / S / kernel_pars only when the host reads them (pull_temporal)
            K = self.A.shape[1]
            k0, k1 = trace_range(K, self.rank, self.world_size)
            L.check(self._lib.cnmfe_update_temporal_finish_part(self._h, k0, k1))
            self._allreduce_temporal_state(K, (0,))
            self._temporal_partial = {1, 2, 3}
        else:
            L.check(self._lib.cnmfe_update_temporal_finish(self._h))
            self._temporal_partial = set()
        if sync_host:
            self.pull_temporal()

    def pull_spatial(self):
        """obj.A from the device (values on the search pattern last given to update_spatial_parallel)."""
        jc, ir = self._ind_jc, self._ind_ir
        vals = np.zeros(ir.size)
        L.check(self._lib.cnmfe_get_spatial(self._h, _ptr(vals)))
        A_new = sp.csc_matrix((vals, ir.copy(), jc.copy()), shape=(self.d1 * self.d2, jc.size - 1))
        A_new.eliminate_zeros()
        self.A = self._freeze(A_new)
        self._mark("A", self.A)

    def exchange_spatial(self):
        """Multi-GPU: every rank solved the rows of its own patches; one all-reduce (disjoint supports => a gather) of
        the values on the search pattern gives every rank the full new A."""
        vals = np.zeros(self._ind_ir.size)
        L.check(self._lib.cnmfe_get_spatial(self._h, _ptr(vals)))
        vals = self._allreduce_sum(vals)
        L.check(self._lib.cnmfe_set_spatial(self._h, _ptr(vals)))

    # obj.C_raw and obj.S: K x T each.  After a temporal update they stay on the device and are copied out when they are
    # READ (the next update reads neither; the reference's own callers -- merging, deletion, saving -- do).  Reading is a local
    # device->host copy, never a collective: with several ranks pull_temporal() (collective) completes the device arrays first.
    @property
    def C_raw(self):
        if "C_raw" in self._lazy_temporal:
            self._fetch_lazy_temporal()
        return self._C_raw

    @C_raw.setter
    def C_raw(self, v):
        self._lazy_temporal.discard("C_raw")
        self._C_raw = v

    @property
    def S(self):
        if "S" in self._lazy_temporal:
            self._fetch_lazy_temporal()
        return self._S

    @S.setter
    def S(self, v):
        self._lazy_temporal.discard("S")
        self._S = v

    def _fetch_lazy_temporal(self):
        K = self.A.shape[1]
        if self._temporal_partial & {1, 2}:
            # reading an attribute must never start a collective (one rank reading would dead-lock the others)
            raise L.CnmfeError("obj.C_raw / obj.S: the rows deconvolved by the other ranks have not been exchanged yet -- "
                               "call pull_temporal() on ALL ranks after update_temporal_parallel(sync_host=False)")
        self._C_raw = self._S = None          # release the buffers of the previous iteration before taking new ones
        Cr = self._pool.take("C_raw", (K, self.T))
        S = self._pool.take("S", (K, self.T))
        L.check(self._lib.cnmfe_get_temporal(self._h, None, _ptr(Cr), _ptr(S), None, None))
        self.d2h_bytes += Cr.nbytes + S.nbytes
        self._C_raw, self._S = self._freeze(Cr), self._freeze(S)
        self._lazy_temporal.clear()

    def pull_temporal(self, lazy=True):
        """obj.C, obj.P.kernel_pars, obj.P.neuron_sn from the device now; obj.C_raw and obj.S when they are read (lazy=False:
        now).  Multi-rank with shard_deconv: a COLLECTIVE call (exchanges, on the devices, the rows of C_raw / S / kernel_pars
        that the other ranks deconvolved)."""
        K = self.A.shape[1]
        if self._temporal_partial:
            # device-side exchange (NVLink) of the rows the other ranks deconvolved; the copies to the host stay lazy
            self._allreduce_temporal_state(K, tuple(sorted(self._temporal_partial)))
            self._temporal_partial = set()
        C = self._pool.take("C", (K, self.T))
        kp = np.zeros((K, 2))
        nsn = np.zeros(K)
        L.check(self._lib.cnmfe_get_temporal(self._h, _ptr(C), None, None, _ptr(kp), _ptr(nsn)))
        self.d2h_bytes += C.nbytes + kp.nbytes + nsn.nbytes
        self.C = self._freeze(C)
        self._lazy_temporal = {"C_raw", "S"}
        if not lazy:
            self._fetch_lazy_temporal()
        self._mark("C", self.C)
        self.P["kernel_pars"], self.P["neuron_sn"] = kp, nsn
        if self.P.get("Ymean") is not None and self._is_ring():          # update_temporal_parallel.m:291-295
            self.b0_new = self.P["Ymean"] - (self.A @ self.C.mean(axis=1)).reshape(self.d1, self.d2, order="F")

    def _is_ring(self):
        return str(self.options["background_model"]).lower() == "ring"

    # aliases named by BASELINE.json north_star
    update_background = update_background_parallel
    update_spatial = update_spatial_parallel
    update_temporal = update_temporal_parallel

    # ------------------------------------------------------------------ multi-GPU exchange (torch.distributed)
    def _allreduce_owned_pixels(self, img):
        """Combine a per-pixel map whose entries are valid only on the pixels of the owned patches."""
        m = np.zeros_like(img)
        for i in self.owned_patches():
            p = self.patch_of(i)
            m[p[0] - 1:p[1], p[2] - 1:p[3]] = img[p[0] - 1:p[1], p[2] - 1:p[3]]
        return self._allreduce_sum(m.ravel()).reshape(img.shape)

    def _allreduce_sum(self, vec):
        if self.world_size == 1:
            return vec
        import torch
        import torch.distributed as dist
        dev = torch.device("cuda", self.device) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.from_numpy(vec).to(dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def _allreduce_merge_buffers(self):
        """C_raw = sum_p aa_p C_raw,p / sum_p aa_p across ranks (update_temporal_parallel.m:269-280): all-reduce of the
        K x T numerator and K denominator, in place on the device buffers (NCCL over NVLink) -- or staged through the
        host for the gloo backend used by the CPU-side tests of the host logic."""
        import torch
        import torch.distributed as dist
        K = self.A.shape[1]
        num = ctypes.c_void_p()
        den = ctypes.c_void_p()
        L.check(self._lib.cnmfe_temporal_merge_buffers(self._h, ctypes.byref(num), ctypes.byref(den)))
        L.check(self._lib.cnmfe_sync(self._h))

        class _Dev:
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = dict(shape=(n,), typestr="<f8", data=(ptr, False), version=3)

        tn = torch.as_tensor(_Dev(num.value, K * self.T), device=torch.device("cuda", self.device))
        td = torch.as_tensor(_Dev(den.value, K), device=torch.device("cuda", self.device))
        if dist.get_backend() == "nccl":
            dist.all_reduce(tn, op=dist.ReduceOp.SUM)
            dist.all_reduce(td, op=dist.ReduceOp.SUM)
            torch.cuda.synchronize(self.device)
        else:
            a, b = tn.cpu(), td.cpu()
            dist.all_reduce(a, op=dist.ReduceOp.SUM)
            dist.all_reduce(b, op=dist.ReduceOp.SUM)
            tn.copy_(a)
            td.copy_(b)
            torch.cuda.synchronize(self.device)

    def _allreduce_temporal_state(self, K, which=(0, 1, 2, 3)):
        """After cnmfe_update_temporal_finish_part: every rank holds its own rows of C, C_raw, S and of the per-trace
        outputs and zeros elsewhere; a SUM all-reduce (disjoint supports => a gather, x + 0 exact) completes them.
        which: indices into (C, C_raw, S, per-trace outputs)."""
        import torch
        import torch.distributed as dist
        ptrs = [ctypes.c_void_p() for _ in range(4)]
        L.check(self._lib.cnmfe_temporal_state_buffers(self._h, *[ctypes.byref(p) for p in ptrs]))
        L.check(self._lib.cnmfe_sync(self._h))

        class _Dev:
            def __init__(self, ptr, n):
                self.__cuda_array_interface__ = dict(shape=(n,), typestr="<f8", data=(ptr, False), version=3)

        for j, (p, n) in enumerate(zip(ptrs, (K * self.T, K * self.T, K * self.T, K * 6))):
            if j not in which:
                continue
            t = torch.as_tensor(_Dev(p.value, n), device=torch.device("cuda", self.device))
            if dist.get_backend() == "nccl":
                dist.all_reduce(t, op=dist.ReduceOp.SUM)
            else:
                a = t.cpu()
                dist.all_reduce(a, op=dist.ReduceOp.SUM)
                t.copy_(a)
        torch.cuda.synchronize(self.device)

    # ------------------------------------------------------------------ timing helpers (bench.py)
    def phase_ms(self):
        ms = (ctypes.c_float * 7)()
        L.check(self._lib.cnmfe_last_phase_ms(self._h, ms))
        return list(ms)

    def sync(self):
        L.check(self._lib.cnmfe_sync(self._h))
