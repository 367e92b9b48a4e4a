"""Host-side mirror of the reference's per-trace interface: deconvolveCa, GetSn, HALS_temporal (from projections).

Argument names/meaning follow OASIS_matlab/deconvolveCa.m:208-355 (the name/value parser) and
ca_source_extraction/utilities/HALS_temporal.m:1-19.  All compute runs in libcnmfe_b200.so on a CUDA device."""
import ctypes
import numpy as np
from . import _lib as L

_TYPES = {"ar1": 1, "ar2": 2}
_METHODS = {"foopsi": 0, "constrained": 1, "thresholded": 2}


def make_deconv_opts(options=None, **kw):
    """Struct-then-name/value merge as deconvolveCa.m:233-246.  Returns (DeconvOpts, pars or None, sn or None)."""
    o = {}
    if options:
        o.update(options)
    o.update(kw)
    d = L.DeconvOpts()
    L.lib().cnmfe_deconv_defaults(ctypes.byref(d))
    pars = o.pop("pars", None)
    sn = o.pop("sn", None)
    for k, v in o.items():
        if k == "type":
            if v not in _TYPES:
                raise ValueError("type %r not supported (ar1, ar2)" % (v,))
            d.type = _TYPES[v]
        elif k == "method":
            if v not in _METHODS:
                raise ValueError("method %r not supported (foopsi, constrained, thresholded)" % (v,))
            d.method = _METHODS[v]
        elif k in ("lambda", "lam"):
            d.lam = float(v)
        elif k == "tau_range":
            if v is not None and len(v) == 2:
                d.has_tau_range = 1
                d.tau_range[0], d.tau_range[1] = float(v[0]), float(v[1])
        elif k in ("optimize_b", "optimize_pars"):
            setattr(d, k, int(bool(v)))
        elif k == "maxIter":
            d.maxIter = int(v)
        elif k in ("smin", "b", "max_tau", "thresh_factor", "p_noise"):
            setattr(d, k, float(v))
        elif k in ("optimize_smin", "window", "shift", "extra_params", "remove_large_residuals"):
            if k == "remove_large_residuals" and v:
                raise ValueError("remove_large_residuals needs `fastsmooth`, absent from the reference tree")
        else:
            raise ValueError("unknown deconvolveCa option %r" % (k,))
    return d, pars, sn


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def deconvolveCa_batch(Y, options=None, device=0, **kw):
    """Y: (N, T).  Returns dict(c, s (N,T), b, sn, smin, lam (N,), pars (N,2))."""
    Y = np.ascontiguousarray(Y, dtype=np.float64)
    if Y.ndim == 1:
        Y = Y[None, :]
    N, T = Y.shape
    d, pars, sn = make_deconv_opts(options, **kw)
    sn_in = None if sn is None else np.ascontiguousarray(np.broadcast_to(np.asarray(sn, dtype=np.float64), (N,)))
    pars_in = None
    if pars is not None and np.size(pars) > 0:
        p = np.atleast_2d(np.asarray(pars, dtype=np.float64))
        if p.shape[0] == 1 and N > 1:
            p = np.repeat(p, N, axis=0)
        pars_in = np.zeros((N, 2))
        pars_in[:, :p.shape[1]] = p
    c = np.empty((N, T)); s = np.empty((N, T))
    b = np.empty(N); po = np.empty((N, 2)); sno = np.empty(N); smin = np.empty(N); lam = np.empty(N)
    L.check(L.lib().cnmfe_deconvolve(_ptr(Y), T, N, ctypes.byref(d), _ptr(sn_in), _ptr(pars_in), _ptr(c), _ptr(s),
                                     _ptr(b), _ptr(po), _ptr(sno), _ptr(smin), _ptr(lam), device))
    return dict(c=c, s=s, b=b, pars=po, sn=sno, smin=smin, lam=lam)


def deconvolveCa(y, options=None, device=0, **kw):
    """[c, s, options] = deconvolveCa(y, varargin)  (single trace)."""
    r = deconvolveCa_batch(np.asarray(y, dtype=np.float64).ravel()[None, :], options, device, **kw)
    d, _, _ = make_deconv_opts(options, **kw)
    opts = dict(options or {})
    opts.update(kw)
    opts.update(b=float(r["b"][0]), sn=float(r["sn"][0]), smin=float(r["smin"][0]), lam=float(r["lam"][0]),
                pars=r["pars"][0, :d.type].copy())
    return r["c"][0], r["s"][0], opts


def GetSn(Y, device=0):
    """sn = GetSn(Y): Y (N,T) or (T,)."""
    Y = np.asarray(Y, dtype=np.float64)
    scalar = Y.ndim == 1
    Y2 = np.ascontiguousarray(Y[None, :] if scalar else Y)
    N, T = Y2.shape
    sn = np.empty(N)
    L.check(L.lib().cnmfe_get_sn(_ptr(Y2), T, N, _ptr(sn), device))
    return float(sn[0]) if scalar else sn


def HALS_temporal_uv(U, V, C, maxIter=1, deconv_options=None, device=0):
    """HALS_temporal given U = A'*Y (K,T), V = A'*A (K,K).  Returns C, C_raw, results_deconv, S."""
    U = np.asfortranarray(U, dtype=np.float64)
    V = np.asfortranarray(V, dtype=np.float64)
    Cw = np.asfortranarray(np.array(C, dtype=np.float64))
    K, T = U.shape
    C_raw = np.zeros((K, T), order="F"); S = np.zeros((K, T), order="F")
    sn = np.zeros(K); kp = np.zeros((K, 2))
    dptr = None
    if deconv_options:
        d, _, _ = make_deconv_opts(deconv_options)
        dptr = ctypes.byref(d)
    L.check(L.lib().cnmfe_hals_temporal_uv(_ptr(U), _ptr(V), K, T, _ptr(Cw), int(maxIter), dptr, _ptr(C_raw),
                                           _ptr(S), _ptr(sn), _ptr(kp), device))
    res = dict(sn=sn, kernel_pars=kp) if deconv_options else None
    return np.ascontiguousarray(Cw), np.ascontiguousarray(C_raw), res, np.ascontiguousarray(S)
