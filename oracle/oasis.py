"""
oracle/oasis.py -- TEST INFRASTRUCTURE ONLY (parity oracle; the product never imports this).

Float64 CPU restatement of the reference's OASIS deconvolution stack, following (file:line under
/root/reference):

  GetSn                   OASIS_matlab/functions/GetSn.m:1-46           (pwelch restated, see `pwelch_psd`)
  estimate_time_constant  OASIS_matlab/functions/estimate_time_constant.m:1-67
  choose_smin / max_ht / ar2exp / exp2ar   OASIS_matlab/functions/{choose_smin,max_ht,ar2exp,exp2ar}.m
  fminbnd                 MathWorks `fminbnd` (closed source): restated from the published
                          Forsythe-Malcolm-Moler golden-section/parabolic routine, TolX=1e-4, MaxFunEvals=500
  quantile                MathWorks `quantile`: linear interpolation at plotting positions (i-0.5)/n
  oasisAR1 / oasisAR2     OASIS_matlab/packages/oasis/oasisAR1.m, oasisAR2.m        (loops in oasis_core.c)
  foopsi_oasisAR1 (+update_g)   OASIS_matlab/packages/oasis/foopsi_oasisAR1.m:1-179
  foopsi_oasisAR2         OASIS_matlab/packages/oasis/foopsi_oasisAR2.m:1-113 (optimize_* unreachable from deconvolveCa)
  constrained_oasisAR1    OASIS_matlab/packages/oasis/constrained_oasisAR1.m:1-260
  thresholded_oasisAR1    OASIS_matlab/packages/oasis/thresholded_oasisAR1.m:1-277
  deconvolveCa            OASIS_matlab/deconvolveCa.m:1-356

PARITY UNPINNED: the reference ships no assertions/golden vectors and MATLAB/Octave are not
installed here, so this restatement is pinned only by known-answer tests (tests/test_oracle_oasis.py):
independent NNLS solve of the FOOPSI problem, RSS = sn^2 T for the constrained form, PAV invariants,
scipy.signal.welch cross-check of the PSD.  Deviation from the reference, by necessity: the `randn` root jitter of
estimate_time_constant.m:62-64 is replaced by 0 (deterministic).
"""
import ctypes
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_core.so")


def build_core(force=False):
    src = os.path.join(_HERE, "oasis_core.c")
    if force or (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", _SO, src, "-lm"])
    return _SO


_lib = None


def _core():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build_core())
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        _lib.oasis_ar1_pools.restype = ctypes.c_int
        _lib.oasis_ar1_pools.argtypes = [dp, dp, ip, ip, ctypes.c_int, ctypes.c_double, ctypes.c_double]
        _lib.oasis_ar1_solution.restype = None
        _lib.oasis_ar1_solution.argtypes = [dp, dp, ip, ip, ctypes.c_int, ctypes.c_double, ctypes.c_int, dp, dp]
        _lib.rss_g_ar1.restype = ctypes.c_double
        _lib.rss_g_ar1.argtypes = [dp, ctypes.c_int, ip, ip, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                   ctypes.c_int, dp, dp, dp]
        _lib.rebuild_pools_ar1.restype = None
        _lib.rebuild_pools_ar1.argtypes = [dp, ip, ip, ctypes.c_int, ctypes.c_double, ctypes.c_double,
                                           ctypes.c_int, dp, dp, dp]
        _lib.oasis_ar2.restype = ctypes.c_int
        _lib.oasis_ar2.argtypes = [dp, ctypes.c_int] + [ctypes.c_double] * 6 + [dp, dp, dp, dp, ip, ip]
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int))


# ----------------------------------------------------------------------------- toolbox restatements
def quantile(x, p):
    """MATLAB quantile(x,p): linear interpolation, sorted x(i) at probability (i-0.5)/n (== numpy 'hazen')."""
    return float(np.quantile(np.asarray(x, dtype=np.float64), p, method="hazen"))


def fminbnd(f, ax, bx, tol=1e-4, maxfun=500, maxiter=500):
    """MATLAB fminbnd (FMM `fmin`), default TolX 1e-4.  Returns (xf, last_x_evaluated)."""
    seps = np.sqrt(np.finfo(np.float64).eps)
    c = 0.5 * (3.0 - np.sqrt(5.0))
    a, b = float(ax), float(bx)
    v = a + c * (b - a)
    w = v
    xf = v
    d = 0.0
    e = 0.0
    x = xf
    fx = f(x)
    funccount = 1
    it = 0
    fv = fx
    fw = fx
    xm = 0.5 * (a + b)
    tol1 = seps * abs(xf) + tol / 3.0
    tol2 = 2.0 * tol1
    while abs(xf - xm) > (tol2 - 0.5 * (b - a)):
        gs = True
        if abs(e) > tol1:
            gs = False
            r = (xf - w) * (fx - fv)
            q = (xf - v) * (fx - fw)
            p = (xf - v) * q - (xf - w) * r
            q = 2.0 * (q - r)
            if q > 0.0:
                p = -p
            q = abs(q)
            r = e
            e = d
            if (abs(p) < abs(0.5 * q * r)) and (p > q * (a - xf)) and (p < q * (b - xf)):
                d = p / q
                x = xf + d
                if ((x - a) < tol2) or ((b - x) < tol2):
                    si = np.sign(xm - xf) + ((xm - xf) == 0)
                    d = tol1 * si
            else:
                gs = True
        if gs:
            e = (a - xf) if xf >= xm else (b - xf)
            d = c * e
        si = np.sign(d) + (d == 0)
        x = xf + si * max(abs(d), tol1)
        fu = f(x)
        funccount += 1
        it += 1
        if fu <= fx:
            if x >= xf:
                a = xf
            else:
                b = xf
            v, fv = w, fw
            w, fw = xf, fx
            xf, fx = x, fu
        else:
            if x < xf:
                a = x
            else:
                b = x
            if (fu <= fw) or (w == xf):
                v, fv = w, fw
                w, fw = x, fu
            elif (fu <= fv) or (v == xf) or (v == w):
                v, fv = x, fu
        xm = 0.5 * (a + b)
        tol1 = seps * abs(xf) + tol / 3.0
        tol2 = 2.0 * tol1
        if funccount >= maxfun or it >= maxiter:
            break
    return xf, x


def pwelch_psd(x):
    """pwelch(x,[],[],[],1) restated (MathWorks Signal toolbox, closed source; documented defaults):
    Hamming window L=fix(N/4.5), noverlap=fix(L/2), nfft=max(256,2^nextpow2(L)), one-sided PSD, fs=1."""
    x = np.asarray(x, dtype=np.float64).ravel()
    N = x.size
    L = int(np.fix(N / 4.5))
    noverlap = int(np.fix(0.5 * L))
    k = int(np.fix((N - noverlap) / (L - noverlap)))
    nfft = max(256, 1 << int(np.ceil(np.log2(L))))
    n = np.arange(L)
    win = 0.54 - 0.46 * np.cos(2 * np.pi * n / (L - 1))
    U = np.sum(win * win)
    step = L - noverlap
    P = np.zeros(nfft // 2 + 1)
    for i in range(k):
        seg = x[i * step:i * step + L] * win
        X = np.fft.rfft(seg, nfft)
        P += (X.real ** 2 + X.imag ** 2)
    P /= (k * U)
    P[1:-1] *= 2.0
    ff = np.arange(nfft // 2 + 1) / nfft
    return P, ff


def GetSn(Y, range_ff=(0.25, 0.5), method="logmexp"):
    """GetSn.m:18-46.  Y: (N,T) or (T,) -> sn (N,) or float."""
    Y = np.asarray(Y, dtype=np.float64)
    if Y.ndim == 1 or 1 in Y.shape:
        rows = [Y.ravel()]
        scalar = True
    else:
        rows = list(Y)
        scalar = False
    out = np.zeros(len(rows))
    for i, x in enumerate(rows):
        P, ff = pwelch_psd(x)
        ind = (ff >= range_ff[0]) & (ff <= range_ff[1])
        if method == "mean":
            out[i] = np.sqrt(np.mean(P[ind] / 2))
        elif method == "median":
            out[i] = np.sqrt(np.median(P[ind] / 2))
        else:
            out[i] = np.sqrt(np.exp(np.mean(np.log(P[ind] / 2))))
    return float(out[0]) if scalar else out


def xcov_biased(y, lags):
    y = np.asarray(y, dtype=np.float64).ravel()
    yn = y - y.mean()
    xc = np.array([np.dot(yn[k:], yn[:yn.size - k]) for k in range(lags + 1)]) / y.size
    return np.concatenate([xc[:0:-1], xc])


def estimate_time_constant(y, p=2, sn=None, lags=5, fudge_factor=1.0):
    """estimate_time_constant.m:21-67 (randn jitter -> 0)."""
    y = np.asarray(y, dtype=np.float64).ravel()
    if sn is None:
        sn = GetSn(y)
    lags = lags + p
    xc = xcov_biased(y, lags)
    col = xc[lags + np.arange(0, lags)]
    row = xc[lags + np.arange(0, p)]
    A = np.empty((lags, p))
    for i in range(lags):
        for j in range(p):
            A[i, j] = col[i - j] if i >= j else row[j - i]
    A = A - sn ** 2 * np.eye(lags, p)
    g = np.linalg.pinv(A) @ xc[lags + 1:]
    while np.max(np.abs(np.roots(np.concatenate([[1.0], -g.ravel()])))) > 1 and p < 5:
        p = p + 1
        g = np.atleast_1d(estimate_time_constant(y, p, sn, lags))
    if p == 5:
        g = np.array([0.0])
    rg = np.roots(np.concatenate([[1.0], -np.atleast_1d(g).ravel()]))
    if np.iscomplexobj(rg) and np.any(rg.imag != 0):
        rg = rg.real + 0.0
    rg = np.real(rg).astype(np.float64)
    rg[rg > 1] = 0.95
    rg[rg < 0] = 0.15
    pg = np.poly(fudge_factor * rg)
    return -pg[1:]


def ar2exp(g):
    g = np.atleast_1d(np.asarray(g, dtype=np.float64))
    if g.size == 1:
        g = np.array([g[0], 0.0])
    temp = np.roots([1.0, -g[0], -g[1]])
    d, r = np.max(temp.real), np.min(temp.real)
    return np.array([-1 / np.log(d), -1 / np.log(r)])


def exp2ar(tau_dr):
    d = np.exp(-1 / tau_dr[0])
    r = np.exp(-1 / tau_dr[1])
    return np.array([d + r, -d * r])


def max_ht(pars):
    """max_ht.m:11-29 (single-argument form)."""
    pars = np.atleast_1d(pars)
    if pars.size == 1:
        return 1.0
    taus = ar2exp(pars)
    t = np.arange(1, int(np.ceil(taus[0] * 2)) + 1, dtype=np.float64)
    d = np.exp(-1.0 / taus[0])
    r = np.exp(-1.0 / taus[1])
    ht = (np.exp(np.log(d) * t) - np.exp(np.log(r) * t)) / (d - r)
    return float(np.max(ht))


def choose_smin(kernel, sn, prob=0.99999):
    """choose_smin.m:28-42."""
    from scipy.stats import norm
    kernel = np.atleast_1d(np.asarray(kernel, dtype=np.float64))
    if kernel.size <= 2:
        a = np.concatenate([[1.0], -kernel])
        h = np.zeros(1000)
        for n in range(1000):
            acc = 1.0 if n == 0 else 0.0
            for j in range(1, a.size):
                if n - j >= 0:
                    acc -= a[j] * h[n - j]
            h[n] = acc
        kernel = h
    return float(sn / np.linalg.norm(kernel) * norm.ppf(prob))


# ----------------------------------------------------------------------------- PAV cores
class Pools:
    """active_set of the reference: columns (v, w, t, l); t is 0-based here."""
    __slots__ = ("v", "w", "t", "l")

    def __init__(self, v, w, t, l):
        self.v = np.ascontiguousarray(v, dtype=np.float64)
        self.w = np.ascontiguousarray(w, dtype=np.float64)
        self.t = np.ascontiguousarray(t, dtype=np.int32)
        self.l = np.ascontiguousarray(l, dtype=np.int32)

    def copy(self):
        return Pools(self.v.copy(), self.w.copy(), self.t.copy(), self.l.copy())

    def __len__(self):
        return self.v.size


def oasisAR1(y, g, lam=0.0, smin=0.0, active_set=None):
    """oasisAR1.m:30-109.  y may be None/empty when warm-starting from `active_set`."""
    lib = _core()
    if y is not None and len(y) == 0:
        y = None
    if y is None:
        T = int(np.sum(active_set.l))
    else:
        y = np.asarray(y, dtype=np.float64).ravel()
        T = y.size
    g = np.atleast_1d(g)
    if g.size > 1:
        return np.zeros(T), np.zeros(T), Pools([], [], [], [])
    g = float(g[0])
    lam = 0.0 if lam is None else float(lam)
    smin = 0.0 if smin is None else float(smin)
    if active_set is None or len(active_set) == 0:
        v = y - lam * (1 - g)
        v[-1] = y[-1] - lam
        P = Pools(v, np.ones(T), np.arange(T), np.ones(T, dtype=np.int32))
    else:
        P = active_set.copy()
    n = lib.oasis_ar1_pools(_dp(P.v), _dp(P.w), _ip(P.t), _ip(P.l), len(P), g, smin)
    P = Pools(P.v[:n], P.w[:n], P.t[:n], P.l[:n])
    c = np.zeros(T)
    s = np.zeros(T)
    lib.oasis_ar1_solution(_dp(P.v), _dp(P.w), _ip(P.t), _ip(P.l), n, g, T, _dp(c), _dp(s))
    return c, s, P


def oasisAR1_py(y, g, lam=0.0, smin=0.0):
    """Pure-Python stack form of oasisAR1.m:57-98 for small cases (cross-checks oasis_core.c)."""
    y = np.asarray(y, dtype=np.float64)
    T = y.size
    v = list(y - lam * (1 - g))
    v[-1] = y[-1] - lam
    pools = [[v[0], 1.0, 0, 1]]
    i = 1
    while i < T:
        nxt = [v[i], 1.0, i, 1]
        i += 1
        cur = pools[-1]
        if nxt[0] / nxt[1] >= cur[0] / cur[1] * g ** cur[3] + smin:
            pools.append(nxt)
            continue
        cur[0] += nxt[0] * g ** cur[3]
        cur[1] += nxt[1] * g ** (2 * cur[3])
        cur[3] += nxt[3]
        while len(pools) > 1:
            pr = pools[-2]
            cur = pools[-1]
            if cur[0] / cur[1] < max(0.0, pr[0] / pr[1] * g ** pr[3]) + smin:
                pr[0] += cur[0] * g ** pr[3]
                pr[1] += cur[1] * g ** (2 * pr[3])
                pr[3] += cur[3]
                pools.pop()
            else:
                break
    c = np.zeros(T)
    s = np.zeros(T)
    for (pv, pw, pt, pl) in pools:
        c[pt:pt + pl] = max(0.0, pv / pw) * g ** np.arange(pl)
    for (pv, pw, pt, pl) in pools[1:]:
        s[pt] = c[pt] - g * c[pt - 1]
    return c, s, pools


def oasisAR2(y, g, lam=0.0, smin=0.0):
    """oasisAR2.m:31-156 (cold start, T_over_ISI=1, no jitter)."""
    lib = _core()
    y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).ravel())
    T = y.size
    g1, g2 = float(g[0]), float(g[1])
    lam = 0.0 if lam is None else float(lam)
    smin = 0.0 if smin is None else float(smin)
    temp = np.roots([1.0, -g1, -g2])
    if np.any(np.abs(temp.imag) > 0):
        raise ValueError("oasisAR2 oracle: complex AR(2) roots are not restated")
    d, r = float(np.max(temp.real)), float(np.min(temp.real))
    c = np.zeros(T)
    s = np.zeros(T)
    pv = np.zeros(T)
    pw = np.zeros(T)
    pt = np.zeros(T, dtype=np.int32)
    pl = np.zeros(T, dtype=np.int32)
    n = lib.oasis_ar2(_dp(y), T, g1, g2, d, r, lam, smin, _dp(c), _dp(s), _dp(pv), _dp(pw), _ip(pt), _ip(pl))
    return c, s, Pools(pv[:n], pw[:n], pt[:n], pl[:n])


# ----------------------------------------------------------------------------- update_g (shared by AR1 drivers)
def _update_g_ar1(y, active_set, lam, smin, g_range):
    """update_g of foopsi_oasisAR1.m:124-179 (== constrained_:202-259 with smin=[], thresholded_:233-277 with lam=0)."""
    lib = _core()
    y = np.ascontiguousarray(np.asarray(y, dtype=np.float64).ravel())
    T = y.size
    n = len(active_set)
    maxl = int(np.max(active_set.l))
    h = np.zeros(maxl + 1)
    hh = np.zeros(maxl + 1)
    cbuf = np.zeros(T)
    t_, l_ = active_set.t, active_set.l

    def rss(gv):
        return lib.rss_g_ar1(_dp(y), T, _ip(t_), _ip(l_), n, float(gv), float(lam), maxl, _dp(h), _dp(hh), _dp(cbuf))

    g, _ = fminbnd(rss, g_range[0], g_range[1])
    P = active_set.copy()
    lib.rebuild_pools_ar1(_dp(y), _ip(P.t), _ip(P.l), n, float(g), float(lam), maxl, _dp(hh), _dp(P.v), _dp(P.w))
    c, s, P = oasisAR1(y, g, lam, smin, P)
    return c, P, g, s


def foopsi_oasisAR1(y, g=None, lam=0.0, smin=0.0, optimize_b=False, optimize_g=False, decimate=None,
                    maxIter=10, tau_range=None, gmax=None):
    """foopsi_oasisAR1.m:36-122."""
    y = np.asarray(y, dtype=np.float64).ravel()
    if g is None or np.size(g) == 0:
        g = estimate_time_constant(y, 1)
    g = float(np.atleast_1d(g)[0])
    lam = 0.0 if lam is None else lam
    if smin is None:
        smin = 0.0
    elif smin < 0:
        smin = abs(smin) * GetSn(y)
    if tau_range is None or np.size(tau_range) == 0:
        g_range = (0.0, 1.0)
    else:
        g_range = tuple(np.exp(-1.0 / np.asarray(tau_range, dtype=np.float64)))
        g = min(max(g, g_range[0]), g_range[1])
    if not optimize_b:
        b = 0.0
        solution, spks, active_set = oasisAR1(y, g, lam, smin)
        if optimize_g:
            solution, active_set, g, spks = _update_g_ar1(y, active_set, lam, smin, g_range)
    else:
        b = quantile(y, 0.15)
        solution, spks, active_set = oasisAR1(y - b, g, lam, smin)
        for _ in range(int(maxIter)):
            b = float(np.mean(y - solution))
            if optimize_g:
                if len(active_set) == 0:
                    break
                g0 = g
                if gmax is not None and g > gmax:
                    g = float(estimate_time_constant(y, 1)[0])
                    solution, spks, active_set = oasisAR1(y - b, g, lam, smin)
                    break
                solution, active_set, g, spks = _update_g_ar1(y - b, active_set, lam, smin, g_range)
                if abs(g - g0) / g0 < 1e-3:
                    optimize_g = False
            else:
                break
    return solution, spks, b, g, active_set


def foopsi_oasisAR2(y, g, lam=0.0, smin=0.0):
    """foopsi_oasisAR2.m:36-113 with optimize_b=optimize_g=false (the only form deconvolveCa.m:128-129 reaches)."""
    solution, spks, active_set = oasisAR2(y, g, lam, smin)
    return solution, spks, 0.0, np.asarray(g, dtype=np.float64), active_set


def constrained_oasisAR1(y, g=None, sn=None, optimize_b=False, optimize_g=False, decimate=None, maxIter=10,
                         tau_range=None):
    """constrained_oasisAR1.m:37-199."""
    y = np.asarray(y, dtype=np.float64).ravel()
    T = y.size
    if g is None or np.size(g) == 0:
        g = estimate_time_constant(y, 1)
    g = float(np.atleast_1d(g)[0])
    if sn is None:
        sn = GetSn(y)
    if tau_range is None or np.size(tau_range) == 0:
        g_range = (0.0, 1.0)
    else:
        g_range = tuple(np.exp(-1.0 / np.asarray(tau_range, dtype=np.float64)))
        g = min(max(g, g_range[0]), g_range[1])
    thresh = sn * sn * T
    lam = 0.0
    tol = 1e-4
    st = {}

    def update_phi(res, RSS):
        """constrained_oasisAR1.m:151-187; returns False if dphi is complex (flag_phi)."""
        aset = st["aset"]
        n = len(aset)
        zeta = np.zeros(T)
        maxl = int(np.max(aset.l))
        h = g ** np.arange(0, maxl + 1)
        for ii in range(n):
            ti, li = int(aset.t[ii]), int(aset.l[ii])
            if ii < n - 1:
                zeta[ti:ti + li] = (1 - g ** li) / aset.w[ii] * h[:li]
            else:
                zeta[ti:ti + li] = 1 / aset.w[ii] * h[:li]
        if optimize_b:
            zeta = zeta - np.mean(zeta)
            tmp_res = res - np.mean(res)
            aa = zeta @ zeta
            bb = tmp_res @ zeta
            cc = tmp_res @ tmp_res - thresh
        else:
            aa = zeta @ zeta
            bb = res @ zeta
            cc = RSS - thresh
        disc = bb ** 2 - aa * cc
        if disc < 0:
            dphi = complex(-bb, np.sqrt(-disc)) / aa
            if dphi.imag > 1e-9:
                st["dphi"] = dphi
                return False
            dphi = dphi.real
        else:
            dphi = (-bb + np.sqrt(disc)) / aa
        st["dphi"] = dphi
        aset = aset.copy()
        aset.v = aset.v - dphi * (1 - g ** aset.l.astype(np.float64))
        sol, spk, aset = oasisAR1(None, g, st["lam"], None, aset)
        st["aset"], st["solution"], st["spks"] = aset, sol, spk
        return True

    def update_lam_b():
        """constrained_oasisAR1.m:189-199."""
        db = float(np.mean(y - st["solution"])) - st["b"]
        st["b"] = st["b"] + db
        dlam = -db / (1 - g)
        st["lam"] = max(0.0, st["lam"] + dlam)
        aset = st["aset"]
        aset.v[-1] = aset.v[-1] - st["lam"] * g ** int(aset.l[-1])
        ti, li = int(aset.t[-1]), int(aset.l[-1])
        st["solution"][ti:ti + li] = max(0.0, aset.v[-1] / aset.w[-1]) * g ** np.arange(li)

    st["lam"] = lam
    g_converged = False
    if not optimize_b:
        st["b"] = 0.0
        sol, spk, aset = oasisAR1(y, g, lam)
        st["aset"], st["solution"], st["spks"] = aset, sol, spk
        for _ in range(int(maxIter)):
            if optimize_g and not g_converged:
                g0 = g
                sol, aset, g, spk = _update_g_ar1(y, st["aset"], st["lam"], None, g_range)
                st["aset"], st["solution"], st["spks"] = aset, sol, spk
                if abs(g - g0) / g0 < 1e-3:
                    g_converged = True
            res = y - st["solution"]
            RSS = float(res @ res)
            if RSS > thresh:
                break
            else:
                update_phi(res, RSS)
                st["lam"] = st["lam"] + np.real(st["dphi"])
    else:
        st["b"] = quantile(y, 0.15)
        sol, spk, aset = oasisAR1(y - st["b"], g, lam)
        st["aset"], st["solution"], st["spks"] = aset, sol, spk
        update_lam_b()
        g_converged = False
        for _ in range(int(maxIter)):
            res = y - st["solution"] - st["b"]
            RSS = float(res @ res)
            if abs(RSS - thresh) < tol or np.sum(st["solution"]) < 1e-9:
                break
            else:
                update_phi(res, RSS)
                update_lam_b()
                if optimize_g and not g_converged:
                    g0 = g
                    sol, aset, g, spk = _update_g_ar1(y - st["b"], st["aset"], st["lam"], None, g_range)
                    st["aset"], st["solution"], st["spks"] = aset, sol, spk
                    if abs(g - g0) / g0 < 1e-4:
                        g_converged = True
    return st["solution"], st["spks"], st["b"], g, st["lam"], st["aset"]


def hist_centers(y, centers):
    """nums = hist(y, centers) (MATLAB hist.m with a vector of bin CENTRES): edges half-way between the centres, the first and
    the last bin extended to min(y) / max(y), bins of the form (edge_k, edge_k+1] (hist.m shifts the edges by eps for that)."""
    y = np.asarray(y, dtype=np.float64).ravel()
    xx = np.asarray(centers, dtype=np.float64).ravel()
    binwidth = np.concatenate([np.diff(xx), [0.0]])
    edges = np.concatenate([[xx[0] - binwidth[0] / 2], xx + binwidth / 2])
    edges[0] = min(edges[0], y.min())
    edges[-1] = max(edges[-1], y.max())
    bins = edges + np.spacing(edges)
    k = np.searchsorted(bins, y, side="right")          # number of shifted edges <= y
    k = np.clip(k, 1, xx.size) - 1                       # first / last histc bins are merged into their neighbours
    return np.bincount(k, minlength=xx.size).astype(np.float64)


def _solve_small(M, b):
    """M \\ b for a small square system: Gaussian elimination with partial pivoting (what mldivide does for a general square
    matrix), written out so that the CUDA path performs the same operations in the same order."""
    M = np.array(M, dtype=np.float64)
    b = np.array(b, dtype=np.float64)
    n = b.size
    for c in range(n):
        piv = c + int(np.argmax(np.abs(M[c:, c])))
        if piv != c:
            M[[c, piv]] = M[[piv, c]]
            b[[c, piv]] = b[[piv, c]]
        for r in range(c + 1, n):
            f = M[r, c] / M[c, c]
            M[r, c:] = M[r, c:] - f * M[c, c:]
            b[r] = b[r] - f * b[c]
    x = np.zeros(n)
    for r in range(n - 1, -1, -1):
        x[r] = (b[r] - M[r, r + 1:] @ x[r + 1:]) / M[r, r]
    return x


def fit_gauss1(x, y, thr=0.1, maxIter=5, mu_fix=False):
    """functions/fit_gauss1.m:1-92 (Guo 2011, iteratively re-weighted log-parabola fit)."""
    x = np.asarray(x, dtype=np.float64).ravel()
    y = np.asarray(y, dtype=np.float64).ravel()
    ind = y > y.max() * thr
    x, y = x[ind], y[ind]
    x2, x3, x4 = x ** 2, x ** 3, x ** 4
    y2 = y ** 2
    logy = np.log(y)
    y2logy = y2 * logy
    p = None
    for _ in range(int(maxIter)):
        if mu_fix:
            M = [[y2.sum(), x2 @ y2], [x2 @ y2, x4 @ y2]]
            p = _solve_small(M, [y2logy.sum(), x2 @ y2logy])
            logy = p[0] + p[1] * x2
        else:
            M = [[y2.sum(), x @ y2, x2 @ y2], [x @ y2, x2 @ y2, x3 @ y2], [x2 @ y2, x3 @ y2, x4 @ y2]]
            p = _solve_small(M, [y2logy.sum(), x @ y2logy, x2 @ y2logy])
            logy = p[0] + p[1] * x + p[2] * x2
        y = np.exp(logy)
        y2 = y ** 2
        y2logy = y2 * logy
    if mu_fix:
        return 0.0, float(np.sqrt(-0.5 / p[1])), float(np.exp(p[0]))
    return float(-p[1] / 2 / p[2]), float(abs(np.sqrt(-0.5 / p[2] + 0j))), float(np.exp(p[0] - 0.25 * p[1] ** 2 / p[2]))


def estimate_baseline_noise(y, bmin=-np.inf):
    """functions/estimate_baseline_noise.m:1-40: histogram on a grid from the deciles, Gaussian fit of its peak.  The grid
    `temp(1):dbin:temp(end)` is restated as temp(1) + k*dbin (MATLAB's colon operator fills long ranges from both ends; the
    two agree to rounding)."""
    y = np.asarray(y, dtype=np.float64).ravel()
    temp = np.array([quantile(y, q) for q in np.arange(0, 11) / 10.0])
    dbin = max(np.min(np.diff(temp)) / 3, (temp.max() - temp.min()) / 1000)
    nb = int(np.floor((temp[-1] - temp[0]) / dbin + 1e-10)) + 1 if dbin > 0 else 0
    if nb <= 0:
        return float(np.mean(y)), 0.0
    bins = temp[0] + dbin * np.arange(nb)
    nums = hist_centers(y, bins)
    b, sn, _ = fit_gauss1(bins, nums, 0.3, 3)
    if b < bmin:
        b = bmin
        _, sn, _ = fit_gauss1(bins - bmin, nums, 0.3, 3, False)
    return b, sn


def thresholded_oasisAR1(y, g=None, sn=None, optimize_b=False, optimize_g=False, decimate=None, maxIter=10,
                         thresh_factor=1.0, p_noise=0.9999, tau_range=None):
    """thresholded_oasisAR1.m:40-184, both branches (optimize_b: baseline from estimate_baseline_noise, then
    b = mean(y - solution) after every update_smin, :141-181)."""
    y = np.asarray(y, dtype=np.float64).ravel()
    T = y.size
    g = float(np.atleast_1d(g)[0])
    smin = choose_smin(g, sn, p_noise)
    thresh = thresh_factor * sn * sn * T
    if tau_range is None or np.size(tau_range) == 0:
        g_range = (0.0, 1.0)
    else:
        g_range = tuple(np.exp(-1.0 / np.asarray(tau_range, dtype=np.float64)))
        g = min(max(g, g_range[0]), g_range[1])
    g_converged = False
    tol = 1e-4
    b = estimate_baseline_noise(y)[0] if optimize_b else 0.0
    solution, spks, aset = oasisAR1(y - b, g, None, smin)
    res = y - solution - b
    RSS0 = float(res @ res)

    def update_smin(yy, smin, solution, spks, aset, thr):
        """thresholded_oasisAR1.m:186-213."""
        n = len(aset)
        s_max = float(np.max(aset.v / aset.w))
        sv = np.linspace(smin, s_max, min(9, n))
        ind_start, ind_end = 1, sv.size
        while (ind_end - ind_start) > 1:
            ind = (ind_start + ind_end) // 2
            tmp_smin = sv[ind - 1]
            ts, tk, ta = oasisAR1(None, g, None, tmp_smin, aset)
            sq = float(np.linalg.norm(yy - ts))
            if sq < thr:
                solution, spks, aset, smin = ts, tk, ta, tmp_smin
                ind_start = ind
            elif sq > thr:
                ind_end = ind
            else:
                break
        return smin, solution, spks, aset

    for _ in range(int(maxIter)):
        if len(aset) == 0:
            break
        if optimize_g and not g_converged:
            g0 = g
            solution, aset, g, spks = _update_g_ar1(y - b, aset, 0.0, smin, g_range)
            if abs(g - g0) / g0 < 1e-4:
                g_converged = True
                if optimize_b:                       # :156 -- only the optimize_b branch re-runs a cold oasisAR1
                    solution, spks, aset = oasisAR1(y - b, g, None, smin)
        res = y - solution - b
        RSS = float(res @ res)
        if abs(RSS - RSS0) < tol:
            break
        if abs(RSS - thresh) < tol or np.sum(solution) < 1e-9:
            break
        RSS0 = RSS
        smin, solution, spks, aset = update_smin(y - b, smin, solution, spks, aset, np.sqrt(thresh))
        if optimize_b:
            b = float(np.mean(y - solution))         # :180
    return solution, spks, b, g, smin, aset


def thresholded_oasisAR2(y, g, sn, smin=None, optimize_b=False, optimize_g=False, decimate=None, maxIter=10,
                         thresh_factor=1.0):
    """thresholded_oasisAR2.m:38-166 with optimize_g = false.  NB: in that configuration both loops (:96-126 and
    :133-163) exit at their first `abs(RSS-RSS0)<tol` test (RSS is recomputed from the unchanged solution), so the result
    is ONE oasisAR2 pass with smin = choose_smin(g, sn, 0.99999999) (:72) -- fitted to y - b with
    b = estimate_baseline_noise(y) (:129-130) when optimize_b; update_smin (:175-201, warm-started AR2) is unreachable and
    not restated.  optimize_g (update_g of the AR(2) kernel, :236-321) is not restated."""
    if optimize_g:
        raise NotImplementedError("thresholded_oasisAR2 oracle: optimize_g branch not restated")
    y = np.asarray(y, dtype=np.float64).ravel()
    T = y.size
    smin = choose_smin(g, sn, 0.99999999)
    thresh = thresh_factor * sn * sn * T
    tol = 1e-4
    b = 0.0
    if optimize_b:
        b, _ = estimate_baseline_noise(y)
    solution, spks, aset = oasisAR2(y - b, g, None, smin)
    res = y - solution - b
    RSS0 = float(res @ res)
    for _ in range(int(maxIter)):
        if len(aset) == 0:
            break
        res = y - solution - b
        RSS = float(res @ res)
        if abs(RSS - RSS0) < tol:
            break
        raise NotImplementedError("thresholded_oasisAR2 oracle: update_smin reached")   # pragma: no cover
    _ = thresh
    return solution, spks, b, np.asarray(g, dtype=np.float64)[:2], smin, aset


# ----------------------------------------------------------------------------- deconvolveCa
_DEFAULTS = dict(type="ar1", pars=None, sn=None, b=0.0, lam=0.0, optimize_b=False, optimize_pars=False,
                 optimize_smin=False, method="constrained", window=200, shift=100, smin=0.0, maxIter=10,
                 thresh_factor=1.0, extra_params=None, p_noise=0.9999, max_tau=100.0, tau_range=None,
                 remove_large_residuals=False)


def parse_options(*structs, **kw):
    """deconvolveCa.m:208-355: defaults <- struct fields <- name/value pairs ('lambda' spelled `lam`)."""
    o = dict(_DEFAULTS)
    for s in structs:
        if s:
            for k, v in s.items():
                o["lam" if k == "lambda" else k] = v
    for k, v in kw.items():
        o["lam" if k == "lambda" else k] = v
    return o


def deconvolveCa(y, options=None, **kw):
    """deconvolveCa.m:60-206.  Returns (c, s, options).  Supports type in {ar1, ar2} and
    method in {foopsi, constrained (ar1), thresholded (ar1)} -- the combinations reachable from the
    CNMF-E demos (SURVEY.md §2 row 18 lists the rest as out of scope)."""
    y = np.asarray(y, dtype=np.float64).ravel()
    o = parse_options(options, **kw)
    if y.size == 0:
        return np.array([]), np.array([]), o
    if o["sn"] is None or np.size(o["sn"]) == 0:
        o["sn"] = GetSn(y)
    pars = o["pars"]
    if pars is None or np.size(pars) == 0 or np.all(np.asarray(pars) == 0):
        if o["type"] == "ar1":
            try:
                o["pars"] = estimate_time_constant(y, 1, o["sn"])
            except Exception:
                return y * 0, y * 0, o
            if np.size(o["pars"]) != 1:
                o["pars"] = 0.0
                return np.zeros(y.size), np.zeros(y.size), o
        elif o["type"] == "ar2":
            o["pars"] = estimate_time_constant(y, 2, o["sn"])
            if np.size(o["pars"]) != 2:
                o["pars"] = np.array([0.0, 0.0])
                return np.zeros(y.size), np.zeros(y.size), o
        else:
            raise NotImplementedError(o["type"])
    b0 = o["b"]
    method = o["method"].lower()
    if method == "foopsi":
        if o["type"] == "ar1":
            if o["smin"] < 0:
                o["smin"] = abs(o["smin"]) * o["sn"]
            gmax = np.exp(-1.0 / o["max_tau"])
            c, s, b, g, _ = foopsi_oasisAR1(y - b0, o["pars"], o["lam"], o["smin"], o["optimize_b"],
                                            o["optimize_pars"], None, o["maxIter"], o["tau_range"], gmax)
            o["b"] = b + b0
            o["pars"] = g
        elif o["type"] == "ar2":
            if o["smin"] < 0:
                o["smin"] = abs(o["smin"]) * o["sn"] / max_ht(o["pars"])
            c, s, b, g, _ = foopsi_oasisAR2(y - b0, o["pars"], o["lam"], o["smin"])
            o["b"] = b + b0
            o["pars"] = g
        else:
            raise NotImplementedError(o["type"])
    elif method == "constrained":
        if o["type"] != "ar1":
            raise NotImplementedError("constrained_foopsi (legacy CVX/LARS path) is out of scope")
        c, s, b, g, lam, _ = constrained_oasisAR1(y, o["pars"], o["sn"], o["optimize_b"], o["optimize_pars"],
                                                  None, o["maxIter"], o["tau_range"])
        o["b"], o["pars"], o["lam"] = b, g, lam
    elif method == "thresholded":
        if o["type"] == "ar2":
            c, s, b, g, smin, _ = thresholded_oasisAR2(y, o["pars"], o["sn"], o["smin"], o["optimize_b"],
                                                       o["optimize_pars"], None, o["maxIter"], o["thresh_factor"])
            o["b"], o["pars"], o["smin"] = b, g, smin
            c = np.array(c, dtype=np.float64)
            c[~np.isfinite(c)] = 0
            return c, s, o
        c, s, b, g, smin, _ = thresholded_oasisAR1(y, o["pars"], o["sn"], o["optimize_b"], o["optimize_pars"],
                                                   None, o["maxIter"], o["thresh_factor"], o["p_noise"],
                                                   o["tau_range"])
        o["b"], o["pars"], o["smin"] = b, g, smin
    else:
        raise NotImplementedError(method)
    c = np.array(c, dtype=np.float64)
    c[~np.isfinite(c)] = 0
    return c, s, o


def gen_data(gam=0.95, noise=0.3, T=3000, framerate=30, firerate=0.5, b=0.0, N=20, seed=13):
    """functions/gen_data.m:30-41.  Spikes use MATLAB's rand stream (MT19937, column-major fill); the Gaussian
    noise cannot be reproduced (MATLAB ziggurat) and uses NumPy's instead."""
    rs = np.random.RandomState(seed)
    trueSpikes = (rs.rand(T, N).T < firerate / framerate)
    truth = trueSpikes.astype(np.float64)
    gam = np.atleast_1d(np.asarray(gam, dtype=np.float64))
    p = gam.size
    gv = np.concatenate([gam[::-1], [1.0]])
    for t in range(p, T):
        truth[:, t] = truth[:, t - p:t + 1] @ gv
    Y = b + truth + noise * rs.randn(N, T)
    return Y, truth, trueSpikes
