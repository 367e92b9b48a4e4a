"""CPU, gloo, world_size 2: the multi-GPU host logic -- patch ownership, the all-reduce of the energy-weighted merge
buffers (update_temporal_parallel.m:269-280) and of the disjoint A rows -- reproduces the single-process result.
The per-patch arithmetic is supplied by the oracle here (no GPU in this container); on the GPU box the same exchange
code runs over NCCL on the library's device buffers."""
import os
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_patch_owners_partition():
    from cnmf_e_b200.sources2d import patch_owners
    for npatch, ws in [(16, 8), (16, 3), (4, 2), (3, 8), (1, 1)]:
        o = patch_owners(npatch, ws)
        assert o.min() >= 0 and o.max() < ws and np.all(np.diff(o) >= 0)
        if npatch >= ws:
            cnt = np.bincount(o, minlength=ws)
            assert cnt.max() - cnt.min() <= 1 and cnt.min() >= 1


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import gen, cnmfe as OC
    from cnmf_e_b200.sources2d import patch_owners, merge_temporal
    import scipy.sparse as sp
    D = gen.make_synthetic(60, 50, 300, 5, seed=4, nblob=2)
    o = OC.OracleSources2D(D["Y"], (30, 50), ring_radius=6, options=dict(deconv_flag=False, maxIter=2))
    o.A, o.C = D["A0"], D["C0"].copy()
    o.update_background_parallel()
    patches = o.patches()
    owner = patch_owners(len(patches), world)
    K, T = o.C.shape
    num, den = np.zeros((K, T)), np.zeros(K)
    Acsr = sp.csr_matrix(o.A)
    for i, mp_ in enumerate(patches):
        if owner[i] != rank:
            continue
        tb, tp = o.block_pos[mp_], o.patch_pos[mp_]
        bm = o._block_mask(tb)
        ind = np.nonzero(np.asarray(Acsr[bm, :].sum(axis=0)).ravel() > 0)[0]
        if ind.size == 0:
            continue
        ipm = OC.ind_patch_mask(tp, tb).ravel(order="F")
        A_patch = np.asarray(Acsr[bm, :][:, ind].todense())[ipm, :]
        _, Craw_p, _, _ = OC.HALS_temporal(o._ysig(mp_, "temporal"), A_patch, o.C[ind], 2, None)
        aa = np.sum(A_patch ** 2, axis=0)
        num[ind] += Craw_p * aa[:, None]
        den[ind] += aa
    tn, td = torch.from_numpy(num), torch.from_numpy(den)
    dist.all_reduce(tn)
    dist.all_reduce(td)
    merged = merge_temporal(tn.numpy(), td.numpy())
    if rank == 0:
        ref = OC.OracleSources2D(D["Y"], (30, 50), ring_radius=6, options=dict(deconv_flag=False, maxIter=2))
        ref.A, ref.C = D["A0"], D["C0"].copy()
        ref.update_background_parallel()
        ref.update_temporal_parallel()
        # update_temporal_parallel subtracts the row minimum after the merge when deconv_flag is false
        q.put(float(np.abs((merged - merged.min(axis=1, keepdims=True)) - ref.C_raw).max()))
    dist.destroy_process_group()


def test_two_rank_merge_matches_single_process():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 1000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert q.get(timeout=10) < 1e-9


def test_trace_range_partition():
    """Split of the final deconvTemporal over ranks (opt-in `shard_deconv`): contiguous, complete, balanced."""
    from cnmf_e_b200.sources2d import trace_range
    for K, w in [(300, 8), (7, 3), (2, 4), (0, 2), (2400, 8), (1000, 1)]:
        r = [trace_range(K, i, w) for i in range(w)]
        assert r[0][0] == 0 and r[-1][1] == K
        assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1
