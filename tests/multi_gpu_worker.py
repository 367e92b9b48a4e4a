"""Worker of tests/test_gpu_multi.py (run under torchrun, one rank per GPU, NCCL): a 2 x 2-patch problem sharded over the
ranks, every stage of two update iterations compared with the float64 oracle on EVERY rank -- the owned patches' ring
weights, the exchanged A (exchange of the patches' rows), the cross-rank energy-weighted merge
(update_temporal_parallel.m:269-280) and the split final deconvTemporal -- through the host-buffer path and the resident
(sync_host=False) path.  Prints MULTI_GPU_PARITY_OK on every rank."""
import os
import sys

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def close(a, b, what, tol=1e-7):
    a, b = np.asarray(a), np.asarray(b)
    scale = max(1.0, float(np.abs(b).max()) if b.size else 1.0)
    err = float(np.abs(a - b).max()) if b.size else 0.0
    assert err <= tol * scale, "%s: max abs err %g (scale %g)" % (what, err, scale)


def main():
    import torch
    import torch.distributed as dist
    import gpu_cases as GC
    from oracle import cnmfe as OC, oasis as O
    from cnmf_e_b200.sources2d import Sources2D
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    D = GC.synthetic("multi_gpu")
    d1, d2, T = D["Y"].shape
    patch, rr = (48, 40), 9
    sn = O.GetSn(D["Y"].reshape(-1, T, order="F").astype(np.float64)).reshape(d1, d2, order="F")
    for shard_deconv in (True, False):
        for resident in (False, True):
            orc = OC.OracleSources2D(D["Y"], patch, ring_radius=rr)
            gpu = Sources2D(d1, d2, T, patch, ring_radius=rr, device=local, rank=rank, world_size=world,
                            options=dict(shard_deconv=shard_deconv))
            assert gpu.npatch == 4 and len(gpu.owned_patches()) == 4 // world
            gpu.load_video(D["Y"])
            for o in (orc, gpu):
                o.A, o.C = D["A0"].copy(), D["C0"].copy()
                o.P["sn"] = sn
            for it, alg in enumerate(("hals_thresh", "nnls")):
                orc.options["spatial_algorithm"] = gpu.options["spatial_algorithm"] = alg
                orc.update_background_parallel(); orc.update_spatial_parallel(IND=D["IND"]); orc.update_temporal_parallel()
                if resident:
                    if it == 0:
                        gpu._push_options(); gpu.push_neurons(); gpu.push_prev(); gpu.push_ring()
                        gpu.set_sn(sn)
                    gpu.update_background_parallel(sync_host=False)
                    gpu.update_spatial_parallel(IND=D["IND"], sync_host=False)
                    if world > 1:
                        gpu.exchange_spatial()
                    gpu.update_temporal_parallel(sync_host=False)
                    gpu.pull_ring(); gpu.pull_spatial(); gpu.pull_temporal()
                else:
                    gpu.update_background_parallel()
                    gpu.update_spatial_parallel(IND=D["IND"])
                    gpu.update_temporal_parallel()
                tag = "shard=%d resident=%d it=%d rank=%d" % (shard_deconv, resident, it, rank)
                for i in gpu.owned_patches():
                    mp = orc.patches()[i]
                    close(gpu.ring_as_sparse(i).toarray(), sp.csr_matrix(orc.W[mp]).toarray(), "W " + tag)
                    close(gpu.b0[i], orc.b0[mp], "b0 " + tag)
                close(gpu.A.toarray(), orc.A.toarray(), "A " + tag)
                close(gpu.C_raw, orc.C_raw, "C_raw " + tag)
                close(gpu.C, orc.C, "C " + tag)
                assert np.array_equal(gpu.S > 0, orc.S > 0), "spike support " + tag
                close(gpu.P["neuron_sn"], orc.P["neuron_sn"], "neuron_sn " + tag)
                if not resident:
                    close(gpu.b0_new, orc.b0_new, "b0_new " + tag)
            gpu.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    print("MULTI_GPU_PARITY_OK rank %d of %d" % (rank, world))


if __name__ == "__main__":
    main()
