"""Multi-GPU parity: the same 2 x 2-patch problem under torchrun with 2 (and, when present, 4) ranks -- one per
GPU, NCCL -- against the float64 oracle on every rank (tests/multi_gpu_worker.py).  Covers exchange_spatial, the cross-rank
merge of update_temporal_parallel.m:269-280 on device buffers and the split deconvTemporal (cnmfe_update_temporal_finish_part).
Skipped on a single-GPU box; world size 1 runs everywhere as the control."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _run(world):
    port = 29600 + (os.getpid() % 300) + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "multi_gpu_worker.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-4000:]
    assert r.stdout.count("MULTI_GPU_PARITY_OK") == world, r.stdout[-4000:]


def test_world1_control(built_lib):
    _run(1)


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_patches_match_oracle(built_lib, world):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    _run(world)
