"""Multi-GPU parity on hardware: ONE batch sharded by bases over 2 GPUs (torchrun, one process per GPU), processed
through the C ABI, merged with fpl_allreduce_stats (NCCL on the library's stream), records gathered — compared with a
single pass of the oracle.  Needs two GPUs (run with `gpurun --gpus 2`); skipped on a one-GPU box."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("optset", ["cut_polyx_cplx", "default_se"])
def test_two_gpu_shard_merge_gather(optset):
    if _n_gpus() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_worker.py"), optset]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-3000:]
    assert "MGPU_OK world=2" in p.stdout
