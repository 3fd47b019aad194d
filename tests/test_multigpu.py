"""GPU (>= 2 devices): the sharded search step under torchrun/NCCL — identical parameters and rewards on every
rank (eager and CUDA-graph steps), and 1-vs-2 result parity on the same global batch with SyncBN statistics.  Skipped on a single-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu`."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_rank_search_step_in_lockstep():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "scripts", "multigpu_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    print(p.stdout)
    assert p.returncode == 0, (p.stdout[-3000:], p.stderr[-3000:])
    for name in ("LOCKSTEP OK", "LOCKSTEP_GRAPH OK", "PARITY_1_vs_2 OK"):
        assert "MULTIGPU_CHECK " + name in p.stdout, (name, p.stdout[-3000:])
