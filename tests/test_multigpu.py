"""GPU (>= 2 devices): the sharded search step under torchrun/NCCL — identical parameters and rewards on every
rank.  Skipped on a single-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu`."""
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
    assert p.returncode == 0 and "MULTIGPU_CHECK OK" in p.stdout, (p.stdout[-2000:], p.stderr[-2000:])
