"""Partitioned model on 2 real GPUs (NCCL) against the single-GPU model: tests/multigpu_parity.py under torchrun.
Skipped on boxes with fewer than 2 GPUs."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_partitioned_model_matches_single_gpu():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + os.getpid() % 300), os.path.join(ROOT, "tests", "multigpu_parity.py")]
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=300)
    assert "MULTIGPU PARITY PASS" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
