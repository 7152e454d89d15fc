"""Config 5 on more than one GPU: parallel.gather_paths / RolloutStats.reduce over NCCL.  Needs >= 2 GPUs
(run with `gpurun --gpus 2`); skipped on a single-GPU box.  The gloo twin of this test runs on the CPU
(tests/test_parallel.py)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_nccl_gather_paths_two_ranks():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "nccl_gather_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "NCCL_GATHER_OK world=2" in r.stdout
