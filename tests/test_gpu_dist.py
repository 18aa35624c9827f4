"""Data-parallel correctness on hardware (VERDICT r1 missing 2; SURVEY.md 4 `tests/dist/`, 8d C3; reference: Lightning DDP,
train.py:174-179): 2 (or more) NCCL ranks, one per GPU.  Skipped on a single-GPU box."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_overlapped_allreduce_equals_rank_sum_and_replicas_stay_identical():
    n = min(torch.cuda.device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29631", os.path.join(HERE, "_dist_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "DIST_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
