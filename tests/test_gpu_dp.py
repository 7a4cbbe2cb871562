"""Data-parallel learner (BASELINE.json config 5): W ranks with a gradient all-reduce == one rank on the concatenated
batch. Two forms: NCCL over two GPUs (skipped on a single-GPU box; run with `gpurun --gpus 2`) — which also checks that
the captured multi-rank iteration graphs equal the eager loop — and two ranks SHARING one GPU over gloo, which exercises
the same CUDA kernels (all-reduced gradient span, grad_scale = 1/W in Adam, deferred temperature step) on any box."""
import os
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parent.parent


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_data_parallel_matches_single_rank():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29533", str(REPO / "tests" / "dp_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=REPO)
    print(p.stdout[-2000:], p.stderr[-2000:])
    assert p.returncode == 0 and "DP_OK" in p.stdout


def test_two_rank_data_parallel_on_one_gpu_over_gloo():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29534", str(REPO / "tests" / "dp_worker.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=REPO, env=dict(os.environ, B2RL_DP_BACKEND="gloo"))
    print(p.stdout[-3000:], p.stderr[-3000:])
    assert p.returncode == 0 and "DP_OK" in p.stdout
    assert p.stdout.count("DP_RESULT") == 4
