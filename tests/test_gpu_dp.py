"""Data-parallel learner (BASELINE.json config 5): W ranks with an NCCL gradient all-reduce == one rank on
the concatenated batch. Needs >= 2 GPUs (skipped on the single-GPU box; run with `gpurun --gpus 2`)."""
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
