#!/usr/bin/env python
"""Minimal target for ncu: a few learner iterations of one workload (graphs on), nothing else.
    python tools/profile_target.py [workload] [iterations]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "td3_hopper"
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    bench.WORKLOADS[workload] = (*bench.WORKLOADS[workload][:4], 200_000)  # small replay: faster start-up
    ag, rb, eng, fmt = bench.build_learner(workload, "cuda:0", seed=1)
    for i in range(iters):
        eng.iteration(i)
    torch.cuda.synchronize()
    print("ok", workload, iters, {k: float(v) for k, v in eng.logs().items()})


if __name__ == "__main__":
    main()
