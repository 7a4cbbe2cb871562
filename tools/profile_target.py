#!/usr/bin/env python
"""Minimal target for ncu: a few learner iterations of one workload (graphs on), nothing else.
    python tools/profile_target.py [workload] [iterations]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402


def gather_target(rows=1 << 20):
    """A few stand-alone sampler launches at a bandwidth-relevant size (bench.py's roofline_gather)."""
    bench.WORKLOADS["td3_hopper"] = (*bench.WORKLOADS["td3_hopper"][:4], 2_000_000)
    ag, rb, eng, fmt = bench.build_learner("td3_hopper", "cuda:0", seed=1)
    for _ in range(6):
        rb.sample(rows)
    torch.cuda.synchronize()
    print("ok gather", rows)


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "td3_hopper"
    if workload == "gather":
        return gather_target()
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    bench.WORKLOADS[workload] = (*bench.WORKLOADS[workload][:4], 200_000)  # small replay: faster start-up
    ag, rb, eng, fmt = bench.build_learner(workload, "cuda:0", seed=1)
    for i in range(iters):
        eng.iteration(i)
    torch.cuda.synchronize()
    print("ok", workload, iters, {k: float(v) for k, v in eng.logs().items()})


if __name__ == "__main__":
    main()
