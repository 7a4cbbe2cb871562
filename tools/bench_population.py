#!/usr/bin/env python
"""Throughput of a stacked population (BASELINE.json config 4) on one GPU: agent-updates/s for N agents.
    python tools/bench_population.py [--wide 3xtf32|tf32|row] [--algo sac|td3] N [N ...]"""
import argparse
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import make_synthetic_transitions  # noqa: E402  (synthetic data only)
from sac_td3_cudagraphs_pytorch_b200 import sac_hps, td3_hps  # noqa: E402
from sac_td3_cudagraphs_pytorch_b200.population import Population  # noqa: E402


def run(n, wide, algo, rb_rows=20_000, iters=None):
    td = make_synthetic_transitions(rb_rows, 11, 3, [-1.0] * 3, [1.0] * 3)
    t0 = time.time()
    hps = sac_hps() if algo == "sac" else td3_hps()
    pop = Population(range(n), 11, 3, [-1.0] * 3, [1.0] * 3, hps, "cuda", seed=1, rb_capacity=rb_rows,
                     wide=None if wide == "row" else wide)
    pop.fill_replay(td)
    t_init = time.time() - t0
    for i in range(6):
        pop.iteration()
    torch.cuda.synchronize()
    K = iters or (30 if n >= 256 else 150)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        pop.iteration()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"N={n:5d} {algo} {wide:7s} {ms:9.3f} ms/iteration  {n / ms * 1e3:10.0f} agent-updates/s  "
          f"(init {t_init:.1f}s, mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB, finite={bool(torch.isfinite(pop.out).all())})",
          flush=True)
    del pop
    torch.cuda.empty_cache()
    return ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--wide", default="3xtf32")
    ap.add_argument("--algo", default="sac")
    ap.add_argument("--iters", type=int, default=0)
    ap.add_argument("n", type=int, nargs="*", default=[1, 8, 64])
    a = ap.parse_args()
    for n in a.n:
        run(n, a.wide, a.algo, iters=a.iters or None)


if __name__ == "__main__":
    main()
