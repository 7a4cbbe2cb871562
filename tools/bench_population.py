#!/usr/bin/env python
"""Throughput of a stacked population (config 4) on one GPU: agent-updates/s for N agents."""
import sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import make_synthetic_transitions
from sac_td3_cudagraphs_pytorch_b200 import sac_hps
from sac_td3_cudagraphs_pytorch_b200.population import Population

def main():
    ns = [int(x) for x in sys.argv[1:]] or [1, 8, 64]
    td = make_synthetic_transitions(20_000, 11, 3, [-1.0] * 3, [1.0] * 3)
    for n in ns:
        t0 = time.time()
        pop = Population(range(n), 11, 3, [-1.0] * 3, [1.0] * 3, sac_hps(), "cuda", seed=1, rb_capacity=20_000)
        pop.fill_replay(td)
        t_init = time.time() - t0
        for i in range(6):
            pop.iteration()
        torch.cuda.synchronize()
        K = 30 if n >= 256 else 150
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(K):
            pop.iteration()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print(f"N={n:5d}  {ms:9.3f} ms/iteration  {n / ms * 1e3:10.0f} agent-updates/s  (init {t_init:.1f}s, finite={bool(torch.isfinite(pop.out).all())})", flush=True)
        del pop; torch.cuda.empty_cache()

if __name__ == "__main__":
    main()
