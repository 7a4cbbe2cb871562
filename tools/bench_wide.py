#!/usr/bin/env python
"""Critic update at large batch: the wide (tensor-core) path vs the row-group path. python tools/bench_wide.py [B ...]"""
import sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import make_synthetic_transitions
from sac_td3_cudagraphs_pytorch_b200 import sac_hps
from sac_td3_cudagraphs_pytorch_b200.agents.agent import Agent
from sac_td3_cudagraphs_pytorch_b200.replay import pack_rows
from sac_td3_cudagraphs_pytorch_b200.wide import WideCritic

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

for B in [int(x) for x in sys.argv[1:]] or [4096, 65536]:
    td = make_synthetic_transitions(B, 11, 3, [-1.0] * 3, [1.0] * 3, seed=1)
    torch.manual_seed(0)
    ag = Agent({"ob_shape": (11,), "ac_shape": (3,)}, np.full(3, -1.0, np.float32), np.full(3, 1.0, np.float32),
               torch.device("cuda"), sac_hps(batch_size=B), seed=1)
    rows = pack_rows({k: v.cuda() for k, v in td.items()}, ag.fmt)
    args = ag.update_args(rows)
    t_row = timeit(lambda: ag.enqueue_critic_step(args, fused_opt=False))
    res = {}
    for prec in ("3xtf32", "tf32"):
        wc = WideCritic(ag, B, prec)
        res[prec] = timeit(lambda: wc.update_qnets(rows))
    gf = 81.0 * B / 65536  # 81.0 GF per critic step at B = 65 536
    print(f"B={B:6d} critic update: row-group path {t_row:8.3f} ms ({gf / t_row:6.1f} TFLOP/s) | wide 3xTF32 {res['3xtf32']:7.3f} ms "
          f"({gf / res['3xtf32']:6.1f} TFLOP/s, {t_row / res['3xtf32']:.1f}x) | wide TF32 {res['tf32']:7.3f} ms ({t_row / res['tf32']:.1f}x)"
          f"  finite={bool(torch.isfinite(ag.out).all())}", flush=True)
