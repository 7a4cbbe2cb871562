#!/usr/bin/env python
"""Config 5: SAC Hopper, batch 65 536 per GPU, 1M-transition replay per GPU, data-parallel over W ranks
(NCCL gradient all-reduce). Launch with torchrun for W > 1. Prints updates/s (each update consumes W*B
transitions)."""
import os, sys, time
from pathlib import Path
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import make_synthetic_transitions
from sac_td3_cudagraphs_pytorch_b200 import sac_hps
from sac_td3_cudagraphs_pytorch_b200.agents.agent import Agent
from sac_td3_cudagraphs_pytorch_b200.dp import DataParallelLearner, GradComm
from sac_td3_cudagraphs_pytorch_b200.replay import ReplayBuffer

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    wide = sys.argv[3] if len(sys.argv) > 3 and sys.argv[3] != "row" else None
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rb = ReplayBuffer(1_000_000, dev, seed=11, agent_id=rank)
    for c in range(4):
        td = make_synthetic_transitions(250_000, 11, 3, [-1.0] * 3, [1.0] * 3, seed=1234 + c + 10 * rank)
        rb.extend({k: v.to(dev) for k, v in td.items()})
    torch.manual_seed(0)
    ag = Agent({"ob_shape": (11,), "ac_shape": (3,)}, np.full(3, -1.0, np.float32), np.full(3, 1.0, np.float32), dev,
               sac_hps(batch_size=B), rb=rb, seed=11, agent_id=rank)
    dp = DataParallelLearner(ag, rb, B, GradComm(), wide=wide)
    for i in range(3):
        dp.iteration(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(3, 3 + K):
        dp.iteration(i)
    e1.record()
    torch.cuda.synchronize()
    el = torch.tensor([e0.elapsed_time(e1) * 1e-3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(el, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(el) / K * 1e3
        gf = 128.9 * (B / 65536)
        print(f"DP world={world} B={B}/gpu: {ms:.3f} ms/update, {1e3 / ms:.1f} updates/s, {world * B / ms * 1e3 / 1e6:.2f} M transitions/s, "
              f"{world * gf / ms:.1f} TFLOP/s aggregate (critic: {wide or 'row-group FFMA2'} path), finite={bool(torch.isfinite(ag.out).all())}, out={ag.out[:4].tolist()}", flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()

if __name__ == "__main__":
    main()
