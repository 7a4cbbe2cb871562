#!/usr/bin/env python
"""Time critic_fused / actor_fused alone for several batch sizes (how does time scale with #CTAs?)."""
import ctypes as C, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from sac_td3_cudagraphs_pytorch_b200 import _lib as L, sac_hps, td3_hps
from sac_td3_cudagraphs_pytorch_b200.agents.agent import Agent

lib = L.load()
for algo in ("td3", "sac"):
    for B in (4, 32, 64, 128, 256, 512):
        hps = (sac_hps if algo == "sac" else td3_hps)(batch_size=B)
        torch.manual_seed(0)
        ag = Agent({"ob_shape": (11,), "ac_shape": (3,)}, np.full(3, -1.0, np.float32), np.full(3, 1.0, np.float32),
                   torch.device("cuda"), hps)
        rows = torch.randn(B, ag.fmt.row_stride, device="cuda")
        rows[:, 15] = 0
        a = ag.update_args(rows)
        st = lambda: torch.cuda.current_stream().cuda_stream
        tc = bench.time_kernel(lambda: L.check(lib.b2rl_launch_single(C.byref(a), 0, st()))) * 1e6
        ta = bench.time_kernel(lambda: L.check(lib.b2rl_launch_single(C.byref(a), 2, st()))) * 1e6
        tw = bench.time_kernel(lambda: L.check(lib.b2rl_launch_single(C.byref(a), 1, st()))) * 1e6
        print(f"{algo} B={B:4d} CTAs={2*B//4:4d}  critic_fused {tc:7.2f} us  actor_fused {ta:7.2f} us  critic_wgrad {tw:6.2f} us", flush=True)
