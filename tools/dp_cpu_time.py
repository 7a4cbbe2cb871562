#!/usr/bin/env python
"""Host time to ENQUEUE one data-parallel wide iteration vs the device time it takes (is config 5 launch-bound?)."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import make_synthetic_transitions
from sac_td3_cudagraphs_pytorch_b200 import sac_hps
from sac_td3_cudagraphs_pytorch_b200.agents.agent import Agent
from sac_td3_cudagraphs_pytorch_b200.dp import DataParallelLearner, GradComm
from sac_td3_cudagraphs_pytorch_b200.replay import ReplayBuffer
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda")
rb = ReplayBuffer(1_000_000, dev, seed=11)
td = make_synthetic_transitions(250_000, 11, 3, [-1.0] * 3, [1.0] * 3, seed=1234)
for _ in range(4):
    rb.extend({k: v.to(dev) for k, v in td.items()})
torch.manual_seed(0)
ag = Agent({"ob_shape": (11,), "ac_shape": (3,)}, np.full(3, -1.0, np.float32), np.full(3, 1.0, np.float32), dev, sac_hps(batch_size=B), rb=rb, seed=11)
dp = DataParallelLearner(ag, rb, B, GradComm(), wide="3xtf32")
for i in range(6):
    dp.iteration(i)
torch.cuda.synchronize()
host = []
t_all = time.perf_counter()
for i in range(6, 36):
    t0 = time.perf_counter(); dp.iteration(i); host.append(time.perf_counter() - t0)
torch.cuda.synchronize()
t_all = time.perf_counter() - t_all
print(f"B={B}: host enqueue {np.mean(host) * 1e3:.3f} ms/iteration (critic-only {np.mean(host[1::3] + host[2::3]) * 1e3:.3f}, with actor {np.mean(host[0::3]) * 1e3:.3f}); wall {t_all / 30 * 1e3:.3f} ms/iteration")
