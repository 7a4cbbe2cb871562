#!/usr/bin/env python
"""Device-timed learner iterations/s of one workload with and without in-kernel sampling.
python tools/bench_engine.py [td3_hopper|sac_hopper|sac_humanoid] [iterations]"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench as B

wl = sys.argv[1] if len(sys.argv) > 1 else "sac_humanoid"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
dev = torch.device("cuda")
for fused in (False, True):
    ag, rb, eng, fmt = B.build_learner(wl, dev, seed=3)
    from sac_td3_cudagraphs_pytorch_b200.engine import LearnerEngine
    eng = LearnerEngine(ag, rb, fused_sample=fused)
    for i in range(30):
        eng.iteration(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(30, 30 + K):
        eng.iteration(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    print(f"{wl} fused_sample={fused}: {ms * 1e3:.2f} us/iteration = {1e3 / ms:.0f} updates/s", flush=True)
    del eng, ag, rb
