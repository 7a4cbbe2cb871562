#!/usr/bin/env python
"""In-kernel phase timing of critic_fused (CTA 0, thread 0; clock64). Needs a timing build:
   B2RL_EXTRA_NVCC_FLAGS=-DB2RL_TIMING python -m sac_td3_cudagraphs_pytorch_b200.build --force"""
import ctypes as C, sys
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from sac_td3_cudagraphs_pytorch_b200 import _lib as L, td3_hps, sac_hps
from sac_td3_cudagraphs_pytorch_b200.agents.agent import Agent
lib = L.load()
algo = sys.argv[1] if len(sys.argv) > 1 else "td3"
B = 256
torch.manual_seed(0)
ag = Agent({"ob_shape": (11,), "ac_shape": (3,)}, np.full(3, -1.0, np.float32), np.full(3, 1.0, np.float32),
           torch.device("cuda"), (sac_hps if algo == "sac" else td3_hps)())
rows = torch.randn(B, ag.fmt.row_stride, device="cuda"); rows[:, 15] = 0
a = ag.update_args(rows)
st = torch.cuda.current_stream().cuda_stream
for _ in range(5):
    L.check(lib.b2rl_launch_single(C.byref(a), 0, st))
torch.cuda.synchronize()
buf = (C.c_longlong * 64)()
lib.b2rl_debug_timing.argtypes = [C.c_void_p]
assert lib.b2rl_debug_timing(buf) == 0
t = np.array(buf[:64], dtype=np.int64)
names = {0: "start", 1: "A.l1 gemm start", 2: "A.l1 gemm end", 3: "A.l1 sync end", 4: "A.l1 epi end", 5: "A.sync end",
         6: "A.l2 gemm end", 7: "A.l2 sync end", 8: "A.l2 epi end", 9: "A.sync end", 10: "A.head rowdot end",
         11: "T start (after sample)", 12: "T.l1 gemm start", 13: "T.l1 gemm end", 14: "T.l1 sync", 15: "T.l1 epi end",
         16: "T.sync", 17: "T.l2 gemm end", 18: "T.l2 sync", 19: "T.l2 epi end", 20: "T.sync", 21: "T.head end",
         22: "cluster exchange done", 23: "Q start (y, load_x)", 24: "Q.l1 gemm start", 25: "Q.l1 gemm end", 26: "Q.l1 sync",
         27: "Q.l1 epi end", 28: "Q.sync", 29: "Q.l2 gemm end", 30: "Q.l2 sync", 31: "Q.l2 epi end", 32: "Q.sync",
         33: "Q.head end", 34: "bwd start", 35: "bwd end"}
prev = t[0]
for k in range(36):
    print(f"{k:2d} {names.get(k,''):28s} +{(t[k]-prev):7d} cyc   total {(t[k]-t[0]):7d}")
    prev = t[k]
