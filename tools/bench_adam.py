#!/usr/bin/env python
"""Adam + Polyak launch of one Hopper learner: the transposing shadow-pair tiles (n_shadow = 3) vs stepping the w2n
shadows element-wise from their own gradient copies (n_shadow = 0)."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from sac_td3_cudagraphs_pytorch_b200 import _lib as L, sac_hps, td3_hps  # noqa: E402
from sac_td3_cudagraphs_pytorch_b200.agents.agent import Agent  # noqa: E402


def main():
    torch.manual_seed(0)
    ag = Agent({"ob_shape": (11,), "ac_shape": (3,)}, np.full(3, -1.0, np.float32), np.full(3, 1.0, np.float32),
               torch.device("cuda"), sac_hps(), seed=1)
    ag.arena.flat[0, 4].normal_()
    ag.counters[:3] = 5
    lib = ag._lib
    st = lambda: torch.cuda.current_stream().cuda_stream
    for polyak in (True, False):
        for what, segs in (("critics", ag.critic_segs(polyak)), ("actor", ag.actor_segs(polyak))):
            for n_sh, dbg in ((3, 0), (0, 0)):
                a = ag._adam_args(segs)
                a.n_shadow, a.reserved2 = n_sh, dbg
                t = bench.time_kernel(lambda: L.check(lib.b2rl_adam_polyak_multi(C.byref(a), st())), iters=400, warm=40)
                print(f"{what:8s} polyak={polyak!s:5s} n_shadow={n_sh} dbg={dbg}: {t * 1e6:7.2f} us")


if __name__ == "__main__":
    main()
