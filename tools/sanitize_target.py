#!/usr/bin/env python
"""Target for compute-sanitizer (memcheck / racecheck / synccheck): a few EAGER learner iterations of the row-group path
(TD3 and SAC, batch 256, the in-kernel sampler and replay write included), one wide (tcgen05) critic + actor step and
one stacked-population iteration. Test tooling; run on a GPU box:
    compute-sanitizer --tool racecheck python tools/sanitize_target.py [row|wide|pop ...]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import make_synthetic_transitions  # noqa: E402  (synthetic data only)
from sac_td3_cudagraphs_pytorch_b200 import sac_hps, td3_hps  # noqa: E402
from sac_td3_cudagraphs_pytorch_b200.agents.agent import Agent  # noqa: E402
from sac_td3_cudagraphs_pytorch_b200.engine import LearnerEngine  # noqa: E402
from sac_td3_cudagraphs_pytorch_b200.replay import ReplayBuffer  # noqa: E402


def learner(algo, O, A, B, n_rows=4096, seed=3):
    hps = (sac_hps if algo == "sac" else td3_hps)(batch_size=B)
    rb = ReplayBuffer(n_rows, "cuda", seed=seed)
    rb.extend({k: v.cuda() for k, v in make_synthetic_transitions(n_rows, O, A, [-1.0] * A, [1.0] * A, seed=11).items()})
    torch.manual_seed(0)
    ag = Agent({"ob_shape": (O,), "ac_shape": (A,)}, np.full(A, -1.0, np.float32), np.full(A, 1.0, np.float32),
               torch.device("cuda"), hps, rb=rb, seed=seed)
    return ag, rb


def row_path():
    for algo, O, A in (("td3", 11, 3), ("sac", 11, 3), ("sac", 376, 17)):
        ag, rb = learner(algo, O, A, 256)
        eng = LearnerEngine(ag, use_graphs=False)
        for i in range(3):
            eng.iteration(i)
        obs = torch.randn(4, O, device="cuda")
        ag.predict_device(obs, explore=True)
        torch.cuda.synchronize()
        print("row", algo, O, A, {k: float(v) for k, v in eng.logs().items()})


def wide_path():
    from sac_td3_cudagraphs_pytorch_b200.dp import DataParallelLearner
    for algo in ("sac", "td3"):
        ag, rb = learner(algo, 11, 3, 1024)
        dp = DataParallelLearner(ag, rb, 1024, wide="3xtf32", graphs=False)
        for i in range(2):
            dp.iteration(i)
        torch.cuda.synchronize()
        print("wide", algo, float(ag.out[0]), float(ag.out[1]))


def pop_path():
    from sac_td3_cudagraphs_pytorch_b200.population import Population
    for wide in (None, "3xtf32"):
        pop = Population(range(3), 11, 3, [-1.0] * 3, [1.0] * 3, sac_hps(batch_size=256), "cuda", seed=5, rb_capacity=2048,
                         use_graphs=False, **({"wide": wide} if wide else {}))
        pop.fill_replay(make_synthetic_transitions(2048, 11, 3, [-1.0] * 3, [1.0] * 3, seed=12))
        for i in range(2):
            pop.iteration()
        torch.cuda.synchronize()
        print("pop", wide, pop.out[:, 0].tolist())


if __name__ == "__main__":
    what = sys.argv[1:] or ["row", "wide"]
    for w in what:
        {"row": row_path, "wide": wide_path, "pop": pop_path}[w]()
    print("sanitize target done")
