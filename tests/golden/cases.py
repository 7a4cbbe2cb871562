"""Golden-fixture case table, shared by make_golden.py (authoring container, runs the
reference) and the tests (any machine, runs the oracle / the CUDA path). Test infrastructure."""
from __future__ import annotations

import numpy as np

from . import portable as P

# keys agents/agent.py reads, with the values of tasks/defaults/{sac,td3}.yml
_COMMON = dict(cuda=False, compile=False, cudagraphs=False, layer_norm=True, actor_lr=3e-4,
               clip_norm=0.0, segment_len=1, batch_size=256, gamma=0.99, polyak=0.005,
               actor_update_delay=2)
SAC = dict(_COMMON, qnets_lr=1e-3, prefer_td3_over_sac=False, bcq_style_targ_mix=False,
           crit_targ_update_freq=1, alpha_init=0.2, autotune=True, log_alpha_lr=1e-3)
TD3 = dict(_COMMON, qnets_lr=3e-4, prefer_td3_over_sac=True, bcq_style_targ_mix=True,
           actor_noise_std=0.1, targ_actor_smoothing=True, td3_std=0.2, td3_c=0.5)

CASES = {
    # BASELINE.json configs[0]/[1]/[2] shapes
    "sac_hopper": dict(base=SAC, ob=11, ac=3, lo=[-1.0] * 3, hi=[1.0] * 3, B=256, N=1024, iters=6, seed=1000),
    "td3_hopper": dict(base=TD3, ob=11, ac=3, lo=[-1.0] * 3, hi=[1.0] * 3, B=256, N=1024, iters=6, seed=2000),
    "sac_humanoid": dict(base=SAC, ob=376, ac=17, lo=[-0.4] * 17, hi=[0.4] * 17, B=64, N=256, iters=4, seed=3000),
    # BASELINE.json configs[2] at its own batch size (the wide first layers' clamped tails are B-dependent code)
    "sac_humanoid_b256": dict(base=SAC, ob=376, ac=17, lo=[-0.4] * 17, hi=[0.4] * 17, B=256, N=1024, iters=4, seed=3500),
    # option coverage: no LayerNorm, fixed alpha, BCQ mix under SAC, asymmetric per-dim bounds, odd dims
    "sac_noln_fixedalpha_bcq": dict(base=SAC, ob=5, ac=2, lo=[-1.0, -0.5], hi=[2.0, 0.5], B=32, N=128, iters=4,
                                    seed=4000, over=dict(layer_norm=False, autotune=False, bcq_style_targ_mix=True,
                                                         batch_size=32)),
    # TD3 hard min, no smoothing, actor grad clipping, A=1
    "td3_hardmin_nosmooth_clip": dict(base=TD3, ob=4, ac=1, lo=[-3.0], hi=[3.0], B=32, N=128, iters=4, seed=5000,
                                      over=dict(bcq_style_targ_mix=False, targ_actor_smoothing=False,
                                                clip_norm=0.05, batch_size=32)),
    # deliberately saturated tanh policy: the reference's own fp32 result is ill-conditioned here
    # (meta.reference_fp32_vs_fp64_oracle is ~1e-2); parity on it is judged against float64
    "sac_saturated": dict(base=SAC, ob=24, ac=8, lo=[-1.0] * 8, hi=[1.0] * 8, B=64, N=256, iters=3, seed=7000,
                          head_scale=4.0, over=dict(batch_size=64)),
    # Ant shapes (27 / 8): the actor's first layer (27 inputs) is staged in shared memory while the critics' (35 inputs) is
    # streamed from global memory — both first-layer forms inside one kernel instantiation; TD3 with the default options
    "td3_ant_mixed_first": dict(base=TD3, ob=27, ac=8, lo=[-1.0] * 8, hi=[1.0] * 8, B=64, N=256, iters=4, seed=9000,
                                over=dict(batch_size=64)),
    # SAC with clipping and a slower target cadence
    "sac_clip_targfreq2": dict(base=SAC, ob=17, ac=6, lo=[-1.0] * 6, hi=[1.0] * 6, B=64, N=256, iters=6, seed=6000,
                               over=dict(clip_norm=0.5, crit_targ_update_freq=2, batch_size=64)),
}


# ragged batches (B % 16 != 0: the last CTA group masks its tail) and B < 16; oracle-checked only (no fixture)
RAGGED_CASES = {
    "sac_ragged40": dict(base=SAC, ob=11, ac=3, lo=[-1.0] * 3, hi=[1.0] * 3, B=40, N=128, iters=1, seed=8000,
                         over=dict(batch_size=40)),
    "td3_ragged7": dict(base=TD3, ob=17, ac=6, lo=[-1.0] * 6, hi=[1.0] * 6, B=7, N=64, iters=1, seed=8100,
                        over=dict(batch_size=7)),
}


def hps_dict(case: dict) -> dict:
    h = dict(case["base"])
    h.update(case.get("over", {}))
    h["batch_size"] = case["B"]
    return h


def case_inputs(name: str) -> dict:
    """All inputs of a case, regenerated bit-identically from seeds."""
    c = CASES[name] if isinstance(name, str) else name
    h = hps_dict(c)
    td3 = h["prefer_td3_over_sac"]
    s = c["seed"]
    a_out = c["ac"] if td3 else 2 * c["ac"]
    inp = dict(
        hps=h, ob=c["ob"], ac=c["ac"], B=c["B"], iters=c["iters"],
        min_ac=np.asarray(c["lo"], np.float32), max_ac=np.asarray(c["hi"], np.float32),
        actor=P.mlp_params(s + 1, c["ob"], a_out, h["layer_norm"], head_scale=c.get("head_scale", 0.25)),
        q1=P.mlp_params(s + 2, c["ob"] + c["ac"], 1, h["layer_norm"]),
        q2=P.mlp_params(s + 3, c["ob"] + c["ac"], 1, h["layer_norm"]),
        storage=P.transitions(s + 4, c["N"], c["ob"], c["ac"], c["lo"], c["hi"]),
        idx=[P.indices(s + 100 + i, c["N"], c["B"]) for i in range(c["iters"])],
        eps_q=[P.noise(s + 200 + i, c["B"], c["ac"]) for i in range(c["iters"])],
        eps_pi=[[P.noise(s + 300 + 10 * i + j, c["B"], c["ac"]) for j in range(h["actor_update_delay"])]
                for i in range(c["iters"])],
        eps_alpha=[[P.noise(s + 400 + 10 * i + j, c["B"], c["ac"]) for j in range(h["actor_update_delay"])]
                   for i in range(c["iters"])],
    )
    return inp


def inputs_digest(inp: dict) -> str:
    """sha256 over every input tensor of a case (guards the numpy stream the fixtures rest on)."""
    import hashlib
    import torch
    h = hashlib.sha256()
    def feed(x):
        if isinstance(x, dict):
            for k in sorted(x):
                feed(x[k])
        elif isinstance(x, (list, tuple)):
            for y in x:
                feed(y)
        elif isinstance(x, torch.Tensor):
            h.update(x.contiguous().numpy().tobytes())
        elif isinstance(x, np.ndarray):
            h.update(np.ascontiguousarray(x).tobytes())
    for k in ("actor", "q1", "q2", "storage", "idx", "eps_q", "eps_pi", "eps_alpha", "min_ac", "max_ac"):
        feed(inp[k])
    return h.hexdigest()
