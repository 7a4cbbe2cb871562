"""A population of independent SAC/TD3 learners stacked on one GPU (BASELINE.json config 4).

The reference's only scale-out story is one OS process per (env, seed) (spawner.py:148-178). Here N
learners — each with its own parameters, optimizer state, replay slice, step counters and Philox streams —
live in one arena `[N, 5, region]` and every kernel of the update takes the agent index from
`blockIdx.y`, so one graph replay advances all N learners by one iteration of orchestrator.py:337-352.
Nothing is shared and nothing is reduced between agents; sharding a population over GPUs is a partition
of the agent ids (`shard`) with no collective, and because every random draw is keyed on the GLOBAL agent
id the results do not depend on the partition (tests/test_gpu_population.py).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Mapping, Optional

import numpy as np
import torch

from . import _lib as L
from .agents.agent import HID_DIMS
from .agents.nets import Actor, Critic, TanhGaussActor
from .arena import Arena, make_layout, set_shadow_pairs
from .replay import pack_rows, row_format


def shard(n_total: int, world: int, rank: int) -> range:
    """Agent ids owned by `rank`: contiguous blocks, sizes differing by at most one."""
    q, r = divmod(n_total, world)
    lo = rank * q + min(rank, r)
    return range(lo, lo + q + (1 if rank < r else 0))


def init_agent_params(agent_id: int, seed: int, ob_dim: int, ac_dim: int, td3: bool, layer_norm: bool,
                      min_ac: torch.Tensor, max_ac: torch.Tensor, noise_std: float = 0.1):
    """Reference initialisation (orthogonal / zeros / ones, agents/nets.py:34-49) from a CPU generator
    seeded by (seed, GLOBAL agent id): the same agent gets the same weights on any shard."""
    state, threads = torch.get_rng_state(), torch.get_num_threads()
    torch.manual_seed((int(seed) * 1_000_003 + int(agent_id)) & 0x7FFFFFFFFFFFFFFF)
    torch.set_num_threads(1)  # LAPACK's QR rounds differently with different thread counts
    try:
        kw = {"layer_norm": layer_norm}
        if td3:
            actor = Actor((ob_dim,), (ac_dim,), HID_DIMS, min_ac, max_ac, exploration_noise=noise_std, device="cpu", **kw)
        else:
            actor = TanhGaussActor((ob_dim,), (ac_dim,), HID_DIMS, min_ac, max_ac, device="cpu", **kw)
        q1 = Critic((ob_dim,), (ac_dim,), HID_DIMS, device="cpu", **kw)
        q2 = Critic((ob_dim,), (ac_dim,), HID_DIMS, device="cpu", **kw)
    finally:
        torch.set_rng_state(state)
        torch.set_num_threads(threads)
    sd = lambda m: {k: v for k, v in m.state_dict().items() if k.startswith(("fc_stack", "head"))}
    return sd(actor), sd(q1), sd(q2)


class Population:
    def __init__(self, agent_ids, ob_dim: int, ac_dim: int, min_ac, max_ac, hps, device, seed: int = 0,
                 rb_capacity: int = 100_000, batch_size: Optional[int] = None, use_graphs: bool = True,
                 wide: Optional[str] = None):
        """wide: None = the batch-256 row-group kernels with the agent in blockIdx.y (each agent's layers stream from L2
        once per 8 rows: latency-optimal for ONE agent, ~10 TFLOP/s however many are stacked); "3xtf32" / "tf32" = the
        layer-by-layer tensor-core path (wide.py) with the agents stacked along the row dimension — n_agents x batch rows
        per launch, per-agent weights fetched by rank-3 TMA maps: what a population of hundreds of agents wants."""
        self.ids = list(agent_ids)
        assert self.ids == list(range(self.ids[0], self.ids[0] + len(self.ids))), "agent ids must be contiguous"
        self.N, self.base = len(self.ids), self.ids[0]
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise L.B2rlError("Population needs a CUDA device (there is no CPU path)")
        self._lib = L.load()
        L.init_device(self.device)
        self.hps, self.seed = hps, int(seed)
        self.td3 = bool(hps.prefer_td3_over_sac)
        self.ob_dim, self.ac_dim = int(ob_dim), int(ac_dim)
        self.B = int(batch_size or hps.batch_size)
        self.use_graphs = use_graphs
        self.min_ac = torch.tensor(np.asarray(min_ac), dtype=torch.float, device=self.device)
        self.max_ac = torch.tensor(np.asarray(max_ac), dtype=torch.float, device=self.device)
        ln = bool(hps.layer_norm)
        self.fmt = row_format(self.ob_dim, self.ac_dim)
        self.layout = make_layout(self.ob_dim, self.ac_dim, self.td3, ln)
        self.arena = Arena(self.layout, self.device, n_agents=self.N)
        N, dev, lay = self.N, self.device, self.layout
        self.autotune = (not self.td3) and bool(hps.autotune)

        for g, aid in enumerate(self.ids):
            a, q1, q2 = init_agent_params(aid, seed, ob_dim, ac_dim, self.td3, ln, self.min_ac.cpu(), self.max_ac.cpu(),
                                          float(hps.actor_noise_std) if self.td3 else 0.1)
            for net, src in ((lay.actor, a), (lay.critic[0], q1), (lay.critic[1], q2)):
                with torch.no_grad():
                    for name, view in self.arena.named(net, L.REGION_P, g).items():
                        view.copy_(src[name])
        with torch.no_grad():
            self.arena.flat[:, L.REGION_T].copy_(self.arena.flat[:, L.REGION_P])
        self.arena.sync_shadows()

        self.counters = torch.zeros(N, 8, dtype=torch.int64, device=dev)
        self.out = torch.zeros(N, 8, dtype=torch.float32, device=dev)
        self.alpha_state = torch.zeros(N, 5, dtype=torch.float32, device=dev)
        if not self.td3:
            self.alpha_state[:, 0] = math.log(hps.alpha_init)
        self.sumsq = torch.zeros(N, dtype=torch.float32, device=dev)
        self.sumsq_scratch = torch.zeros(N * 64, dtype=torch.float32, device=dev)
        ws = self._lib.b2rl_workspace_floats(self.B)
        if ws < 0:
            raise L.B2rlError(f"batch size {self.B} must be >= 1")
        self.workspace = torch.zeros(N, ws, dtype=torch.float32, device=dev)
        self.rows = torch.zeros(N, self.B, self.fmt.row_stride, dtype=torch.float32, device=dev)
        self.idx = torch.zeros(N, self.B, dtype=torch.int64, device=dev)
        self.capacity = int(rb_capacity)
        self.storage = torch.zeros(N, self.capacity, self.fmt.row_stride, dtype=torch.float32, device=dev)
        self.size = 0
        self.iterations = 0
        self.graphs: dict = {}
        self.launches_per_variant: dict = {}
        self._hyper = L.Hyper(
            td3=int(self.td3), bcq_mix=int(bool(hps.bcq_style_targ_mix)),
            targ_smoothing=int(bool(hps.targ_actor_smoothing)) if self.td3 else 0, autotune=int(self.autotune),
            gamma=float(hps.gamma), td3_std=float(hps.td3_std) if self.td3 else 0.0,
            td3_c=float(hps.td3_c) if self.td3 else 0.0, targ_ent=float(-self.ac_dim), seed=self.seed)
        self.args = self._make_args()
        self._lo: Optional[torch.Tensor] = None
        self.wide = wide
        self.wide_q = self.wide_pi = None
        if wide:
            from .wide import WideActor, WideCritic, _Owner
            own = _Owner(self)
            self.wide_q = WideCritic(own, self.B, wide)
            self.wide_pi = WideActor(own, self.B, wide)

    # ------------------------------------------------------------------ replay
    def fill_replay(self, td: Mapping[str, torch.Tensor], agent: Optional[int] = None) -> None:
        """Load transitions (the reference's six keys) into one agent's slice, or the same data into all."""
        rows = pack_rows({k: v.to(self.device) for k, v in td.items()}, self.fmt)
        n = rows.shape[0]
        assert n <= self.capacity
        if agent is None:
            self.storage[:, :n] = rows
        else:
            self.storage[agent, :n] = rows
        self.size = max(self.size, n)
        self.counters[:, L.CTR_SIZE] = self.size

    # ------------------------------------------------------------------ enqueue
    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def _make_args(self) -> L.UpdateArgs:
        lay = self.layout
        a = L.UpdateArgs()
        a.hp, a.fmt = self._hyper, self.fmt
        a.actor = lay.actor.c_struct()
        a.critic[0], a.critic[1] = lay.critic[0].c_struct(), lay.critic[1].c_struct()
        a.batch, a.n_agents, a.agent_base = self.B, self.N, self.base
        a.region_stride, a.arena_agent_stride = lay.region, self.arena.agent_stride
        a.arena, a.rows, a.rows_agent_stride = self.arena.flat.data_ptr(), self.rows.data_ptr(), self.rows[0].numel()
        a.min_ac, a.max_ac = self.min_ac.data_ptr(), self.max_ac.data_ptr()
        a.log_alpha = self.alpha_state.data_ptr()
        a.counters = self.counters.data_ptr()
        a.workspace, a.workspace_agent_stride = self.workspace.data_ptr(), self.workspace[0].numel()
        a.out = self.out.data_ptr()
        return a

    def _adam(self, segs) -> None:
        a = L.AdamArgs()
        for i, s in enumerate(segs):
            a.seg[i] = s
        a.n_seg, a.n_agents = len(segs), self.N
        a.polyak, a.clip_norm = float(self.hps.polyak), float(self.hps.clip_norm)
        a.beta1, a.beta2, a.eps = 0.9, 0.999, 1e-8
        a.region_stride, a.arena_agent_stride = self.layout.region, self.arena.agent_stride
        a.arena, a.counters, a.grad_sumsq = self.arena.flat.data_ptr(), self.counters.data_ptr(), self.sumsq.data_ptr()
        a.lo, a.lo_agent_stride = L.ptr(self._lo), 2 * self.layout.region
        set_shadow_pairs(a, self.layout)
        L.check(self._lib.b2rl_adam_polyak_multi(C.byref(a), self._stream()), "adam_polyak_multi")

    def lo_mirror(self) -> torch.Tensor:
        """[N][2][region] lo parts of every agent's online / target regions for the 3xTF32 products (Agent.lo_mirror);
        kept current by every Adam / Polyak launch of the population."""
        if self._lo is None:
            self._lo = torch.zeros(self.N, 2, self.layout.region, dtype=torch.float32, device=self.device)
            self.refresh_lo()
        return self._lo

    def refresh_lo(self) -> None:
        if self._lo is not None:
            stk = L.Stack(self.N, self.base, self.arena.agent_stride, 2 * self.layout.region, 0, 0, 0)
            L.check(self._lib.b2rl_tc_split_lo(self.arena.flat.data_ptr(), self._lo.data_ptr(), 2 * self.layout.region,
                                               C.byref(stk), self._stream()), "tc_split_lo")

    def _enqueue(self, do_actor: bool, do_polyak: bool) -> int:
        lib, st, lay, h = self._lib, self._stream(), self.layout, self.hps
        seg = lambda b, e, lr, adam, pol, ctr=0, clip=False: L.Seg(b, e, lr, int(adam), int(pol), ctr, 1.0, int(clip))
        n = 0
        L.check(lib.b2rl_replay_sample_gather(
            self.storage.data_ptr(), self.storage[0].numel(), 0, self.fmt, self.B, self.N, None, self.idx.data_ptr(),
            self.rows.data_ptr(), C.c_uint64(self.seed), self.counters.data_ptr(), L.CTR_Q, 0, self.base, st), "gather")
        delay = int(h.actor_update_delay) if do_actor else 0
        extra = [seg(lay.actor.begin, lay.actor.end, 0.0, False, True)] if (self.td3 and do_polyak and delay == 0) else []
        if self.wide_q is not None:  # the tensor-core path, agents stacked along the rows (launch counts: tools/bench_population.py)
            self.wide_q.update_qnets(self.rows, polyak=do_polyak, extra_segs=extra)
            for j in range(delay):
                self.wide_pi.update_actor(self.rows, polyak=self.td3 and do_polyak and j == delay - 1)
            return -1
        fn = lib.b2rl_critic_update_td3 if self.td3 else lib.b2rl_critic_update_sac
        L.check(fn(C.byref(self.args), st), "critic_update")
        self._adam([seg(lay.critic[0].begin, lay.critic[1].end, float(h.qnets_lr), True, do_polyak, L.CTR_Q)] + extra)
        n += 4
        for j in range(delay):
            fn = lib.b2rl_actor_update_td3 if self.td3 else lib.b2rl_actor_update_sac
            L.check(fn(C.byref(self.args), st), "actor_update")
            clip = h.clip_norm > 0
            if clip:
                L.check(lib.b2rl_grad_sumsq(self.arena.flat.data_ptr(), lay.region, self.arena.agent_stride,
                                            lay.actor.begin, lay.actor.core_end, self.N, self.sumsq.data_ptr(),
                                            self.sumsq_scratch.data_ptr(), st), "grad_sumsq")
                n += 2
            self._adam([seg(lay.actor.begin, lay.actor.end, float(h.actor_lr), True,
                            self.td3 and do_polyak and j == delay - 1, L.CTR_PI, clip)])
            n += 3
            if self.autotune:
                L.check(lib.b2rl_alpha_update(C.byref(self.args), float(h.log_alpha_lr), st), "alpha_update")
                n += 1
        return n

    def iteration(self, i: Optional[int] = None) -> None:
        """One learner iteration (orchestrator.py:337-352) for every agent of the population."""
        i = self.iterations if i is None else i
        h = self.hps
        do_actor = (i % (h.actor_update_delay + 1) == 0)
        do_polyak = self.td3 or ((self.iterations + 1) % h.crit_targ_update_freq == 0)
        key = (do_actor, do_polyak)
        if not self.use_graphs:
            self.launches_per_variant[key] = self._enqueue(*key)
        else:
            g = self.graphs.get(key)
            if g is None:
                g = torch.cuda.CUDAGraph()
                torch.cuda.synchronize(self.device)
                with torch.cuda.graph(g):
                    self.launches_per_variant[key] = self._enqueue(*key)
                self.graphs[key] = g
            g.replay()
        self.iterations += 1

    # ------------------------------------------------------------------ inspection
    def agent_arena(self, g: int) -> torch.Tensor:
        return self.arena.flat[g]

    def losses(self) -> dict[str, torch.Tensor]:
        d = {"loss/qf_loss": self.out[:, L.OUT_QF_LOSS], "loss/actor_loss": self.out[:, L.OUT_ACTOR_LOSS]}
        if not self.td3:
            d["vitals/alpha"] = self.out[:, L.OUT_ALPHA]
            if self.autotune:
                d["loss/alpha_loss"] = self.out[:, L.OUT_ALPHA_LOSS]
        return d
