"""Large-batch data-parallel learner (BASELINE.json config 5): W replicas of one agent, each with its own
replay buffer and its own batch; gradients of the mean loss are summed over ranks with one
`all_reduce` per optimizer step on the flat gradient span of the arena (NCCL over NVLink on GPUs) and the
`1/W` is folded into the Adam kernel (`grad_scale`). The reference has no counterpart (no
torch.distributed anywhere); the parity oracle is a single rank run on the concatenation of
the W local batches (mean of equal-sized means = global mean).

Replicas stay bit-identical: every rank applies the same reduced gradient with the same arithmetic. The
natural-layout shadows (`w2n`) never see a gradient of their own — a ring all-reduce may sum two positions of the
buffer in different orders — the Adam launch steps `w2t` and writes the result to both layouts (adam.cu).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib as L
from .arena import Arena, ArenaLayout


class GradComm:
    """Sum-all-reduce over a torch.distributed group (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def all_reduce_sum(self, t: torch.Tensor) -> None:
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)


def reduce_grad_span(arena: Arena, layout: ArenaLayout, comm: GradComm, which: str, agent: int = 0) -> None:
    """All-reduce the gradients (region 4) of the critics or of the actor, then rebuild the w2n shadows."""
    if comm.world == 1:  # nothing to reduce, and the weight-gradient kernels wrote the shadows themselves
        return
    nets = layout.critic if which == "critic" else [layout.actor]
    g = arena.flat[agent, L.REGION_G]
    # one contiguous bucket (a shadow's gradient slot in between rides along unused: the Adam launch derives the w2n
    # shadows from the reduced w2t gradient through its transposing tiles, so replicas stay bit-identical with no copy)
    comm.all_reduce_sum(g[nets[0].begin:nets[-1].core_end])


class DataParallelLearner:
    """One rank of the data-parallel learner. `agent` must be constructed identically on every rank (same
    torch seed => same initial parameters) except `agent_id=rank`, which keys its index/noise draws."""

    def __init__(self, agent, rb, batch_size: int, comm: Optional[GradComm] = None, wide: Optional[str] = None,
                 graphs: Optional[bool] = None):
        """wide: None = row-group kernels; "3xtf32" / "tf32" = the critic step on the layer-by-layer tensor-core path
        (wide.WideCritic), which is what a batch of tens of thousands wants."""
        self.agent, self.rb, self.B = agent, rb, int(batch_size)
        self.wide = self.wide_actor = None
        if wide:
            from .wide import WideActor, WideCritic
            self.wide = WideCritic(agent, self.B, wide)
            if not agent.hps.clip_norm > 0:  # (the wide actor step has no gradient clipping)
                self.wide_actor = WideActor(agent, self.B, wide)
        self.comm = comm or GradComm()
        dev = agent.device
        self.rows = torch.zeros(self.B, agent.fmt.row_stride, dtype=torch.float32, device=dev)
        self.idx = torch.zeros(self.B, dtype=torch.int64, device=dev)
        self.launches = 0
        # One CUDA graph per iteration variant (actor updates? Polyak?) when the batch is sampled on the device: the wide
        # path is ~100 launches per iteration, ~0.9 ms of host time to enqueue — all of a batch-16 384 iteration
        # (tools/dp_cpu_time.py). With W > 1 the NCCL all-reduces are captured INSIDE the iteration graph (NCCL collectives
        # are stream-capturable; every rank replays the same sequence), so the multi-rank iteration is one graph launch
        # too; other backends (gloo in the CPU-side tests) stay eager.
        capturable = self.comm.world == 1 or (dist.is_initialized() and dist.get_backend(self.comm.group) == "nccl")
        self.graphs = (capturable and rb is not None and bool(agent.hps.cudagraphs)) if graphs is None else bool(graphs)
        self._graphs: dict[tuple, torch.cuda.CUDAGraph] = {}
        self._size_on_device = -1
        self._warm: set = set()

    def _seg(self, begin, end, lr, polyak, counter, clip=False):
        return L.Seg(begin, end, lr, 1, int(polyak), counter, 1.0 / self.comm.world, int(clip))

    def iteration(self, i: int, rows: Optional[torch.Tensor] = None, eps_q=None, eps_pi=None, eps_alpha=None) -> None:
        """One iteration of orchestrator.py:337-352 with gradient averaging over ranks. `rows` / `eps_*`
        inject this rank's batch and noise (parity tests); default: device-side sampling."""
        ag, lib, lay, h, W = self.agent, self.agent._lib, self.agent.layout, self.agent.hps, self.comm.world
        do_actor = (i % (h.actor_update_delay + 1) == 0)
        do_polyak = ag.td3 or ((ag.qnet_updates_so_far + 1) % h.crit_targ_update_freq == 0)
        if rows is None and self._size_on_device != len(self.rb):
            self._size_on_device = len(self.rb)
            ag.counters[L.CTR_SIZE] = self._size_on_device
        if self.graphs and rows is None and eps_q is None and eps_pi is None and eps_alpha is None:
            key = (do_actor, do_polyak)
            g = self._graphs.get(key)
            if g is None:  # first occurrence of the variant runs eagerly (lazy allocations), the second is captured
                if key not in self._warm:
                    self._warm.add(key)
                else:
                    g = torch.cuda.CUDAGraph()
                    torch.cuda.synchronize(ag.device)
                    with torch.cuda.graph(g):
                        self._enqueue(do_actor, do_polyak, None, None, None, None)
                    self._graphs[key] = g  # (capture does not execute: the replay below is this iteration)
            if g is not None:
                g.replay()
                return self._count(do_actor)
        self._enqueue(do_actor, do_polyak, rows, eps_q, eps_pi, eps_alpha)
        self._count(do_actor)

    def _count(self, do_actor: bool) -> None:
        ag = self.agent
        ag.qnet_updates_so_far += 1
        if do_actor:
            ag.actor_updates_so_far += int(ag.hps.actor_update_delay)

    def _enqueue(self, do_actor, do_polyak, rows, eps_q, eps_pi, eps_alpha) -> None:
        ag, lib, lay, h, W = self.agent, self.agent._lib, self.agent.layout, self.agent.hps, self.comm.world
        st = ag._stream()
        if rows is None:
            L.check(lib.b2rl_replay_sample_gather(
                self.rb.storage.data_ptr(), 0, 0, self.rb.fmt, self.B, 1, None, self.idx.data_ptr(),
                self.rows.data_ptr(), C.c_uint64(ag.seed), ag.counters.data_ptr(), L.CTR_Q, 0, ag.agent_id, st),
                "replay_sample_gather")
            rows = self.rows
        delay = int(h.actor_update_delay) if do_actor else 0

        if self.wide is not None:
            self.wide.update_qnets(rows, eps=ag._noise(eps_q, rows), adam=False)
        else:
            args = ag.update_args(rows, eps=ag._noise(eps_q, rows))
            fn = lib.b2rl_critic_update_td3 if ag.td3 else lib.b2rl_critic_update_sac
            L.check(fn(C.byref(args), st), "critic_update")
        reduce_grad_span(ag.arena, lay, self.comm, "critic")
        segs = [self._seg(lay.critic[0].begin, lay.critic[1].end, float(h.qnets_lr), do_polyak, L.CTR_Q)]
        if ag.td3 and do_polyak and delay == 0:
            segs.append(L.Seg(lay.actor.begin, lay.actor.end, 0.0, 0, 1, 0, 1.0, 0))
        ag._launch_adam(segs)
        for j in range(delay):
            e1 = None if eps_pi is None else eps_pi[j]
            e2 = None if eps_alpha is None else eps_alpha[j]
            args = ag.update_args(rows, eps=ag._noise(e1, rows), eps2=ag._noise(e2, rows))
            if self.wide_actor is not None:
                self.wide_actor.update_actor(rows, eps=ag._noise(e1, rows), adam=False)
            else:
                fn = lib.b2rl_actor_update_td3 if ag.td3 else lib.b2rl_actor_update_sac
                L.check(fn(C.byref(args), st), "actor_update")
            reduce_grad_span(ag.arena, lay, self.comm, "actor")
            clip = h.clip_norm > 0
            if clip:  # clip_grad_norm_ acts on the averaged gradient: sum of squares AFTER the all-reduce
                L.check(lib.b2rl_grad_sumsq(ag.arena.flat.data_ptr(), lay.region, ag.arena.agent_stride,
                                            lay.actor.begin, lay.actor.core_end, 1, ag._sumsq.data_ptr(),
                                            ag._sumsq_scratch.data_ptr(), st), "grad_sumsq")
            ag._launch_adam([self._seg(lay.actor.begin, lay.actor.end, float(h.actor_lr),
                                       ag.td3 and do_polyak and j == delay - 1, L.CTR_PI, clip)])
            if ag.autotune:
                if self.wide_actor is not None:
                    self.wide_actor.alpha_grad(rows, ag._noise(e2, rows))
                else:
                    L.check(lib.b2rl_alpha_update(C.byref(args), 0.0, st), "alpha_update (gradient only)")
                self.comm.all_reduce_sum(ag._alpha_state[1:2])
                L.check(lib.b2rl_alpha_adam(ag._alpha_state.data_ptr(), ag.counters.data_ptr(), 1,
                                            float(h.log_alpha_lr), 1.0 / W, ag.out.data_ptr(), st), "alpha_adam")
