"""The wide (layer-by-layer, tensor-core) critic update for LARGE batches (BASELINE.json config 5).

``Agent.update_qnets`` (agents/agent.py:183-242) restated as a sequence of batch-parallel kernels
(csrc/wide.cu) around the tcgen05 hidden layers (csrc/tc_linear.cu): the row-group kernels that win at
batch 256 stream every layer's weights from L2 once per 8 rows and stay at ~10 TFLOP/s however large the
batch is; here weights are read once per 128 rows (TMA) and the 256x256 products run on the tensor cores
as TF32. Everything outside the products (LayerNorm, heads, TD target, losses, optimizer) is fp32 and the
Philox noise is keyed exactly as in the row path, so both paths draw the same samples; results agree to
TF32 accuracy (~1e-3, the north star's "looser stated bound" for tensor-core modes; tests/test_gpu_wide.py).
Operates on an ``Agent``'s arena, counters, workspace and optimizer: it is a drop-in for the critic step.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib as L

STREAM_CRITIC_EPS = 1  # csrc/rng.cuh


class WideCritic:
    def __init__(self, agent, batch: int, precision: str = "3xtf32"):
        """precision: "3xtf32" (default; hi/lo operand split, three MMAs per product: fp32-level accuracy) or "tf32"
        (single MMA, ~1e-3 per product, ~1e-2 on gradients because dLoss/dQ is a difference of Q and the TD target)."""
        assert precision in ("3xtf32", "tf32")
        self.x3 = precision == "3xtf32"
        self.ag, self.M = agent, int(batch)
        ag, M, dev = agent, self.M, agent.device
        self._lib = ag._lib
        f32 = dict(dtype=torch.float32, device=dev)
        O, A = ag.fmt.ob_dim, ag.ac_dim
        self.ldn = (O + A + 3) & ~3
        self.t1, self.t2 = torch.empty(M, 256, **f32), torch.empty(M, 256, **f32)     # scratch activations
        self.xn = torch.zeros(M, self.ldn, **f32)                                      # [next_obs | a']
        self.logp, self.qn, self.q = torch.zeros(M, **f32), torch.zeros(2, M, **f32), torch.zeros(2, M, **f32)
        self.xh1, self.xh2 = torch.empty(2, M, 256, **f32), torch.empty(2, M, 256, **f32)
        self.st1, self.st2 = torch.zeros(2, M, 2, **f32), torch.zeros(2, M, 2, **f32)
        self.P128, self.P8 = (M + 127) // 128, (M + 7) // 8
        self.part1, self.part2 = torch.zeros(2, self.P128, 3, 256, **f32), torch.zeros(2, self.P128, 3, 256, **f32)
        self.sq = torch.zeros(2, self.P8, 2, **f32)  # per-CTA {sum sq err, sum dQ} of wide_q_head
        self.ws = ag.workspace(M)
        self.wlo = torch.zeros(7, 256 * 256, **f32)  # lo parts of the seven 256x256 matrices a critic step multiplies by

    # pointers into the arena: region r, float offset o
    def _p(self, region: int, off: int) -> int:
        return self.ag.arena.flat.data_ptr() + 4 * (region * self.ag.layout.region + off)

    def _ws(self, which: int, slot: int) -> int:  # H1 H2 DZ1 DZ2 of common.cuh::ws_carve
        return self.ws.data_ptr() + 4 * ((which * 2 + slot) * self.M * 256)

    def _dz3(self, slot: int) -> int:
        return self.ws.data_ptr() + 4 * (8 * self.M * 256 + slot * self.M * L.MAX_OUT)

    def update_qnets(self, rows: torch.Tensor, eps: Optional[torch.Tensor] = None, eps_out: Optional[torch.Tensor] = None,
                     targ_out: Optional[torch.Tensor] = None, adam: bool = True) -> dict:
        """rows: the sampled batch [M][row_stride] (replay.Batch.rows). Enqueue-only (graph-capturable).
        adam=False: stop after the gradients (data parallel: all-reduce them, then step)."""
        ag, lib, M = self.ag, self._lib, self.M
        assert rows.shape == (M, ag.fmt.row_stride) and rows.is_contiguous()
        lay, st = ag.layout, ag._stream()
        O, A, rs = ag.fmt.ob_dim, ag.ac_dim, ag.fmt.row_stride
        ln = int(bool(ag.hps.layer_norm))
        RP, RT, RG = L.REGION_P, L.REGION_T, 4
        none = None

        def first(x_ptr, ldx, K, net, region, H, XH, stat):
            o = net.off
            L.check(lib.b2rl_wide_first(x_ptr, ldx, M, K, self._p(region, o["w1t"]), self._p(region, o["b1"]),
                                        self._p(region, o["g1"]) if ln else none, self._p(region, o["be1"]) if ln else none, ln,
                                        H, XH, stat, st), "wide_first")

        def lo_of(slot, w_ptr):  # the weights change every step: their lo parts are recomputed (256 KB each)
            if not self.x3:
                return none
            dst = self.wlo[slot].data_ptr()
            L.check(lib.b2rl_tc_split_lo(w_ptr, dst, 256 * 256, st), "tc_split_lo")
            return dst

        def hidden(x_ptr, net, region, H, XH, stat, slot):
            o = net.off
            w = self._p(region, o["w2n"])
            L.check(lib.b2rl_tc_linear(x_ptr, 256, M, w, lo_of(slot, w), self._p(region, o["b2"]),
                                       self._p(region, o["g2"]) if ln else none, self._p(region, o["be2"]) if ln else none, ln, 1,
                                       H, XH, stat, st), "tc_linear")

        # ---- next action: SAC samples from the ONLINE actor (agent.py:205), TD3 uses the TARGET actor (:194-202)
        ra = RT if ag.td3 else RP
        act = lay.actor
        first(rows.data_ptr() + 4 * (O + A + 2), rs, O, act, ra, self.t1.data_ptr(), none, none)
        hidden(self.t1.data_ptr(), act, ra, self.t2.data_ptr(), none, none, 0)
        p = L.WidePolicy()
        p.h2, p.w3, p.b3 = self.t2.data_ptr(), self._p(ra, act.off["w3"]), self._p(ra, act.off["b3"])
        p.rows, p.min_ac, p.max_ac = rows.data_ptr(), ag.min_ac.data_ptr(), ag.max_ac.data_ptr()
        p.eps, p.eps_out, p.xn = L.ptr(eps), L.ptr(eps_out), self.xn.data_ptr()
        p.logp = None if ag.td3 else self.logp.data_ptr()
        p.counters = ag.counters.data_ptr()
        p.M, p.O, p.A, p.out_dim, p.row_stride, p.ldn, p.src_off = M, O, A, act.out_dim, rs, self.ldn, O + A + 2
        p.td3, p.smoothing = int(ag.td3), int(ag._hyper.targ_smoothing)
        p.counter_idx, p.stream_id = L.CTR_Q, STREAM_CRITIC_EPS
        p.td3_std, p.td3_c, p.seed, p.agent = ag._hyper.td3_std, ag._hyper.td3_c, ag._hyper.seed, ag.agent_id
        L.check(lib.b2rl_wide_policy_head(C.byref(p), st), "wide_policy_head")

        def q_head(net, region, k, mode):
            q = L.WideQ()
            q.h2, q.w3, q.b3 = (self._ws(1, k) if mode else self.t2.data_ptr()), self._p(region, net.off["w3"]), self._p(region, net.off["b3"])
            q.q_out = (self.q if mode else self.qn)[k].data_ptr()
            q.qn0, q.qn1, q.logp = self.qn[0].data_ptr(), self.qn[1].data_ptr(), self.logp.data_ptr()
            q.rows, q.log_alpha = rows.data_ptr(), ag._alpha_state.data_ptr()
            q.dz3, q.sq_part = self._dz3(k), self.sq[k].data_ptr()
            q.targ_out = L.ptr(targ_out) if (mode and k == 0) else None
            q.M, q.mode, q.row_stride, q.rd_off, q.td3, q.bcq_mix = M, mode, rs, O + A, int(ag.td3), int(ag._hyper.bcq_mix)
            q.gamma = ag._hyper.gamma
            L.check(lib.b2rl_wide_q_head(C.byref(q), st), "wide_q_head")

        # ---- twin target Q on (next_obs, a')  (agent.py:208-210)
        for k in range(2):
            net = lay.critic[k]
            first(self.xn.data_ptr(), self.ldn, O + A, net, RT, self.t1.data_ptr(), none, none)
            hidden(self.t1.data_ptr(), net, RT, self.t2.data_ptr(), none, none, 1 + k)
            q_head(net, RT, k, 0)
        # ---- twin online Q, TD target, loss, backward (agent.py:212-235)
        G = self._p(RG, 0)
        for k in range(2):
            net, o = lay.critic[k], lay.critic[k].off
            first(rows.data_ptr(), rs, O + A, net, RP, self._ws(0, k), self.xh1[k].data_ptr(), self.st1[k].data_ptr())
            hidden(self._ws(0, k), net, RP, self._ws(1, k), self.xh2[k].data_ptr(), self.st2[k].data_ptr(), 3 + k)
            q_head(net, RP, k, 1)
            L.check(lib.b2rl_wide_ln_bwd(self._dz3(k), 1, self._p(RP, o["w3"]), self.xh2[k].data_ptr(), self.st2[k].data_ptr(),
                                         self._p(RP, o["g2"]) if ln else none, self._p(RP, o["be2"]) if ln else none, ln, M,
                                         self._ws(3, k), self.part2[k].data_ptr(), st), "wide_ln_bwd")
            w2t = self._p(RP, o["w2t"])
            L.check(lib.b2rl_tc_linear_bwd(self._ws(3, k), M, w2t, lo_of(5 + k, w2t), self.xh1[k].data_ptr(), self.st1[k].data_ptr(),
                                           self._p(RP, o["g1"]) if ln else none, self._p(RP, o["be1"]) if ln else none, ln,
                                           self._ws(2, k), self.part1[k].data_ptr(), st), "tc_linear_bwd")
            L.check(lib.b2rl_wide_colsum(self.part2[k].data_ptr(), self.P128, G, o["b2"], o["g2"] if ln else 0, o["be2"] if ln else 0,
                                         ln, st), "wide_colsum")
            L.check(lib.b2rl_wide_colsum(self.part1[k].data_ptr(), self.P128, G, o["b1"], o["g1"] if ln else 0, o["be1"] if ln else 0,
                                         ln, st), "wide_colsum")
        L.check(lib.b2rl_wide_critic_scalars(self.sq[0].data_ptr(), self.sq[1].data_ptr(), self.P8, self._dz3(0), self._dz3(1), M, G,
                                             lay.critic[0].off["b3"], lay.critic[1].off["b3"], ag.out.data_ptr(), st),
                "wide_critic_scalars")
        # ---- weight gradients (wgrad.cu reads rows / H1 / H2 / DZ1 / DZ2 / DZ3 of the workspace), then Adam
        args = ag.update_args(rows)
        L.check(lib.b2rl_wgrad(C.byref(args), 0, L.CTR_Q, 1, st), "wgrad")
        self._args = args
        if adam:
            ag._launch_adam(ag.critic_segs(False))
        return {"loss/qf_loss": ag.out[L.OUT_QF_LOSS]}
