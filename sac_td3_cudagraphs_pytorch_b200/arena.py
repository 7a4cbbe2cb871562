"""Flat fp32 parameter arena: layout computation and torch views (include/b2rl.h "parameter arena").

One allocation holds, per agent, five identically laid out regions
``online | target | exp_avg | exp_avg_sq | grad``. Networks are contiguous spans of a region, so
"Adam over the critics" or "Polyak over everything" is one flat span for the multi-tensor kernel,
and the nn.Parameters the reference's API exposes (agents/nets.py state_dict names) are strided
views of region 0.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from . import _lib as L

HID = L.HID
TENSOR_NAMES = (  # reference state_dict names (agents/nets.py:66-84) -> arena field
    ("fc_stack.fc_block_1.fc.weight", "w1t"), ("fc_stack.fc_block_1.fc.bias", "b1"),
    ("fc_stack.fc_block_1.ln.weight", "g1"), ("fc_stack.fc_block_1.ln.bias", "be1"),
    ("fc_stack.fc_block_2.fc.weight", "w2t"), ("fc_stack.fc_block_2.fc.bias", "b2"),
    ("fc_stack.fc_block_2.ln.weight", "g2"), ("fc_stack.fc_block_2.ln.bias", "be2"),
    ("head.weight", "w3"), ("head.bias", "b3"),
)


def _up4(n: int) -> int:
    return (n + 3) & ~3


@dataclass
class NetLayout:
    in_dim: int
    out_dim: int
    layer_norm: bool
    off: dict  # field -> float offset inside a region (-1 when absent)
    begin: int
    core_end: int  # end of the tensors the reference knows about (clip_grad_norm_ spans these)
    end: int       # including the w2n shadow

    def c_struct(self) -> L.Net:
        o = self.off
        return L.Net(self.in_dim, self.out_dim, int(self.layer_norm), 0, o["w1t"], o["b1"], o["g1"], o["be1"],
                     o["w2t"], o["b2"], o["g2"], o["be2"], o["w3"], o["b3"], o["w2n"], self.begin, self.end)

    def numel(self, f: str) -> int:
        return {"w1t": self.in_dim * HID, "w2t": HID * HID, "w2n": HID * HID, "w3": self.out_dim * HID,
                "b3": self.out_dim}.get(f, HID)


def _lay_net(cur: int, in_dim: int, out_dim: int, layer_norm: bool) -> tuple[NetLayout, int]:
    begin = cur
    off = {}
    for f, n in (("w1t", in_dim * HID), ("b1", HID), ("g1", HID), ("be1", HID), ("w2t", HID * HID), ("b2", HID),
                 ("g2", HID), ("be2", HID), ("w3", out_dim * HID), ("b3", out_dim)):
        if f in ("g1", "be1", "g2", "be2") and not layer_norm:
            off[f] = -1
            continue
        off[f] = cur
        cur += _up4(n)
    core_end = cur
    off["w2n"] = cur
    cur += HID * HID
    return NetLayout(in_dim, out_dim, layer_norm, off, begin, core_end, cur), cur


@dataclass
class ArenaLayout:
    critic: list  # two NetLayout with identical internal structure
    actor: NetLayout
    region: int   # floats per region (multiple of 4)

    @property
    def critic_stride(self) -> int:
        return self.critic[1].begin - self.critic[0].begin


def make_layout(ob_dim: int, ac_dim: int, td3: bool, layer_norm: bool) -> ArenaLayout:
    cur = 0
    c0, cur = _lay_net(cur, ob_dim + ac_dim, 1, layer_norm)
    c1, cur = _lay_net(cur, ob_dim + ac_dim, 1, layer_norm)
    a, cur = _lay_net(cur, ob_dim, ac_dim if td3 else 2 * ac_dim, layer_norm)
    return ArenaLayout([c0, c1], a, _up4(cur))


def set_shadow_pairs(adam_args, layout: ArenaLayout) -> None:
    """Tell an Adam / Polyak launch where every net's (w2t, w2n) pair lives (b2rl_adam_args_t.shadow_*): the step is
    computed at w2t and written to both layouts, the shadow's own gradient / moments are never read."""
    nets = (*layout.critic, layout.actor)
    for i, net in enumerate(nets):
        adam_args.shadow_src[i], adam_args.shadow_dst[i] = net.off["w2t"], net.off["w2n"]
    adam_args.n_shadow = len(nets)


class Arena:
    """The device allocation + view helpers. ``n_agents`` > 1 stacks independent learners."""

    def __init__(self, layout: ArenaLayout, device, n_agents: int = 1):
        self.layout, self.n_agents = layout, n_agents
        self.flat = torch.zeros(n_agents, 5, layout.region, dtype=torch.float32, device=device)
        assert self.flat.data_ptr() % 16 == 0

    @property
    def agent_stride(self) -> int:
        return 5 * self.layout.region

    def region(self, r: int, agent: int = 0) -> torch.Tensor:
        return self.flat[agent, r]

    def tensor(self, net: NetLayout, field: str, region: int = L.REGION_P, agent: int = 0) -> torch.Tensor:
        """View with the reference's shape: [out,in] weights (a transposed view for w1t/w2t), [n] vectors."""
        o = net.off[field]
        flat = self.flat[agent, region, o:o + net.numel(field)]
        if field == "w1t":
            return flat.view(net.in_dim, HID).t()
        if field == "w2t":
            return flat.view(HID, HID).t()
        if field == "w2n":
            return flat.view(HID, HID)
        if field == "w3":
            return flat.view(net.out_dim, HID)
        return flat

    def named(self, net: NetLayout, region: int = L.REGION_P, agent: int = 0) -> dict[str, torch.Tensor]:
        return {name: self.tensor(net, f, region, agent) for name, f in TENSOR_NAMES if net.off[f] >= 0}

    def stacked(self, region: int = L.REGION_P, agent: int = 0) -> dict[str, torch.Tensor]:
        """Twin critics as [2, ...] views, the shape of the reference's ``qnet_params`` (agent.py:106)."""
        c0, stride = self.layout.critic[0], self.layout.critic_stride
        out = {}
        for name, f in TENSOR_NAMES:
            if c0.off[f] < 0:
                continue
            v = self.tensor(c0, f, region, agent)
            out[name] = torch.as_strided(v, (2, *v.shape), (stride, *v.stride()), v.storage_offset())
        return out

    @torch.no_grad()
    def sync_shadows(self, regions=(L.REGION_P, L.REGION_T, L.REGION_M, L.REGION_V)) -> None:
        """w2n <- transpose(w2t) after any host-side write to the weights / optimizer state."""
        for a in range(self.n_agents):
            for r in regions:
                for net in (*self.layout.critic, self.layout.actor):
                    self.tensor(net, "w2n", r, a).copy_(self.tensor(net, "w2t", r, a))
