"""Actor / twin-critic network containers with the reference's constructor signatures and
``state_dict`` names (agents/nets.py:52-234), re-designed as *views*: after ``bind()`` every
parameter is a strided window into the flat fp32 arena that the CUDA kernels read and the fused
Adam/Polyak kernel writes, so torch code (checkpointing, the inference policy, tests) and the
kernels always see the same bytes.

The ``forward`` methods are the torch statement of the same maths; the learner update never
calls them (it runs csrc/critic.cu, csrc/actor.cu). Shapes: hidden width is 256 for both
layers (agents/agent.py:56,101); Linear -> LayerNorm|Identity -> ReLU twice, then a head.
"""
from __future__ import annotations

import math
from typing import Callable, Union

import torch
from torch import nn

SAC_LOG_STD_BOUNDS = [-5.0, 2.0]  # agents/nets.py:13

Device = Union[str, torch.device]


def init(constant_bias: float = 0.0) -> Callable[[nn.Module], None]:
    """Initializer factory: orthogonal weights / constant bias for Linear-like layers, identity
    affine for normalisation layers (agents/nets.py:34-49)."""
    linear_like = (nn.Linear, nn.Conv2d, nn.Bilinear)
    norm_like = (nn.LayerNorm, nn.BatchNorm2d)

    def apply(m: nn.Module) -> None:
        if isinstance(m, linear_like):
            nn.init.orthogonal_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, constant_bias)
        elif isinstance(m, norm_like):
            nn.init.ones_(m.weight)
            if m.bias is not None:
                nn.init.zeros_(m.bias)

    return apply


def log_module_info(model: nn.Module) -> str:
    n = sum(p.numel() for p in model.parameters())
    unit = f"{n / 1e6:.2f} M" if n >= 10 ** 6 else (f"{n / 1e3:.2f} k" if n >= 10 ** 3 else str(n))
    return f"{model.__class__.__name__}: {unit} params"


def _block(n_in: int, n_out: int, layer_norm: bool, device: Device) -> nn.Sequential:
    blk = nn.Sequential()
    blk.add_module("fc", nn.Linear(n_in, n_out, device=device))
    blk.add_module("ln", nn.LayerNorm(n_out, device=device) if layer_norm else nn.Identity())
    blk.add_module("nl", nn.ReLU())
    return blk


class _Mlp(nn.Module):
    """Shared trunk + head; subclasses only differ in what they feed in and how they read the head."""

    def __init__(self, in_dim: int, out_dim: int, hid_dims: tuple[int, int], layer_norm: bool, device: Device):
        super().__init__()
        self.layer_norm = layer_norm
        self.fc_stack = nn.Sequential()
        self.fc_stack.add_module("fc_block_1", _block(in_dim, hid_dims[0], layer_norm, device))
        self.fc_stack.add_module("fc_block_2", _block(hid_dims[0], hid_dims[1], layer_norm, device))
        self.head = nn.Linear(hid_dims[1], out_dim, device=device)
        if str(device) != "meta":
            self.fc_stack.apply(init())
            self.head.apply(init())

    def trunk_head(self, x: torch.Tensor) -> torch.Tensor:
        return self.head(self.fc_stack(x))

    @torch.no_grad()
    def bind(self, views: dict[str, torch.Tensor], grads: dict[str, torch.Tensor] | None = None,
             requires_grad: bool = True, copy_from_self: bool = False) -> None:
        """Re-point every parameter at ``views[name]`` (arena windows). With ``copy_from_self`` the
        current values are written into the arena first."""
        for name, view in views.items():
            *path, leaf = name.split(".")
            mod = self
            for p in path:
                mod = getattr(mod, p)
            if copy_from_self:
                view.copy_(getattr(mod, leaf).detach().to(view.device))
            param = nn.Parameter(view, requires_grad=requires_grad)
            if grads is not None:
                param.grad = grads[name]
            mod._parameters[leaf] = param


class Critic(_Mlp):
    """Q(ob, ac) -> [B, 1]  (agents/nets.py:52-92)."""

    def __init__(self, ob_shape: tuple[int, ...], ac_shape: tuple[int, ...], hid_dims: tuple[int, int], *,
                 layer_norm: bool, device: Device):
        super().__init__(ob_shape[-1] + ac_shape[-1], 1, hid_dims, layer_norm, device)

    def forward(self, ob: torch.Tensor, ac: torch.Tensor) -> torch.Tensor:
        return self.trunk_head(torch.cat([ob, ac], dim=-1))


class _Policy(_Mlp):
    def __init__(self, ob_dim, out_dim, hid_dims, min_ac, max_ac, layer_norm, device):
        super().__init__(ob_dim, out_dim, hid_dims, layer_norm, device)
        self.register_buffer("action_scale", (max_ac - min_ac) / 2.0)
        self.register_buffer("action_bias", (max_ac + min_ac) / 2.0)

    def squash(self, x: torch.Tensor) -> torch.Tensor:
        return torch.tanh(x) * self.action_scale + self.action_bias


class Actor(_Policy):
    """TD3 deterministic policy with additive exploration noise (agents/nets.py:95-159)."""

    def __init__(self, ob_shape: tuple[int, ...], ac_shape: tuple[int, ...], hid_dims: tuple[int, int],
                 min_ac: torch.Tensor, max_ac: torch.Tensor, *, exploration_noise: float, layer_norm: bool,
                 device: Device):
        super().__init__(ob_shape[-1], ac_shape[-1], hid_dims, min_ac, max_ac, layer_norm, device)
        self.register_buffer("exploration_noise", torch.as_tensor(exploration_noise, device=device))

    def forward(self, ob: torch.Tensor) -> torch.Tensor:
        return self.squash(self.trunk_head(ob))

    def exploit(self, ob: torch.Tensor) -> dict[str, torch.Tensor]:
        return {"action": self(ob)}

    def explore(self, ob: torch.Tensor) -> dict[str, torch.Tensor]:
        ac = self(ob)
        return {"action": ac + torch.randn_like(ac) * (self.action_scale * self.exploration_noise)}


class TanhGaussActor(_Policy):
    """SAC tanh-Gaussian policy, log-std squashed into SAC_LOG_STD_BOUNDS (agents/nets.py:162-234)."""

    def __init__(self, ob_shape: tuple[int, ...], ac_shape: tuple[int, ...], hid_dims: tuple[int, int],
                 min_ac: torch.Tensor, max_ac: torch.Tensor, *, layer_norm: bool, device: Device):
        super().__init__(ob_shape[-1], 2 * ac_shape[-1], hid_dims, min_ac, max_ac, layer_norm, device)

    @staticmethod
    def bound_log_std(log_std: torch.Tensor) -> torch.Tensor:
        lo, hi = SAC_LOG_STD_BOUNDS
        return lo + 0.5 * (hi - lo) * (torch.tanh(log_std) + 1)

    def forward(self, ob: torch.Tensor) -> tuple[torch.Tensor, torch.Tensor]:
        mean, log_std = self.trunk_head(ob).chunk(2, dim=-1)
        return mean, self.bound_log_std(log_std).exp()

    def get_action(self, ob: torch.Tensor, eps: torch.Tensor | None = None) -> dict[str, torch.Tensor]:
        mean, std = self(ob)
        x_t = mean + std * (torch.randn_like(mean) if eps is None else eps)
        y_t = torch.tanh(x_t)
        log_prob = (-((x_t - mean) ** 2) / (2 * std ** 2) - std.log() - math.log(math.sqrt(2 * math.pi))
                    - torch.log(self.action_scale * (1 - y_t.pow(2)) + 1e-6)).sum(1, keepdim=True)
        return {"sample": y_t * self.action_scale + self.action_bias, "log_prob": log_prob,
                "mode": self.squash(mean)}
