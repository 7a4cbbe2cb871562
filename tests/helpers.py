"""Shared helpers for the parity tests (test infrastructure)."""
from __future__ import annotations

import numpy as np
import torch

from oracle import OracleAgent, OracleHps
from tests.golden.cases import case_inputs


def rel_dev(a, b, scale=None) -> float:
    """max|a-b| / max|b| — the per-tensor tolerance definition of BASELINE.md §4.6.
    ``scale`` replaces the denominator for scalars that are a difference of large terms."""
    a = torch.as_tensor(a).detach().to("cpu", torch.float64)
    b = torch.as_tensor(b).detach().to("cpu", torch.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    denom = float(scale) if scale is not None else max(float(b.abs().max()), 1e-30)
    return float((a - b).abs().max()) / denom


def oracle_hps(h: dict, capturable: bool = False) -> OracleHps:
    f = OracleHps.__dataclass_fields__
    return OracleHps(**{k: v for k, v in h.items() if k in f}, adam_capturable=capturable)


def make_oracle(inp, dtype=torch.float32, capturable=True) -> OracleAgent:
    cast = lambda d: {k: v.to(dtype) for k, v in d.items()}
    return OracleAgent(inp["ob"], inp["ac"], inp["min_ac"], inp["max_ac"], oracle_hps(inp["hps"], capturable),
                       dtype=dtype, actor_init=cast(inp["actor"]), qnet_init=[cast(inp["q1"]), cast(inp["q2"])])


def make_agent(inp, device="cuda", seed=0):
    from sac_td3_cudagraphs_pytorch_b200 import Hps
    from sac_td3_cudagraphs_pytorch_b200.agents.agent import Agent
    ag = Agent({"ob_shape": (inp["ob"],), "ac_shape": (inp["ac"],)}, inp["min_ac"], inp["max_ac"],
               torch.device(device), Hps(inp["hps"]), rb=None, seed=seed)
    ag.load_params(inp["actor"], inp["q1"], inp["q2"])
    return ag


def batch_of(inp, i, dtype=torch.float32, device="cpu"):
    idx = inp["idx"][i]
    return {k: (v[idx].to(dtype) if v.is_floating_point() else v[idx]).to(device) for k, v in inp["storage"].items()}


def alpha_loss_scale(alpha, ac_dim) -> float:
    """alpha_loss = alpha * (mean(-logpi) - targ_ent) with targ_ent = -A is a difference of two O(A)
    terms (it crosses zero as the entropy reaches its target), so its error is judged against
    alpha * A, the size of the terms, not against the cancelled result."""
    return float(alpha) * float(ac_dim)


def check_close(name, got, ref32, ref64, floor=1e-5, factor=4.0, scale=None):
    """CUDA result vs the fp32 oracle, with the fp32 oracle's own distance to float64 as yardstick:
    dev(cuda, fp32) <= max(floor, factor * dev(fp32, fp64)). Returns the three numbers for reporting."""
    d_cuda = rel_dev(got, ref32, scale)
    d_ref = rel_dev(ref32, ref64, scale)
    d_true = rel_dev(got, ref64, scale)
    tol = max(floor, factor * d_ref)
    assert d_cuda <= tol or d_true <= tol, (
        f"{name}: cuda-vs-oracle32 {d_cuda:.3e}, cuda-vs-oracle64 {d_true:.3e}, "
        f"oracle32-vs-oracle64 {d_ref:.3e}, tol {tol:.3e}")
    return d_cuda, d_true, d_ref
