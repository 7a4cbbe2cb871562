"""Platform-independent synthetic inputs for the golden fixtures and the parity tests.

Everything is derived from numpy PCG64 *integer* draws scaled by powers of two, so
the same seed yields bit-identical float32 tensors on any machine — the fixtures in
this directory therefore store only the *outputs* of the reference, not its inputs.
Test infrastructure; not imported by the product.
"""
from __future__ import annotations

import numpy as np
import torch

HID = 256
PARAM_SHAPES = (  # name, shape as a function of (in_dim, out_dim), kind
    ("fc_stack.fc_block_1.fc.weight", lambda i, o: (HID, i), "w"),
    ("fc_stack.fc_block_1.fc.bias", lambda i, o: (HID,), "b"),
    ("fc_stack.fc_block_1.ln.weight", lambda i, o: (HID,), "g"),
    ("fc_stack.fc_block_1.ln.bias", lambda i, o: (HID,), "b"),
    ("fc_stack.fc_block_2.fc.weight", lambda i, o: (HID, HID), "w"),
    ("fc_stack.fc_block_2.fc.bias", lambda i, o: (HID,), "b"),
    ("fc_stack.fc_block_2.ln.weight", lambda i, o: (HID,), "g"),
    ("fc_stack.fc_block_2.ln.bias", lambda i, o: (HID,), "b"),
    ("head.weight", lambda i, o: (o, HID), "w"),
    ("head.bias", lambda i, o: (o,), "b"),
)


def unit(rng: np.random.Generator, shape) -> np.ndarray:
    """Uniform on the grid k/32768, k in [-32768, 32768) — exact in float32."""
    return (rng.integers(-32768, 32768, size=shape, dtype=np.int64).astype(np.float32)
            / np.float32(32768.0))


def pseudo_normal(rng: np.random.Generator, shape) -> np.ndarray:
    """Irwin-Hall(4) scaled to unit variance: exact dyadic sum times one constant."""
    s = sum(unit(rng, shape) for _ in range(4))
    return (s * np.float32(np.sqrt(3.0 / 4.0))).astype(np.float32)


def mlp_params(seed: int, in_dim: int, out_dim: int, layer_norm: bool,
               head_scale: float = 1.0) -> dict[str, torch.Tensor]:
    """Non-trivial parameters (biases and LN affine not at their init values).
    ``head_scale`` shrinks the head so a tanh policy is not saturated (a saturated
    tanh makes the reference's own fp32 log-prob ill-conditioned: 1 - y^2 cancels)."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shp, kind in PARAM_SHAPES:
        shape = shp(in_dim, out_dim)
        u = unit(rng, shape)
        if kind == "w":
            v = u * np.float32(1.5 / np.sqrt(max(shape)))
            if name == "head.weight":
                v = v * np.float32(head_scale)
        elif kind == "g":
            v = np.float32(1.0) + np.float32(0.25) * u
        else:
            v = np.float32(0.125) * u
        if ".ln." in name and not layer_norm:
            continue
        out[name] = torch.from_numpy(v.astype(np.float32))
    return out


def transitions(seed: int, n: int, ob_dim: int, ac_dim: int, min_ac, max_ac,
                done_prob: float = 0.05) -> dict[str, torch.Tensor]:
    rng = np.random.default_rng(seed)
    lo = np.asarray(min_ac, dtype=np.float32)
    hi = np.asarray(max_ac, dtype=np.float32)
    obs = np.float32(2.0) * unit(rng, (n, ob_dim))
    nxt = obs + np.float32(0.125) * unit(rng, (n, ob_dim))
    act = lo + (hi - lo) * ((unit(rng, (n, ac_dim)) + np.float32(1.0)) * np.float32(0.5))
    rew = unit(rng, (n, 1))
    done = rng.integers(0, 1000, size=(n, 1)) < int(done_prob * 1000)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    return {"observations": t(obs), "next_observations": t(nxt), "actions": t(act.astype(np.float32)),
            "rewards": t(rew), "terminations": t(done), "dones": t(done.copy())}


def indices(seed: int, n: int, batch: int) -> torch.Tensor:
    rng = np.random.default_rng(seed)
    return torch.from_numpy(rng.integers(0, n, size=(batch,), dtype=np.int64))


def noise(seed: int, batch: int, ac_dim: int) -> torch.Tensor:
    return torch.from_numpy(pseudo_normal(np.random.default_rng(seed), (batch, ac_dim)))


# ---------------------------------------------------------------- summaries
N_SAMPLES = 64


def summarize(t) -> np.ndarray:
    """[sum, sum|x|, sum x^2, max|x|] in float64 followed by N_SAMPLES strided elements."""
    a = np.asarray(t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else t, dtype=np.float64).ravel()
    pos = np.linspace(0, a.size - 1, N_SAMPLES).astype(np.int64)
    stats = np.array([a.sum(), np.abs(a).sum(), (a * a).sum(), np.abs(a).max() if a.size else 0.0])
    return np.concatenate([stats, a[pos]])


def summary_close(got, want, rtol) -> tuple[bool, float]:
    """Compare two summaries; every entry is measured relative to the tensor's scale:
    samples and max against max|x|, the three sums against sum|x| (resp. sum x^2)."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    amax = max(abs(want[3]), 1e-30)
    errs = [abs(got[0] - want[0]) / max(want[1], 1e-30),
            abs(got[1] - want[1]) / max(want[1], 1e-30),
            abs(got[2] - want[2]) / max(want[2], 1e-30),
            abs(got[3] - want[3]) / amax]
    errs.append(np.abs(got[4:] - want[4:]).max() / amax)
    e = float(max(errs))
    return e <= rtol, e
