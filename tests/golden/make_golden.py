#!/usr/bin/env python
"""Generate tests/golden/<case>.npz by RUNNING THE REFERENCE ITSELF.

Authoring-container only (needs /root/reference). The reference's unmodified
``agents/agent.py`` + ``agents/nets.py`` are imported; the three packages it needs that
are not installed here (tensordict, torchrl, omegaconf) are replaced by the container
stand-ins in ``_ref_shims/`` (no arithmetic of their own). Its random draws are
redirected to the portable noise of ``portable.py`` by patching the two torch entry
points it samples through (``Normal.rsample`` -> ``_standard_normal``; ``Tensor.normal_``),
so the same noise can be fed to the oracle and to the CUDA kernels.

The learner iterations follow orchestrator.py:337-352. For every case the oracle
(``oracle/sac_td3_oracle.py``) is run on the same inputs and its maximum deviation
from the reference over ALL tensors is stored in the fixture's ``meta``.

    python -m tests.golden.make_golden            # all cases
"""
from __future__ import annotations

import contextlib
import json
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF = Path("/root/reference")

sys.path.insert(0, str(REPO))
from tests.golden import portable as P  # noqa: E402
from tests.golden.cases import CASES, case_inputs, inputs_digest  # noqa: E402


def _import_reference():
    assert REF.exists(), "the reference is only mounted in the authoring container"
    sys.path.insert(0, str(HERE / "_ref_shims"))
    sys.path.insert(0, str(REF))
    from agents.agent import Agent  # the reference, unmodified
    from omegaconf import DictConfig
    from tensordict import TensorDict
    return Agent, DictConfig, TensorDict


@contextlib.contextmanager
def injected_noise(queue: list):
    """Serve the reference's N(0,1) draws from ``queue`` (FIFO of [B,A] tensors)."""
    import torch.distributions.normal as tdn
    orig_sn, orig_normal_ = tdn._standard_normal, torch.Tensor.normal_

    def sn(shape, dtype, device):
        z = queue.pop(0)
        assert tuple(z.shape) == tuple(shape), (z.shape, shape)
        return z.to(dtype=dtype, device=device)

    def normal_(self, mean=0.0, std=1.0, *, generator=None):
        z = queue.pop(0)
        return self.copy_(z * std + mean)

    tdn._standard_normal, torch.Tensor.normal_ = sn, normal_
    try:
        yield
    finally:
        tdn._standard_normal, torch.Tensor.normal_ = orig_sn, orig_normal_


def run_reference(inp, Agent, DictConfig, TensorDict):
    h = inp["hps"]
    agent = Agent(net_shapes={"ob_shape": (inp["ob"],), "ac_shape": (inp["ac"],)},
                  min_ac=inp["min_ac"], max_ac=inp["max_ac"], device=torch.device("cpu"),
                  hps=DictConfig(h), rb=None)
    with torch.no_grad():
        for n, v in inp["actor"].items():
            agent.actor_params[n].data.copy_(v)
            agent.actor_target[n].copy_(v)
        for n in inp["q1"]:
            st = torch.stack([inp["q1"][n], inp["q2"][n]])
            agent.qnet_params[n].data.copy_(st)
            agent.qnet_target[n].copy_(st)
    td3 = h["prefer_td3_over_sac"]
    logs, rec = [], {}
    for i in range(inp["iters"]):
        batch = TensorDict({k: v[inp["idx"][i]] for k, v in inp["storage"].items()})
        q = []
        if (not td3) or h.get("targ_actor_smoothing", False):
            q.append(inp["eps_q"][i])
        with injected_noise(q):
            out = dict(agent.update_qnets(batch).items())
        assert not q
        agent.qnet_updates_so_far += 1
        if i == 0:
            rec["grad_q"] = {n: p.grad.detach().clone() for n, p in agent.qnet.named_parameters()}
        if i % (h["actor_update_delay"] + 1) == 0:
            for j in range(h["actor_update_delay"]):
                q = []
                if not td3:
                    q.append(inp["eps_pi"][i][j])
                    if h["autotune"]:
                        q.append(inp["eps_alpha"][i][j])
                with injected_noise(q):
                    out.update(dict(agent.update_actor(batch).items()))
                assert not q
                agent.actor_updates_so_far += 1
                if i == 0 and j == 0:
                    rec["grad_actor"] = {n: p.grad.detach().clone() for n, p in agent.actor.named_parameters()}
        agent.update_targ_nets()
        logs.append([float(out.get(k, float("nan"))) for k in LOG_KEYS])
    pnames = list(inp["actor"].keys())
    rec["actor"] = {n: agent.actor_params[n].detach().clone() for n in pnames}
    rec["actor_target"] = {n: agent.actor_target[n].detach().clone() for n in pnames}
    rec["qnet"] = {n: agent.qnet_params[n].detach().clone() for n in inp["q1"]}
    rec["qnet_target"] = {n: agent.qnet_target[n].detach().clone() for n in inp["q1"]}
    if not td3:
        rec["log_alpha"] = {"log_alpha": agent.log_alpha.detach().clone().reshape(1)}
    rec["logs"] = np.asarray(logs, dtype=np.float64)
    return rec


LOG_KEYS = ("loss/qf_loss", "loss/actor_loss", "loss/alpha_loss", "vitals/alpha")


def run_oracle(inp, dtype=torch.float32, capturable=False):
    """Same protocol on the oracle. Also used by tests/test_oracle_golden.py."""
    from oracle import OracleAgent, OracleHps
    h = inp["hps"]
    fields = OracleHps.__dataclass_fields__
    hps = OracleHps(**{k: v for k, v in h.items() if k in fields}, adam_capturable=capturable)
    cast = lambda d: {k: v.to(dtype) for k, v in d.items()}
    ag = OracleAgent(inp["ob"], inp["ac"], inp["min_ac"], inp["max_ac"], hps, dtype=dtype,
                     actor_init=cast(inp["actor"]), qnet_init=[cast(inp["q1"]), cast(inp["q2"])])
    logs, rec = [], {}
    for i in range(inp["iters"]):
        batch = {k: (v[inp["idx"][i]].to(dtype) if v.is_floating_point() else v[inp["idx"][i]])
                 for k, v in inp["storage"].items()}
        out = dict(ag.update_qnets(batch, inp["eps_q"][i]))
        ag.qnet_updates_so_far += 1
        if i == 0:
            rec["grad_q"] = {n: p.grad.detach().clone() for n, p in ag.qnet.items()}
        if i % (h["actor_update_delay"] + 1) == 0:
            for j in range(h["actor_update_delay"]):
                out.update(ag.update_actor(batch, inp["eps_pi"][i][j], inp["eps_alpha"][i][j]))
                ag.actor_updates_so_far += 1
                if i == 0 and j == 0:
                    rec["grad_actor"] = {n: p.grad.detach().clone() for n, p in ag.actor.items()}
        ag.update_targ_nets()
        logs.append([float(out.get(k, float("nan"))) for k in LOG_KEYS])
    rec["actor"] = {n: v.detach().clone() for n, v in ag.actor.items()}
    rec["actor_target"] = {n: v.detach().clone() for n, v in ag.actor_target.items()}
    rec["qnet"] = {n: v.detach().clone() for n, v in ag.qnet.items()}
    rec["qnet_target"] = {n: v.detach().clone() for n, v in ag.qnet_target.items()}
    if not ag.td3:
        rec["log_alpha"] = {"log_alpha": ag.log_alpha.detach().clone().reshape(1)}
    rec["logs"] = np.asarray(logs, dtype=np.float64)
    return rec


GROUPS = ("grad_q", "grad_actor", "actor", "actor_target", "qnet", "qnet_target", "log_alpha")


def max_deviation(a, b) -> float:
    """max over tensors of max|a-b| / max|b| (the tolerance definition of BASELINE.md §4.6)."""
    worst = 0.0
    for g in GROUPS:
        if g not in b:
            continue
        for n, tb in b[g].items():
            ta = a[g][n].to(torch.float64)
            tb = tb.to(torch.float64)
            denom = max(float(tb.abs().max()), 1e-30)
            worst = max(worst, float((ta - tb).abs().max()) / denom)
    la, lb = a["logs"], b["logs"]
    m = ~np.isnan(lb)
    worst = max(worst, float(np.max(np.abs(la[m] - lb[m]) / np.maximum(np.abs(lb[m]), 1e-30))))
    return worst


def main(names=None):
    Agent, DictConfig, TensorDict = _import_reference()
    torch.set_num_threads(1)
    for name in (names or CASES):
        inp = case_inputs(name)
        torch.manual_seed(0)
        ref = run_reference(inp, Agent, DictConfig, TensorDict)
        orc = run_oracle(inp)
        dev = max_deviation(orc, ref)
        orc64 = run_oracle(inp, dtype=torch.float64)
        dev64 = max_deviation(ref, orc64)
        print(f"{name:32s} oracle-vs-reference max rel dev {dev:.3e}   reference-vs-fp64 {dev64:.3e}")
        assert dev <= 1e-6, f"oracle does not restate the reference for {name}: {dev}"
        out = {"logs": ref["logs"]}
        for g in GROUPS:
            for n, t in ref.get(g, {}).items():
                out[f"{g}/{n}"] = P.summarize(t)
        out["meta"] = np.frombuffer(json.dumps({
            "case": name, "inputs_sha256": inputs_digest(inp), "torch": torch.__version__, "numpy": np.__version__,
            "oracle_vs_reference_max_rel_dev": dev, "reference_fp32_vs_fp64_oracle": dev64,
            "source": "reference agents/agent.py + agents/nets.py run under tests/golden/_ref_shims",
            "log_keys": LOG_KEYS,
        }).encode(), dtype=np.uint8)
        np.savez_compressed(HERE / f"{name}.npz", **out)


if __name__ == "__main__":
    main(sys.argv[1:] or None)
