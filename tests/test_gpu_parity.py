"""GPU parity: the CUDA path (through the C ABI, via the Agent mirror) against the oracle.

Bars (north_star): gathers bit-exact; loss / gradients / parameters within 1e-5
relative (max|a-b| / max|b| per tensor) in fp32 — with one honest qualification measured in
tests/golden/make_golden.py: two fp32 evaluations of the same update that differ only in summation
order already sit up to `d32_64` apart (the reference's own fp32-vs-float64 gap, 3e-6 .. 2e-2
depending on conditioning), so the bound used is max(1e-5, 4 x d32_64), and d32_64 is printed.
"""
import json
from pathlib import Path

import numpy as np
import pytest
import torch

from tests.golden import portable as P
from tests.golden.cases import CASES, RAGGED_CASES, case_inputs
from tests.golden.make_golden import GROUPS, LOG_KEYS
from tests.helpers import (alpha_loss_scale, batch_of, check_close, check_param_after_first_adam, make_agent,
                           make_oracle, rel_dev)

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


def _dev(x):
    return x.to("cuda")


# ------------------------------------------------------------------------------- single steps
ALL_CASES = {**CASES, **RAGGED_CASES}


@pytest.mark.parametrize("name", list(ALL_CASES))
def test_critic_step_matches_oracle(name):
    inp = case_inputs(ALL_CASES[name])
    ag = make_agent(inp)
    o32, o64 = make_oracle(inp, torch.float32), make_oracle(inp, torch.float64)
    B = inp["B"]
    tq = torch.zeros(B, device="cuda")
    qv = torch.zeros(2, B, device="cuda")
    out = ag.update_qnets({k: _dev(v) for k, v in batch_of(inp, 0).items()}, eps=_dev(inp["eps_q"][0]),
                          dbg_targ_q=tq, dbg_q=qv)
    r32 = o32.update_qnets(batch_of(inp, 0), inp["eps_q"][0])
    r64 = o64.update_qnets(batch_of(inp, 0, torch.float64), inp["eps_q"][0].double())
    torch.cuda.synchronize()
    check_close("targ_q", tq, r32["_targ_q"], r64["_targ_q"])
    check_close("q", qv, r32["_q"], r64["_q"])
    check_close("qf_loss", out["loss/qf_loss"], r32["loss/qf_loss"], r64["loss/qf_loss"])
    for n, p in ag.qnet_params.items():
        check_close(f"grad {n}", p.grad, o32.qnet[n].grad, o64.qnet[n].grad)
    for n, p in ag.qnet_params.items():  # after the Adam step
        check_param_after_first_adam(f"param {n}", p, o32.qnet[n], o64.qnet[n], p.grad, o32.qnet[n].grad,
                                     float(inp["hps"]["qnets_lr"]))
    # the natural-layout shadow of fc2.weight stays bit-identical to the primary copy
    for k, q in enumerate((ag.qnet1, ag.qnet2)):
        w = q.fc_stack.fc_block_2.fc.weight.detach()
        assert torch.equal(ag.arena.tensor(ag.layout.critic[k], "w2n"), w.contiguous())


@pytest.mark.parametrize("name", list(ALL_CASES))
def test_actor_step_matches_oracle(name):
    inp = case_inputs(ALL_CASES[name])
    ag = make_agent(inp)
    o32, o64 = make_oracle(inp, torch.float32), make_oracle(inp, torch.float64)
    e1, e2 = inp["eps_pi"][0][0], inp["eps_alpha"][0][0]
    out = ag.update_actor({k: _dev(v) for k, v in batch_of(inp, 0).items()}, eps=_dev(e1), eps_alpha=_dev(e2))
    r32 = o32.update_actor(batch_of(inp, 0), e1, e2)
    r64 = o64.update_actor(batch_of(inp, 0, torch.float64), e1.double(), e2.double())
    torch.cuda.synchronize()
    for k in r32:
        sc = alpha_loss_scale(inp["hps"]["alpha_init"], inp["ac"]) if k == "loss/alpha_loss" else None
        check_close(k, out[k], r32[k], r64[k], scale=sc)
    for n, p in ag.actor_params.items():
        check_close(f"grad {n}", p.grad, o32.actor[n].grad, o64.actor[n].grad)
        check_param_after_first_adam(f"param {n}", p, o32.actor[n], o64.actor[n], p.grad, o32.actor[n].grad,
                                     float(inp["hps"]["actor_lr"]))
    if not ag.td3:
        check_close("log_alpha", ag.log_alpha, o32.log_alpha, o64.log_alpha)
    w = ag.actor.fc_stack.fc_block_2.fc.weight.detach()
    assert torch.equal(ag.arena.tensor(ag.layout.actor, "w2n"), w.contiguous())


# ------------------------------------------------------------------------------- whole protocols
def run_agent_protocol(inp, ag):
    """The golden protocol (orchestrator.py:337-352 cadence) through the Agent API."""
    h = inp["hps"]
    logs, rec = [], {}
    for i in range(inp["iters"]):
        batch = {k: _dev(v) for k, v in batch_of(inp, i).items()}
        out = dict(ag.update_qnets(batch, eps=_dev(inp["eps_q"][i])))
        ag.qnet_updates_so_far += 1
        if i == 0:
            rec["grad_q"] = {n: p.grad.detach().clone() for n, p in ag.qnet_params.items()}
        if i % (h["actor_update_delay"] + 1) == 0:
            for j in range(h["actor_update_delay"]):
                out.update(ag.update_actor(batch, eps=_dev(inp["eps_pi"][i][j]), eps_alpha=_dev(inp["eps_alpha"][i][j])))
                ag.actor_updates_so_far += 1
                if i == 0 and j == 0:
                    rec["grad_actor"] = {n: p.grad.detach().clone() for n, p in ag.actor_params.items()}
        ag.update_targ_nets()
        logs.append([float(out[k]) if k in out else float("nan") for k in LOG_KEYS])
    rec["actor"] = {n: p.detach().clone() for n, p in ag.actor_params.items()}
    rec["actor_target"] = {n: p.detach().clone() for n, p in ag.actor_target.items()}
    rec["qnet"] = {n: p.detach().clone() for n, p in ag.qnet_params.items()}
    rec["qnet_target"] = {n: p.detach().clone() for n, p in ag.qnet_target.items()}
    if not ag.td3:
        rec["log_alpha"] = {"log_alpha": ag.log_alpha.detach().clone().reshape(1)}
    rec["logs"] = np.asarray(logs, dtype=np.float64)
    return rec


@pytest.mark.parametrize("name", list(CASES))
def test_protocol_matches_reference_fixture(name):
    """CUDA path vs the outputs recorded from the reference's own agents/agent.py."""
    z = np.load(GOLD / f"{name}.npz")
    meta = json.loads(bytes(z["meta"]).decode())
    inp = case_inputs(name)
    rec = run_agent_protocol(inp, make_agent(inp))
    # several Adam steps amplify summation-order noise (the first steps move every weight by ~lr*sign(g), so
    # a gradient that is pure rounding noise flips whole steps): allow 25x the reference's own
    # fp32-vs-float64 gap on the same trajectory (recorded in the fixture); single-step tests are tight
    # — capped at 2e-3 so that the bound can fail on the ill-conditioned cases too (sac_saturated's own gap is 2e-2);
    # achieved deviations: profiles/r2_parity_report.txt (worst 6e-4, sac_clip_targfreq2; sac_saturated 2e-5)
    tol = min(max(2e-5, 25 * meta["reference_fp32_vs_fp64_oracle"]), 2e-3)
    want = z["logs"]
    m = ~np.isnan(want)
    assert (np.isnan(rec["logs"]) == np.isnan(want)).all()
    denom = np.maximum(np.abs(want), 1e-30)
    if not inp["hps"]["prefer_td3_over_sac"]:  # alpha_loss column: see helpers.alpha_loss_scale
        denom[:, LOG_KEYS.index("loss/alpha_loss")] = np.maximum(
            denom[:, LOG_KEYS.index("loss/alpha_loss")], alpha_loss_scale(inp["hps"]["alpha_init"], inp["ac"]))
    rel = np.abs(rec["logs"][m] - want[m]) / denom[m]
    assert rel.max() <= tol, f"log trajectory off by {rel.max():.3e} (tol {tol:.1e})"
    worst = 0.0
    for g in GROUPS:
        for n, t in rec.get(g, {}).items():
            ok, e = P.summary_close(P.summarize(t), z[f"{g}/{n}"], tol)
            worst = max(worst, e)
            assert ok, f"{name}: {g}/{n} off by {e:.3e} (tol {tol:.1e})"
    print(f"\n[{name}] worst deviation from the reference fixture {worst:.3e} "
          f"(reference fp32-vs-fp64 {meta['reference_fp32_vs_fp64_oracle']:.3e})")
    if name == "sac_saturated":
        # float64-anchored: on the deliberately ill-conditioned case the CUDA trajectory must be as close to the
        # float64 oracle as the reference's own fp32 evaluation is, tensor group by tensor group (1.5x + 1e-4)
        from tests.golden.make_golden import run_oracle
        r32, r64 = run_oracle(inp, torch.float32, capturable=True), run_oracle(inp, torch.float64, capturable=True)
        for g in GROUPS:
            for n, t in rec.get(g, {}).items():
                d_cuda, d_ref = rel_dev(t, r64[g][n]), rel_dev(r32[g][n], r64[g][n])
                assert d_cuda <= 1.5 * d_ref + 1e-4, f"{g}/{n}: cuda-vs-fp64 {d_cuda:.3e} vs reference-fp32-vs-fp64 {d_ref:.3e}"


@pytest.mark.parametrize("name,n_iter", [("sac_hopper", 1), ("sac_hopper", 10), ("sac_hopper", 100),
                                         ("td3_hopper", 10), ("td3_hopper", 100)])
def test_params_after_n_updates(name, n_iter):
    """Parameters after N in {1,10,100} iterations (reference cadence) vs the float64 oracle, next to
    the fp32 oracle's own drift from float64 on the same trajectory."""
    inp = case_inputs(name)
    c = CASES[name]
    s = c["seed"] + 50_000
    ag = make_agent(inp)
    o32, o64 = make_oracle(inp, torch.float32), make_oracle(inp, torch.float64)
    delay = inp["hps"]["actor_update_delay"]
    for i in range(n_iter):
        idx = P.indices(s + i, c["N"], c["B"])
        b32 = {k: v[idx] for k, v in inp["storage"].items()}
        b64 = {k: (v.double() if v.is_floating_point() else v) for k, v in b32.items()}
        eq = P.noise(s + 10_000 + i, c["B"], c["ac"])
        ep = [P.noise(s + 20_000 + 10 * i + j, c["B"], c["ac"]) for j in range(delay)]
        ea = [P.noise(s + 30_000 + 10 * i + j, c["B"], c["ac"]) for j in range(delay)]
        o32.iteration(i, b32, eq, ep, ea)
        o64.iteration(i, b64, eq.double(), [e.double() for e in ep], [e.double() for e in ea])
        bd = {k: _dev(v) for k, v in b32.items()}
        ag.update_qnets(bd, eps=_dev(eq))
        ag.qnet_updates_so_far += 1
        if i % (delay + 1) == 0:
            for j in range(delay):
                ag.update_actor(bd, eps=_dev(ep[j]), eps_alpha=_dev(ea[j]))
                ag.actor_updates_so_far += 1
        ag.update_targ_nets()
    torch.cuda.synchronize()
    worst_cuda = worst_ref = 0.0
    groups = [(ag.qnet_params, o32.qnet, o64.qnet), (ag.qnet_target, o32.qnet_target, o64.qnet_target),
              (ag.actor_params, o32.actor, o64.actor)]
    if ag.td3:
        groups.append((ag.actor_target, o32.actor_target, o64.actor_target))
    for got, r32, r64 in groups:
        for n in r32:
            worst_cuda = max(worst_cuda, rel_dev(got[n], r64[n]))
            worst_ref = max(worst_ref, rel_dev(r32[n], r64[n]))
    print(f"\n[{name} N={n_iter}] max rel dev vs float64: cuda {worst_cuda:.3e}, torch-fp32 oracle {worst_ref:.3e}")
    assert np.isfinite(worst_cuda)
    assert worst_cuda <= max(1e-5, 10 * worst_ref)


# ------------------------------------------------------------------------------- Adam / Polyak kernel
def test_adam_kernel_matches_torch_capturable():
    """adam.cu vs torch.optim.Adam(capturable=True) on the GPU (the branch the reference runs)."""
    inp = case_inputs("sac_hopper")
    ag = make_agent(inp)
    g = torch.Generator(device="cuda").manual_seed(7)
    params = {n: p.detach().clone().contiguous().requires_grad_(True) for n, p in ag.qnet_params.items()}
    opt = torch.optim.Adam(list(params.values()), lr=float(ag.hps.qnets_lr), capturable=True, foreach=False)
    for step in range(1, 4):
        for n, p in ag.qnet_params.items():
            gr = torch.randn(p.shape, generator=g, device="cuda") * (10.0 ** (-step))
            p.grad.copy_(gr)
            params[n].grad = gr.clone()
        ag.arena.tensor(ag.layout.critic[0], "w2n", 4).copy_(ag.qnet1.fc_stack.fc_block_2.fc.weight.grad)
        ag.arena.tensor(ag.layout.critic[1], "w2n", 4).copy_(ag.qnet2.fc_stack.fc_block_2.fc.weight.grad)
        ag.q_optimizer.step()
        opt.step()
        for n, p in ag.qnet_params.items():
            assert rel_dev(p, params[n]) <= 1e-6, (n, step, rel_dev(p, params[n]))
    assert ag.q_optimizer.step_count == 3
    # Polyak-only pass == torch.lerp
    before = {n: t.clone() for n, t in ag.qnet_target.items()}
    ag.update_targ_nets()
    for n, t in ag.qnet_target.items():
        want = torch.lerp(before[n], ag.qnet_params[n].detach(), float(ag.hps.polyak))
        assert rel_dev(t, want) <= 1e-7
