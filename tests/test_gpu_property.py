"""Property tests (SURVEY §4 "Property tests"): hypothesis draws the shape and option space of the update — ob/ac dims incl.
odd ones (O = 1..48: both the shared-memory and the global-memory first-layer paths; A = 1..17), ragged batch sizes, LayerNorm
on/off, SAC/TD3, BCQ target mix, target smoothing, autotune, gradient clipping — and for every draw one critic step and one
actor (+ temperature) step on the CUDA path must match the fp32 oracle within max(1e-5, 4 x the oracle's own fp32-vs-fp64
gap) per tensor (tests/helpers.check_close), the bar of the fixed-case parity tests."""
import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from tests.golden.cases import SAC, TD3, case_inputs
from tests.helpers import (alpha_loss_scale, batch_of, check_close, check_param_after_first_adam, make_agent, make_oracle)

pytestmark = pytest.mark.gpu


@st.composite
def update_cases(draw):
    td3 = draw(st.booleans())
    ob, ac = draw(st.integers(1, 48)), draw(st.integers(1, 17))
    B = draw(st.sampled_from([1, 5, 8, 13, 32, 40, 64, 100, 256]))
    over = dict(layer_norm=draw(st.booleans()), bcq_style_targ_mix=draw(st.booleans()),
                clip_norm=draw(st.sampled_from([0.0, 0.0, 0.25])), batch_size=B)
    if td3:
        over.update(targ_actor_smoothing=draw(st.booleans()))
    else:
        over.update(autotune=draw(st.booleans()), alpha_init=draw(st.sampled_from([0.2, 1.0])))
    lo = [-(1.0 + 0.25 * (i % 3)) for i in range(ac)]
    hi = [0.5 + 0.5 * (i % 2) for i in range(ac)]  # asymmetric per-dimension bounds
    return dict(base=TD3 if td3 else SAC, ob=ob, ac=ac, lo=lo, hi=hi, B=B, N=max(2 * B, 16), iters=1,
                seed=draw(st.integers(1, 10_000)) * 7, over=over)


@settings(max_examples=40, deadline=None, derandomize=True, suppress_health_check=list(HealthCheck))
@given(case=update_cases())
def test_single_steps_match_the_oracle_for_any_shape_and_option(case):
    inp = case_inputs(case)
    dev = lambda d: {k: v.cuda() for k, v in d.items()}
    # ---- critic step
    ag = make_agent(inp)
    o32, o64 = make_oracle(inp, torch.float32), make_oracle(inp, torch.float64)
    out = ag.update_qnets(dev(batch_of(inp, 0)), eps=inp["eps_q"][0].cuda())
    r32 = o32.update_qnets(batch_of(inp, 0), inp["eps_q"][0])
    r64 = o64.update_qnets(batch_of(inp, 0, torch.float64), inp["eps_q"][0].double())
    torch.cuda.synchronize()
    check_close("qf_loss", out["loss/qf_loss"], r32["loss/qf_loss"], r64["loss/qf_loss"])
    for n, p in ag.qnet_params.items():
        check_close(f"grad {n}", p.grad, o32.qnet[n].grad, o64.qnet[n].grad)
        check_param_after_first_adam(f"param {n}", p, o32.qnet[n], o64.qnet[n], p.grad, o32.qnet[n].grad, float(inp["hps"]["qnets_lr"]))
    # ---- actor (+ temperature) step from the same initial state
    ag = make_agent(inp)
    o32, o64 = make_oracle(inp, torch.float32), make_oracle(inp, torch.float64)
    e1, e2 = inp["eps_pi"][0][0], inp["eps_alpha"][0][0]
    out = ag.update_actor(dev(batch_of(inp, 0)), eps=e1.cuda(), eps_alpha=e2.cuda())
    r32 = o32.update_actor(batch_of(inp, 0), e1, e2)
    r64 = o64.update_actor(batch_of(inp, 0, torch.float64), e1.double(), e2.double())
    torch.cuda.synchronize()
    for k in r32:
        sc = alpha_loss_scale(inp["hps"]["alpha_init"], inp["ac"]) if k == "loss/alpha_loss" else None
        check_close(k, out[k], r32[k], r64[k], scale=sc)
    if not inp["hps"]["clip_norm"] > 0:  # (a clipped gradient is compared after the clip: see the fixed clip cases)
        for n, p in ag.actor_params.items():
            check_close(f"grad {n}", p.grad, o32.actor[n].grad, o64.actor[n].grad)
            check_param_after_first_adam(f"param {n}", p, o32.actor[n], o64.actor[n], p.grad, o32.actor[n].grad,
                                         float(inp["hps"]["actor_lr"]))
    if not ag.td3 and ag.autotune:
        check_close("log_alpha", ag.log_alpha, o32.log_alpha, o64.log_alpha)
