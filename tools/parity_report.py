#!/usr/bin/env python
"""Print the parity table (CUDA vs fp32 oracle vs float64 oracle) for every golden case.
Test tooling (uses oracle/); run on a GPU box:  python tools/parity_report.py [case ...]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from tests.golden.cases import CASES, case_inputs  # noqa: E402
from tests.helpers import batch_of, make_agent, make_oracle, rel_dev  # noqa: E402


def row(label, got, r32, r64):
    print(f"    {label:44s} cuda-o32 {rel_dev(got, r32):9.2e}  cuda-o64 {rel_dev(got, r64):9.2e}  o32-o64 {rel_dev(r32, r64):9.2e}")


def main(names):
    for name in names:
        inp = case_inputs(name)
        dev = lambda d: {k: v.cuda() for k, v in d.items()}
        print(f"== {name}  (B={inp['B']}, ob={inp['ob']}, ac={inp['ac']})")
        ag = make_agent(inp)
        o32, o64 = make_oracle(inp, torch.float32), make_oracle(inp, torch.float64)
        B = inp["B"]
        tq, qv = torch.zeros(B, device="cuda"), torch.zeros(2, B, device="cuda")
        out = ag.update_qnets(dev(batch_of(inp, 0)), eps=inp["eps_q"][0].cuda(), dbg_targ_q=tq, dbg_q=qv)
        r32 = o32.update_qnets(batch_of(inp, 0), inp["eps_q"][0])
        r64 = o64.update_qnets(batch_of(inp, 0, torch.float64), inp["eps_q"][0].double())
        print("  critic step")
        row("targ_q", tq, r32["_targ_q"], r64["_targ_q"])
        row("q", qv, r32["_q"], r64["_q"])
        row("qf_loss", out["loss/qf_loss"], r32["loss/qf_loss"], r64["loss/qf_loss"])
        for n, p in ag.qnet_params.items():
            row("grad " + n, p.grad, o32.qnet[n].grad, o64.qnet[n].grad)
        for n, p in ag.qnet_params.items():
            row("param " + n, p, o32.qnet[n], o64.qnet[n])
        # actor step from a fresh identical state
        ag = make_agent(inp)
        o32, o64 = make_oracle(inp, torch.float32), make_oracle(inp, torch.float64)
        e1, e2 = inp["eps_pi"][0][0], inp["eps_alpha"][0][0]
        out = ag.update_actor(dev(batch_of(inp, 0)), eps=e1.cuda(), eps_alpha=e2.cuda())
        r32 = o32.update_actor(batch_of(inp, 0), e1, e2)
        r64 = o64.update_actor(batch_of(inp, 0, torch.float64), e1.double(), e2.double())
        print("  actor step")
        for k in r32:
            row(k, out[k], r32[k], r64[k])
            print(f"      values: cuda {float(out[k]):.8f} o32 {float(r32[k]):.8f} o64 {float(r64[k]):.8f}")
        for n, p in ag.actor_params.items():
            row("grad " + n, p.grad, o32.actor[n].grad, o64.actor[n].grad)
        for n, p in ag.actor_params.items():
            row("param " + n, p, o32.actor[n], o64.actor[n])
        if not ag.td3:
            row("log_alpha", ag.log_alpha, o32.log_alpha, o64.log_alpha)


if __name__ == "__main__":
    main(sys.argv[1:] or list(CASES))
