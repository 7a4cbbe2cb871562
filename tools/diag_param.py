#!/usr/bin/env python
"""Diagnostic (test tooling): where does the worst parameter deviation after one actor step come from?"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from tests.golden.cases import case_inputs
from tests.helpers import batch_of, make_agent, make_oracle

name = sys.argv[1] if len(sys.argv) > 1 else "sac_humanoid"
inp = case_inputs(name)
ag = make_agent(inp)
o32, o64 = make_oracle(inp, torch.float32), make_oracle(inp, torch.float64)
e1, e2 = inp["eps_pi"][0][0], inp["eps_alpha"][0][0]
dev = lambda d: {k: v.cuda() for k, v in d.items()}
p0 = {n: p.detach().clone().cpu() for n, p in ag.actor_params.items()}
ag.update_actor(dev(batch_of(inp, 0)), eps=e1.cuda(), eps_alpha=e2.cuda())
o32.update_actor(batch_of(inp, 0), e1, e2)
o64.update_actor(batch_of(inp, 0, torch.float64), e1.double(), e2.double())
torch.cuda.synchronize()
for n, p in ag.actor_params.items():
    g, g32, g64 = p.grad.cpu().double(), o32.actor[n].grad.double(), o64.actor[n].grad
    q, q32, q64 = p.detach().cpu().double(), o32.actor[n].detach().double(), o64.actor[n].detach()
    d = (q - q32).abs()
    i = int(d.argmax())
    fl = lambda t: float(t.reshape(-1)[i])
    print(f"{n}: max|p| {float(q32.abs().max()):.3e} max|g| {float(g32.abs().max()):.3e}")
    print(f"   worst dp {float(d.max()):.3e} at {i}: g cuda {fl(g):.6e} o32 {fl(g32):.6e} o64 {fl(g64):.6e}"
          f" | step cuda {fl(q) - float(p0[n].reshape(-1)[i]):.6e} o32 {fl(q32) - float(p0[n].reshape(-1)[i]):.6e} o64 {fl(q64) - float(p0[n].double().reshape(-1)[i]):.6e}")
    print(f"   grad: max|cuda-o64| {float((g - g64).abs().max()):.3e}  max|o32-o64| {float((g32 - g64).abs().max()):.3e}"
          f"  rms cuda {float((g - g64).pow(2).mean().sqrt()):.3e} rms o32 {float((g32 - g64).pow(2).mean().sqrt()):.3e}")
