#!/usr/bin/env python
"""HBM write-only / read-only / copy bandwidth on this GPU (torch fill_, sum, copy_ on buffers larger than L2): the ceilings
the layer-by-layer kernels that mostly WRITE activations (first layers: 268-537 MB out, ~30 MB in) run against."""
import torch

def t(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3

for mb in (268, 537, 1074):
    n = mb * 1_000_000 // 4
    a, b = torch.empty(n, device="cuda"), torch.empty(n, device="cuda")
    tw = t(lambda: a.fill_(1.0))
    tr = t(lambda: a.sum())
    tc = t(lambda: b.copy_(a))
    print(f"{mb:5d} MB: fill {tw * 1e6:7.1f} us = {mb / tw / 1e6:5.2f} TB/s | read (sum) {tr * 1e6:7.1f} us = {mb / tr / 1e6:5.2f} TB/s | "
          f"copy {tc * 1e6:7.1f} us = {2 * mb / tc / 1e6:5.2f} TB/s (read + write)")
