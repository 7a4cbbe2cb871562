#!/usr/bin/env python
"""Throughput of the tcgen05 hidden layer (b2rl_tc_linear) at large batch, next to torch's fp32 / TF32 GEMM+LN+ReLU."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from sac_td3_cudagraphs_pytorch_b200 import _lib as L

lib = L.load(); L.init_device(torch.device("cuda"))
for M in [int(x) for x in sys.argv[1:]] or [4096, 65536, 262144]:
    X = torch.randn(M, 256, device="cuda"); W = torch.randn(256, 256, device="cuda") / 16
    b = torch.randn(256, device="cuda"); g = torch.ones(256, device="cuda"); be = torch.zeros(256, device="cuda")
    H = torch.empty(M, 256, device="cuda"); XH = torch.empty(M, 256, device="cuda"); stat = torch.empty(M, 2, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    Wlo = torch.empty_like(W)
    L.check(lib.b2rl_tc_split_lo(W.data_ptr(), Wlo.data_ptr(), W.numel(), None, st), "split")
    x3 = False
    def ours(xh=True):
        L.check(lib.b2rl_tc_linear(X.data_ptr(), 256, M, W.data_ptr(), Wlo.data_ptr() if x3 else None, b.data_ptr(), g.data_ptr(), be.data_ptr(), 1, 1,
                                   H.data_ptr(), XH.data_ptr() if xh else None, stat.data_ptr(), None, st), "tc")
    def ref():
        return torch.relu(torch.nn.functional.layer_norm(torch.addmm(b, X, W.t()), (256,), g, be))
    def timeit(fn, n=30):
        for _ in range(5): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e-3
    fl = 2.0 * M * 256 * 256
    t1, t1b = timeit(ours), timeit(lambda: ours(False))
    x3 = True; t1x = timeit(ours); x3 = False
    torch.backends.cuda.matmul.allow_tf32 = False; t2 = timeit(ref)
    torch.backends.cuda.matmul.allow_tf32 = True; t3 = timeit(ref)
    by = M * 256 * 4
    print(f"M={M:7d}: b2rl_tc_linear {t1*1e6:8.1f} us = {fl/t1/1e12:6.1f} TFLOP/s, {3*by/t1/1e9:6.0f} GB/s (X in, H + XH out) | "
          f"3xTF32 {t1x*1e6:8.1f} us | H only {t1b*1e6:8.1f} us {2*by/t1b/1e9:6.0f} GB/s | torch fp32 addmm+LN+ReLU {t2*1e6:8.1f} us | torch TF32 {t3*1e6:8.1f} us", flush=True)
