#!/usr/bin/env python
"""First layer alone at population scale (n_agents x 256 rows, K = 14): FFMA kernel (b2rl_wide_first) vs tensor cores
(b2rl_tc_first), with and without the x-hat output."""
import ctypes as C
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from sac_td3_cudagraphs_pytorch_b200 import _lib as L  # noqa: E402


def main():
    n, M, K, ldx = (int(sys.argv[1]) if len(sys.argv) > 1 else 1024), 256, 14, 28
    lib = L.load()
    L.init_device(torch.device("cuda"))
    ps = K * 256 + 1024
    P = torch.randn(n, ps, device="cuda") * 0.1
    X = torch.randn(n * M, ldx, device="cuda")
    H, XH = torch.empty(n * M, 256, device="cuda"), torch.empty(n * M, 256, device="cuda")
    stat = torch.empty(n * M, 2, device="cuda")
    stk = L.Stack(n, 0, ps, 0, 0, 0, 0)
    st = lambda: torch.cuda.current_stream().cuda_stream
    base = P.data_ptr()
    b, g, be = base + 4 * K * 256, base + 4 * (K * 256 + 256), base + 4 * (K * 256 + 512)
    for xh in (False, True):
        xp, sp = (XH.data_ptr(), stat.data_ptr()) if xh else (None, None)
        t_f = bench.time_kernel(lambda: L.check(lib.b2rl_wide_first(X.data_ptr(), ldx, M, K, base, b, g, be, 1, H.data_ptr(), xp, sp,
                                                                   C.byref(stk), st())), iters=60, warm=10, per_graph=10)
        out = [f"ffma {t_f * 1e6:7.1f} us"]
        for x3 in (0, 1):
            t = bench.time_kernel(lambda: L.check(lib.b2rl_tc_first(X.data_ptr(), ldx, M, K, base, b, g, be, 1, H.data_ptr(), xp, sp, x3,
                                                                    C.byref(stk), st())), iters=60, warm=10, per_graph=10)
            out.append(f"tc {'3xtf32' if x3 else 'tf32  '} {t * 1e6:7.1f} us")
        mb = n * M * 256 * 4 * (2 if xh else 1) / 1e6
        print(f"n={n} rows={n * M} x-hat={xh!s:5s} ({mb:.0f} MB out): " + " | ".join(out))


if __name__ == "__main__":
    main()
