"""tcgen05 (TF32) large-batch hidden layer vs torch: Linear -> LayerNorm -> ReLU (agents/nets.py:66-82).
Tolerance: TF32 products (10-bit mantissa, truncated operands) — stated bound 3e-3 of the output scale; the
LayerNorm statistics and everything after the product are fp32."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("x3", [False, True])
@pytest.mark.parametrize("M,ln,relu", [(128, True, True), (1000, True, True), (4096, False, True), (300, True, False)])
def test_tc_linear_matches_torch(M, ln, relu, x3):
    from sac_td3_cudagraphs_pytorch_b200 import _lib as L
    lib = L.load()
    L.init_device(torch.device("cuda"))
    g = torch.Generator(device="cuda").manual_seed(M)
    X = torch.randn(M, 256, device="cuda", generator=g)
    W = torch.randn(256, 256, device="cuda", generator=g) / 16.0
    b = torch.randn(256, device="cuda", generator=g) * 0.1
    gam = 1.0 + 0.1 * torch.randn(256, device="cuda", generator=g)
    bet = 0.1 * torch.randn(256, device="cuda", generator=g)
    H = torch.full((M, 256), float("nan"), device="cuda")
    XH = torch.full((M, 256), float("nan"), device="cuda")
    stat = torch.zeros(M, 2, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    Wlo = torch.empty_like(W)
    if x3:
        L.check(lib.b2rl_tc_split_lo(W.data_ptr(), Wlo.data_ptr(), W.numel(), st), "split")
    L.check(lib.b2rl_tc_linear(X.data_ptr(), 256, M, W.data_ptr(), Wlo.data_ptr() if x3 else None, b.data_ptr(), gam.data_ptr(),
                               bet.data_ptr(), int(ln),
                               int(relu), H.data_ptr(), XH.data_ptr(), stat.data_ptr(), st), "tc_linear")
    torch.cuda.synchronize()
    z = X.double() @ W.double().T + b.double()
    if ln:
        mu, var = z.mean(1, keepdim=True), z.var(1, unbiased=False, keepdim=True)
        xh = (z - mu) / torch.sqrt(var + 1e-5)
        y = xh * gam.double() + bet.double()
    else:
        xh, y = z, z
    want = torch.relu(y) if relu else y
    assert torch.isfinite(H).all() and torch.isfinite(XH).all()
    tol = 4e-6 if x3 else 3e-3  # 3xTF32: fp32-level; TF32: 10-bit mantissas
    scale = float(want.abs().max())
    eh, ex = float((H.double() - want).abs().max()) / scale, float((XH.double() - xh).abs().max()) / float(xh.abs().max())
    print(f"\nM={M} ln={ln} 3xTF32={x3}: max rel err H {eh:.2e}, x-hat {ex:.2e}")
    assert eh <= tol and ex <= tol
    if ln:
        assert float((stat[:, 0].double() - mu[:, 0]).abs().max()) <= tol * 4
