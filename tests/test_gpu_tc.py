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


@pytest.mark.parametrize("x3", [False, True])
@pytest.mark.parametrize("Bn,MA,lda,transpose", [(4096, 256, 256, True), (1000, 14, 28, False), (300, 6, 64, False),
                                                 (65536, 256, 256, True), (250, 393, 772, False)])
def test_tc_wgrad_matches_torch(Bn, MA, lda, transpose, x3):
    """C = A[:, :MA]^T . Bm on the tensor cores (MN-major operands, split over the batch) vs float64."""
    from sac_td3_cudagraphs_pytorch_b200 import _lib as L
    lib = L.load()
    L.init_device(torch.device("cuda"))
    g = torch.Generator(device="cuda").manual_seed(Bn + MA)
    A = torch.randn(Bn, lda, device="cuda", generator=g)
    Bm = torch.randn(Bn, 256, device="cuda", generator=g)
    Cc = torch.full((MA, 256), float("nan"), device="cuda")
    Ct = torch.full((256, MA), float("nan"), device="cuda") if transpose else None
    scratch = torch.empty(lib.b2rl_tc_wgrad_scratch_floats(MA, Bn), device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    L.check(lib.b2rl_tc_wgrad(A.data_ptr(), lda, lda, MA, Bm.data_ptr(), Bn, Cc.data_ptr(), L.ptr(Ct), scratch.data_ptr(), int(x3), None, st),
            "tc_wgrad")
    torch.cuda.synchronize()
    want = A[:, :MA].double().T @ Bm.double()
    err = float((Cc.double() - want).abs().max()) / float(want.abs().max())
    print(f"\nBn={Bn} MA={MA} 3xTF32={x3}: max rel err {err:.2e}")
    # (the 65 536-row contraction chains 384 fp32 accumulations per TMEM element: fp32-class, not 4e-6)
    assert err <= ((4e-6 if Bn <= 4096 else 2e-5) if x3 else 3e-3)
    if transpose:
        assert torch.equal(Ct, Cc.T.contiguous())
