"""Throughput mode: bf16 operands on the tcgen05 tensor cores (fp32 accumulate, fp32 master weights / activations).
bf16 keeps 8 mantissa bits, so the tolerance here is 2e-2 relative L2 per op (stated; the exact-fp32 mode is held to 1e-3
in test_ops.py / test_model.py)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = 2e-2


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


@pytest.fixture(autouse=True)
def bf16_mode():
    from transfuser_b200 import gemm
    old = gemm.MODE
    gemm.set_mode('bf16')
    yield
    gemm.set_mode(old)


@pytest.mark.parametrize('M,K,N', [(348, 72, 72), (348, 216, 864), (1740, 1512, 1512), (14080, 72, 72), (70400, 32 + 40, 216), (200, 64, 64)])
def test_linear_bf16_fwd_bwd(M, K, N):
    from transfuser_b200 import ops
    g = torch.Generator(device='cuda').manual_seed(M + K)
    x = torch.randn(M, K, device='cuda', generator=g).requires_grad_()
    w = (torch.randn(N, K, device='cuda', generator=g) / math.sqrt(K)).requires_grad_()
    b = torch.randn(N, device='cuda', generator=g).requires_grad_()
    ref = F.relu(F.linear(x, w, b))
    xm, wm, bm = [t.detach().clone().requires_grad_() for t in (x, w, b)]
    out = ops.linear(xm, wm, bm, relu=True)
    assert rel(out, ref) < TOL
    go = torch.randn(M, N, device='cuda', generator=g)
    for a, r in zip(torch.autograd.grad(out, [xm, wm, bm], go), torch.autograd.grad(ref, [x, w, b], go)):
        assert rel(a, r) < TOL


def test_attention_bf16():
    from transfuser_b200 import ops
    B, T, C, nh = 2, 174, 216, 4
    g = torch.Generator(device='cuda').manual_seed(3)
    h = torch.randn(B * T, C, device='cuda', generator=g).requires_grad_()
    ws = [(torch.randn(C, C, device='cuda', generator=g) / math.sqrt(C)).requires_grad_() for _ in range(3)]
    bs = [(torch.randn(C, device='cuda', generator=g) * 0.1).requires_grad_() for _ in range(3)]
    q, k, v = [F.linear(h, ws[i], bs[i]).view(B, T, nh, C // nh).transpose(1, 2) for i in range(3)]
    ref = (F.softmax((q @ k.transpose(-2, -1)) / math.sqrt(C // nh), dim=-1) @ v).transpose(1, 2).reshape(B * T, C)
    hm = h.detach().clone().requires_grad_()
    wm = [w.detach().clone().requires_grad_() for w in ws]
    bm = [b.detach().clone().requires_grad_() for b in bs]
    out = ops.AttentionFn.apply(hm, wm[0], bm[0], wm[1], bm[1], wm[2], bm[2], B, T, nh, 0.0, 1)
    assert rel(out, ref) < TOL
    go = torch.randn(B * T, C, device='cuda', generator=g)
    for a, r in zip(torch.autograd.grad(out, [hm] + wm, go), torch.autograd.grad(ref, [h] + ws, go)):
        assert rel(a, r) < 2 * TOL
