"""Per-op parity of the CUDA kernels (forward and backward, through the C-ABI) against plain fp32 PyTorch ops.
Tolerances: 1e-4 relative L2 for the exact-fp32 kernels (north_star: 1e-3), stated per test otherwise."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
TOL = 1e-4
DEV = 'cuda'   # tests/test_ops_emulated.py re-runs these functions with DEV = 'cpu' over the CPU emulation of the kernels
DROPOUT_N, LN_ROWS, ATT_T = 1 << 20, 348, 174   # (and smaller populations there: emulated thread barriers are slow)


_NOLOG = []
SKIPPED_ON_EMULATOR = []      # tests that skip themselves under the CPU emulation say so here (tests/test_ops_emulated.py checks every other test reached the emulated C-ABI)


def _calls():
    """The emulated library's call log when the test runs over the CPU emulation (tests/test_ops_emulated.py), else _NOLOG."""
    from transfuser_b200 import _lib
    return getattr(_lib._LIB, 'log', _NOLOG)


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return torch.randn(*shape, device=DEV, generator=g) * scale


def check_grads(mine_inputs, ref_inputs, mine_out, ref_out, tol=TOL, gseed=7):
    g = rnd(*ref_out.shape, seed=gseed)
    gm = nhwc(g) if (mine_out.dim() == 4 and mine_out.shape != ref_out.shape) else g.view_as(mine_out)
    ref_g = torch.autograd.grad(ref_out, ref_inputs, g, allow_unused=True)
    my_g = torch.autograd.grad(mine_out, mine_inputs, gm, allow_unused=True)
    for i, (a, b) in enumerate(zip(my_g, ref_g)):
        if a.dim() == 4 and a.shape != b.shape:
            a = nchw(a)
        assert rel(a.reshape(b.shape), b) < tol, ('grad', i, rel(a.reshape(b.shape), b))


@pytest.mark.parametrize('cfg', [
    # N, H, W, Cin, Cout, k, stride, groups, bias, relu
    (2, 16, 24, 3, 32, 3, 2, 1, False, False),      # stem (wgrad: the warp-per-pixel kernel for Cin <= 4)
    (1, 9, 11, 2, 40, 3, 1, 1, True, False),        # 2 input channels, 40 output channels (two 32-wide co blocks), bias
    (2, 10, 7, 4, 8, 3, 2, 1, True, True),          # 4 input channels, stride 2, bias + relu
    (2, 12, 20, 72, 72, 3, 1, 3, False, False),     # grouped 3x3, group width 24
    (2, 12, 20, 72, 72, 3, 2, 3, False, False),     # grouped, stride 2
    (1, 10, 14, 216, 216, 3, 1, 9, False, False),
    (2, 9, 11, 32, 72, 1, 2, 1, False, False),      # strided 1x1 shortcut
    (2, 8, 8, 64, 64, 3, 1, 1, True, True),         # head 3x3 + bias + relu
    (1, 5, 22, 512, 128, 3, 1, 1, True, True),      # decoder first layer
    (2, 12, 16, 32, 7, 3, 1, 1, True, False),       # decoder last layer (Cout 7)
    (2, 12, 16, 32, 1, 3, 1, 1, True, False),       # depth last layer (Cout 1)
    (2, 7, 9, 72, 216, 1, 1, 1, False, False),      # 1x1 -> GEMM path
    (2, 8, 8, 64, 12, 1, 1, 1, True, False),        # head 1x1 with bias
])
def test_conv2d(cfg):
    from transfuser_b200 import ops
    N, H, W, Cin, Cout, k, s, g, has_b, relu = cfg
    x = rnd(N, Cin, H, W, seed=1).requires_grad_()
    w = rnd(Cout, Cin // g, k, k, seed=2, scale=1.0 / math.sqrt(Cin // g * k * k)).requires_grad_()
    b = rnd(Cout, seed=3).requires_grad_() if has_b else None
    ref = F.conv2d(x, w, b, stride=s, padding=k // 2, groups=g)
    ref = F.relu(ref) if relu else ref
    xm = nhwc(x.detach()).requires_grad_()
    wm = w.detach().clone().requires_grad_()
    bm = b.detach().clone().requires_grad_() if has_b else None
    out = ops.conv2d(xm, wm, bm, s, g, relu)
    assert rel(nchw(out), ref) < TOL
    check_grads([xm, wm] + ([bm] if has_b else []), [x, w] + ([b] if has_b else []), out, ref)


@pytest.mark.parametrize('shape,relu', [((3, 10, 12, 72), True), ((2, 6, 7, 216), False), ((1, 5, 22, 1512), True),
                                        ((2, 48, 44, 72), True), ((2, 40, 56, 216), False), ((1, 80, 64, 32), True),      # M >= 4096: the flat column reduction
                                        ((2, 10, 22, 576), True), ((3, 9, 7, 1512), False)])                            # the stage-3 / stage-4 widths
def test_batchnorm_train(shape, relu):
    from transfuser_b200 import ops
    N, H, W, C = shape
    x = (rnd(N, C, H, W, seed=4) * 2 + 0.5).requires_grad_()
    bn = torch.nn.BatchNorm2d(C).to(DEV)
    bn.weight.data = rnd(C, seed=5) * 0.3 + 1
    bn.bias.data = rnd(C, seed=6) * 0.2
    bn2 = torch.nn.BatchNorm2d(C).to(DEV)
    bn2.load_state_dict(bn.state_dict())
    ref = bn(x)
    ref = F.relu(ref) if relu else ref
    xm = nhwc(x.detach()).requires_grad_()
    out = ops.batch_norm(xm, bn2, relu, True)
    assert rel(nchw(out), ref) < TOL
    assert rel(bn2.running_mean, bn.running_mean) < TOL and rel(bn2.running_var, bn.running_var) < TOL
    check_grads([xm, bn2.weight, bn2.bias], [x, bn.weight, bn.bias], out, ref)


@pytest.mark.parametrize('shape', [(3, 10, 12, 72), (2, 6, 7, 216), (1, 5, 22, 1512), (2, 48, 44, 72), (2, 10, 22, 576)])
def test_batchnorm_add_relu_fused(shape):
    """relu(bn(x) + shortcut) — the Bottleneck tail — as ONE BatchNorm call (add + ReLU inside the normalise pass; ReLU mask and the
    shortcut's gradient inside the backward passes) against torch and against the unfused product path (eval mode too)."""
    from transfuser_b200 import ops
    N, H, W, C = shape
    x = (rnd(N, C, H, W, seed=4) * 2 + 0.5).requires_grad_()
    sc = rnd(N, C, H, W, seed=8).requires_grad_()
    bn = torch.nn.BatchNorm2d(C).to(DEV)
    bn.weight.data = rnd(C, seed=5) * 0.3 + 1
    bn.bias.data = rnd(C, seed=6) * 0.2
    state = {k: v.clone() for k, v in bn.state_dict().items()}
    ref = F.relu(bn(x) + sc)
    outs = {}
    old = ops.BN_ADD_FUSED
    try:
        for fused in (True, False):
            ops.BN_ADD_FUSED = fused
            bn2 = torch.nn.BatchNorm2d(C).to(DEV)
            bn2.load_state_dict(state)
            xm, sm = nhwc(x.detach()).requires_grad_(), nhwc(sc.detach()).requires_grad_()
            n0 = len(_calls())
            out = ops.batch_norm(xm, bn2, True, True, residual=sm)
            if _calls() is not _NOLOG:
                assert len(_calls()) - n0 == (1 if fused else 2)
            bn.load_state_dict(state)
            ref = F.relu(bn(x) + sc)
            assert rel(nchw(out), ref) < TOL
            assert rel(bn2.running_mean, bn.running_mean) < TOL and rel(bn2.running_var, bn.running_var) < TOL
            check_grads([xm, sm, bn2.weight, bn2.bias], [x, sc, bn.weight, bn.bias], out, ref)
            outs[fused] = out.detach()
            bn2.eval()
            with torch.no_grad():
                ev = ops.batch_norm(xm.detach(), bn2, True, False, residual=sm.detach())
                bn.eval()
                assert rel(nchw(ev), F.relu(bn(x) + sc)) < TOL
                bn.train()
    finally:
        ops.BN_ADD_FUSED = old
    assert torch.equal(outs[True], outs[False])          # same arithmetic, one pass instead of two


def test_layernorm_linear_dropout():
    from transfuser_b200 import ops
    x = rnd(LN_ROWS, 216, seed=1).requires_grad_()
    ln = torch.nn.LayerNorm(216).to(DEV)
    ln.weight.data = rnd(216, seed=2) * 0.2 + 1
    ln.bias.data = rnd(216, seed=3) * 0.1
    lin = torch.nn.Linear(216, 864).to(DEV)
    ref = F.relu(lin(ln(x)))
    xm = x.detach().clone().requires_grad_()
    ln2, lin2 = torch.nn.LayerNorm(216).to(DEV), torch.nn.Linear(216, 864).to(DEV)
    ln2.load_state_dict(ln.state_dict()); lin2.load_state_dict(lin.state_dict())
    out = ops.linear(ops.layer_norm(xm, ln2), lin2.weight, lin2.bias, relu=True)
    assert rel(out, ref) < TOL
    check_grads([xm, ln2.weight, ln2.bias, lin2.weight, lin2.bias], [x, ln.weight, ln.bias, lin.weight, lin.bias], out, ref)
    # dropout: keep-probability and scaling, same mask regenerated in backward
    y = rnd(DROPOUT_N, seed=9).abs().requires_grad_()
    d = ops.DropoutFn.apply(y, 0.1, 1234)
    kept = (d != 0).float().mean().item()
    assert abs(kept - 0.9) < 5e-3
    assert torch.allclose(d[d != 0], (y / 0.9)[d != 0], rtol=1e-6)
    (gy,) = torch.autograd.grad(d.sum(), y)
    assert torch.equal(gy != 0, d != 0)
    # residual + dropout in one launch: the same mask as DropoutFn with the same seed (incl. the scalar tail of an odd length)
    for n in (DROPOUT_N, 1027):
        yb, res = y.detach()[:n].clone().requires_grad_(), rnd(n, seed=10).requires_grad_()
        fused = ops.AddDropoutFn.apply(res, yb, 0.1, 1234)
        assert torch.equal(fused.detach(), res.detach() + ops.DropoutFn.apply(yb.detach(), 0.1, 1234))
        gr, gb = torch.autograd.grad(fused, (res, yb), torch.ones_like(fused))
        assert torch.equal(gr, torch.ones_like(gr)) and torch.equal(gb != 0, d.detach()[:n] != 0)
    assert ops.add_dropout(res, yb, 0.1, False).equal(res + yb) and ops.add_dropout(res, yb, 0.0, True).equal(res + yb)


@pytest.mark.parametrize('p', [0.0, 0.1])
def test_residual_dropout_layernorm_fused(p):
    """ops.add_dropout_ln == add_dropout followed by layer_norm: same mask (same seed), both outputs and all five gradients, with the
    residual stream consumed twice (by the LayerNorm and by the next residual connection) as in the GPT block."""
    from transfuser_b200 import ops
    C = 216 if DEV != 'cpu' else 72          # (CPU emulation: one OS thread per CUDA thread — keep the rows short)
    ln = torch.nn.LayerNorm(C).to(DEV)
    ln.weight.data = rnd(C, seed=2) * 0.2 + 1
    ln.bias.data = rnd(C, seed=3) * 0.1
    w = rnd(LN_ROWS, C, seed=6)
    outs = {}
    old = ops.ADD_LN_FUSED
    try:
        for fused in (False, True):
            ops.ADD_LN_FUSED = fused
            res, x = rnd(LN_ROWS, C, seed=4).requires_grad_(), rnd(LN_ROWS, C, seed=5).requires_grad_()
            ln.zero_grad()
            ops._SEED['off'] = 0x70000000
            xnew, h = ops.add_dropout_ln(res, x, p, True, ln, emit16=False)
            loss = (h * w).sum() + (xnew * xnew).sum() * 0.5
            g = torch.autograd.grad(loss, (res, x, ln.weight, ln.bias))
            outs[fused] = (xnew.detach(), h.detach()) + tuple(g)
            # the LayerNorm output alone / the residual stream alone
            xnew, h = ops.add_dropout_ln(res, x, 0.0, True, ln)
            (g1,) = torch.autograd.grad(xnew.sum(), res)
            assert torch.equal(g1, torch.ones_like(g1))
    finally:
        ops.ADD_LN_FUSED = old
    assert torch.equal(outs[True][0], outs[False][0])
    for a, b in zip(outs[True][1:], outs[False][1:]):
        assert rel(a, b) < 1e-5
    if p == 0.0:
        res, x = rnd(LN_ROWS, C, seed=4), rnd(LN_ROWS, C, seed=5)
        assert rel(outs[True][1], F.layer_norm(res + x, (C,), ln.weight, ln.bias, ln.eps)) < TOL


@pytest.mark.parametrize('packed', [False, True])
@pytest.mark.parametrize('C,nh', [(72, 4), (216, 4)])
def test_attention(C, nh, packed):
    """packed: query / key / value parameters back to back in one buffer, as optim.FlatParams lays them out -> the projections,
    their dgrad, wgrad and bias gradients each run as ONE GEMM / reduction on the [3C, C] pack (ops._pack3)."""
    from transfuser_b200 import ops
    B, T = 2, ATT_T
    h = rnd(B * T, C, seed=1).requires_grad_()
    ws = [rnd(C, C, seed=10 + i, scale=1 / math.sqrt(C)).requires_grad_() for i in range(3)]
    bs = [rnd(C, seed=20 + i, scale=0.1).requires_grad_() for i in range(3)]
    q = F.linear(h, ws[0], bs[0]).view(B, T, nh, C // nh).transpose(1, 2)
    k = F.linear(h, ws[1], bs[1]).view(B, T, nh, C // nh).transpose(1, 2)
    v = F.linear(h, ws[2], bs[2]).view(B, T, nh, C // nh).transpose(1, 2)
    att = F.softmax((q @ k.transpose(-2, -1)) * (1.0 / math.sqrt(C // nh)), dim=-1)
    ref = (att @ v).transpose(1, 2).reshape(B * T, C)
    hm = h.detach().clone().requires_grad_()
    if packed:
        flat = torch.cat([w.detach().reshape(-1) for w in ws] + [b.detach() for b in bs]).contiguous()
        wm = [flat[i * C * C:(i + 1) * C * C].view(C, C).requires_grad_() for i in range(3)]
        bm = [flat[3 * C * C + i * C:3 * C * C + (i + 1) * C].requires_grad_() for i in range(3)]
        assert ops._pack3(*wm).shape == (3 * C, C) and ops._pack3(*bm).shape == (3 * C,)
    else:
        wm = [w.detach().clone().requires_grad_() for w in ws]
        bm = [b.detach().clone().requires_grad_() for b in bs]
        assert ops._pack3(*wm) is None
    n0 = len(_calls())
    out = ops.AttentionFn.apply(hm, wm[0], bm[0], wm[1], bm[1], wm[2], bm[2], B, T, nh, 0.0, 1)
    if _calls() is not _NOLOG:
        assert len(_calls()) - n0 == (4 if packed else 6)      # q|k|v GEMM(s) + scores + softmax + AV
    assert rel(out, ref) < TOL
    g = rnd(B * T, C, seed=5)
    rg = torch.autograd.grad(ref, [h] + ws + [bs[0], bs[2]], g)
    mg = torch.autograd.grad(out, [hm] + wm + [bm[0], bm[2]], g)
    for a, b_ in zip(mg, rg):
        assert rel(a, b_) < TOL


@pytest.mark.parametrize('dims', [(3, 6, 9, 72, 8), (2, 32, 36, 72, 8), (2, 32, 34, 216, 18)])     # HW >= 1024: the flat SE reduction
def test_se_add_pool(dims):
    from transfuser_b200 import ops
    N, H, W, C, Cr = dims
    x = rnd(N, C, H, W, seed=1).requires_grad_()
    w1, b1 = rnd(Cr, C, 1, 1, seed=2, scale=0.2).requires_grad_(), rnd(Cr, seed=3, scale=0.1).requires_grad_()
    w2, b2 = rnd(C, Cr, 1, 1, seed=4, scale=0.3).requires_grad_(), rnd(C, seed=5, scale=0.1).requires_grad_()
    s = x.mean((2, 3), keepdim=True)
    ref = x * torch.sigmoid(F.conv2d(F.relu(F.conv2d(s, w1, b1)), w2, b2))
    xm = nhwc(x.detach()).requires_grad_()
    ps = [t.detach().clone().requires_grad_() for t in (w1, b1, w2, b2)]
    out = ops.SEFn.apply(xm, *ps)
    assert rel(nchw(out), ref) < TOL
    check_grads([xm] + ps, [x, w1, b1, w2, b2], out, ref)
    # residual add + relu, global pool
    a, b = rnd(N, H, W, C, seed=6).requires_grad_(), rnd(N, H, W, C, seed=7).requires_grad_()
    am, bm = a.detach().clone().requires_grad_(), b.detach().clone().requires_grad_()
    ref2, out2 = F.relu(a + b), ops.add(am, bm, relu=True)
    assert rel(out2, ref2) < 1e-6
    check_grads([am, bm], [a, b], out2, ref2)
    ref3, out3 = a.mean((1, 2)), ops.PoolHWFn.apply(am)
    assert rel(out3, ref3) < 1e-5
    check_grads([am], [a], out3, ref3)


@pytest.mark.parametrize('shape', [(3, 20, 24, 72), (2, 5, 22, 1512), (2, 40, 44, 216)])
def test_batchnorm_with_se_pool(shape):
    """conv2.bn -> SE: the squeeze-excite average pool comes out of BatchNorm's normalise pass (accumulator cleared by the statistics
    kernel, no memset / pooling launch); BatchNorm's own output is unchanged and SE(bn(x)) matches torch forward and backward."""
    from transfuser_b200 import ops
    N, H, W, C = shape
    Cr = 8
    x = (rnd(N, C, H, W, seed=1) * 2 + 0.3).requires_grad_()
    bn = torch.nn.BatchNorm2d(C).to(DEV)
    bn.weight.data = rnd(C, seed=2) * 0.3 + 1
    bn.bias.data = rnd(C, seed=3) * 0.2
    bn2 = torch.nn.BatchNorm2d(C).to(DEV)
    bn2.load_state_dict(bn.state_dict())
    bn3 = torch.nn.BatchNorm2d(C).to(DEV)
    bn3.load_state_dict(bn.state_dict())
    w1, b1 = rnd(Cr, C, 1, 1, seed=4, scale=0.1).requires_grad_(), rnd(Cr, seed=5, scale=0.1).requires_grad_()
    w2, b2 = rnd(C, Cr, 1, 1, seed=6, scale=0.3).requires_grad_(), rnd(C, seed=7, scale=0.1).requires_grad_()
    yb = F.relu(bn(x))
    ref = yb * torch.sigmoid(F.conv2d(F.relu(F.conv2d(yb.mean((2, 3), keepdim=True), w1, b1)), w2, b2))
    xm = nhwc(x.detach()).requires_grad_()
    ps = [t.detach().clone().requires_grad_() for t in (w1, b1, w2, b2)]
    y = ops.batch_norm(xm, bn2, True, True, pool=True)
    assert torch.equal(y.detach(), ops.batch_norm(xm.detach(), bn3, True, True))          # the normalised output itself is unchanged
    assert rel(y._tfb_pooled, y.detach().mean((1, 2))) < 1e-5
    n0 = len(_calls())
    out = ops.SEFn.apply(y, *ps)
    if _calls() is not _NOLOG:
        assert 'tfb_pool_hw_fwd' not in _calls()[n0:]
    assert rel(nchw(out), ref) < TOL
    check_grads([xm, bn2.weight, bn2.bias] + ps, [x, bn.weight, bn.bias, w1, b1, w2, b2], out, ref)


@pytest.mark.parametrize('shape', [(3, 10, 12, 72), (2, 48, 44, 72), (2, 10, 22, 576), (2, 5, 22, 1512), (2, 40, 56, 216)])
def test_batchnorm_squeeze_excite_one_node(shape):
    """ops.bn_se (BatchNorm + ReLU + squeeze-excite as one autograd node, the SE gradient folded into BatchNorm's backward kernels:
    tfb_bn_bwd_se) against torch, forward and all seven gradients. Shapes: small narrow / wide maps (slab column reduction), tall narrow
    maps (M >= 4096, C <= 512: the flat column reduction)."""
    from transfuser_b200 import ops
    N, H, W, C = shape
    Cr = 8
    x = (rnd(N, C, H, W, seed=1) * 2 + 0.3).requires_grad_()
    bn = torch.nn.BatchNorm2d(C).to(DEV)
    bn.weight.data = rnd(C, seed=2) * 0.3 + 1
    bn.bias.data = rnd(C, seed=3) * 0.2
    bn2 = torch.nn.BatchNorm2d(C).to(DEV)
    bn2.load_state_dict(bn.state_dict())
    fc1, fc2 = torch.nn.Conv2d(C, Cr, 1).to(DEV), torch.nn.Conv2d(Cr, C, 1).to(DEV)
    fc1.weight.data, fc1.bias.data = rnd(Cr, C, 1, 1, seed=4, scale=0.1), rnd(Cr, seed=5, scale=0.1)
    fc2.weight.data, fc2.bias.data = rnd(C, Cr, 1, 1, seed=6, scale=0.3), rnd(C, seed=7, scale=0.1)
    w1, b1, w2, b2 = [t.detach().clone().requires_grad_() for t in (fc1.weight, fc1.bias, fc2.weight, fc2.bias)]
    yb = F.relu(bn(x))
    ref = yb * torch.sigmoid(F.conv2d(F.relu(F.conv2d(yb.mean((2, 3), keepdim=True), w1, b1)), w2, b2))
    xm = nhwc(x.detach()).requires_grad_()
    assert ops.bn_se_ok(xm, bn2, fc1)
    n0 = len(_calls())
    out = ops.bn_se(xm, bn2, fc1, fc2)
    assert rel(nchw(out), ref) < TOL
    assert rel(bn2.running_mean, bn.running_mean) < TOL and rel(bn2.running_var, bn.running_var) < TOL
    check_grads([xm, bn2.weight, bn2.bias, fc1.weight, fc1.bias, fc2.weight, fc2.bias], [x, bn.weight, bn.bias, w1, b1, w2, b2], out, ref)
    if _calls() is not _NOLOG:
        log = _calls()[n0:]
        assert 'tfb_bn_bwd_se' in log and 'tfb_se_bwd_apply' not in log and 'tfb_pool_hw_fwd' not in log


@pytest.mark.parametrize('N,C,Cr', [(10, 1512, 144), (10, 576, 54), (16, 216, 18), (2, 72, 8), (1, 80, 3)])
def test_se_fused_mlp_backward(N, C, Cr):
    """tfb_se_mlp_bwd (two launches) against the eight-launch path it replaces and against torch autograd, at the RegNetY-3.2GF
    SE sizes (C up to 1512, reduction width = round(block input width / 4))."""
    from transfuser_b200 import ops
    if DEV == 'cpu' and C > 600:
        SKIPPED_ON_EMULATOR.append('test_se_fused_mlp_backward')
        pytest.skip('CPU emulation: the 576-channel case covers the same code in a third of the time')
    H, W = 3, 4
    x = rnd(N, C, H, W, seed=1).requires_grad_()
    w1, b1 = rnd(Cr, C, 1, 1, seed=2, scale=0.1).requires_grad_(), rnd(Cr, seed=3, scale=0.1).requires_grad_()
    w2, b2 = rnd(C, Cr, 1, 1, seed=4, scale=0.3).requires_grad_(), rnd(C, seed=5, scale=0.1).requires_grad_()
    ref = x * torch.sigmoid(F.conv2d(F.relu(F.conv2d(x.mean((2, 3), keepdim=True), w1, b1)), w2, b2))
    g = rnd(N, C, H, W, seed=6)
    want = torch.autograd.grad(ref, [x, w1, b1, w2, b2], g)
    got = {}
    old = ops.SE_FUSED_BWD
    try:
        for fused in (True, False):
            ops.SE_FUSED_BWD = fused
            xm = nhwc(x.detach()).requires_grad_()
            ps = [t.detach().clone().requires_grad_() for t in (w1, b1, w2, b2)]
            out = ops.SEFn.apply(xm, *ps)
            got[fused] = torch.autograd.grad(out, [xm] + ps, nhwc(g))
    finally:
        ops.SE_FUSED_BWD = old
    for i, (a, u, b_) in enumerate(zip(got[True], got[False], want)):
        if i == 0:
            a, u = nchw(a), nchw(u)
        assert rel(a.reshape(b_.shape), b_) < TOL, ('fused vs torch', i, rel(a.reshape(b_.shape), b_))
        assert rel(a, u) < 1e-5, ('fused vs unfused', i, rel(a, u))


@pytest.mark.parametrize('C,hw', [(72, ((40, 176), (64, 64))), (576, ((10, 44), (16, 16))), (1512, ((5, 22), (8, 8)))])
def test_gpt_tokens_and_view_quirk(C, hw):
    """Token build (adaptive avg pool + permute + pos_emb) and the reference's non-inverse `.view` on the way back
    (transfuser.py:346-364), then bilinear upsample + add (transfuser.py:154-157)."""
    from transfuser_b200 import ops
    (Hi, Wi), (Hl, Wl) = hw
    B = 2
    img, lid = rnd(B, C, Hi, Wi, seed=1).requires_grad_(), rnd(B, C, Hl, Wl, seed=2).requires_grad_()
    pos = rnd(1, 174, C, seed=3, scale=0.1).requires_grad_()
    ie, le = F.adaptive_avg_pool2d(img, (5, 22)), F.adaptive_avg_pool2d(lid, (8, 8))
    tok_ref = pos + torch.cat((ie.permute(0, 2, 3, 1).reshape(B, -1, C), le.permute(0, 2, 3, 1).reshape(B, -1, C)), dim=1)
    im, lm, pm = nhwc(img.detach()).requires_grad_(), nhwc(lid.detach()).requires_grad_(), pos.detach().clone().requires_grad_()
    tok = ops.TokensFn.apply(im, lm, pm, 5, 22, 8, 8, 0.0, 1)
    assert rel(tok, tok_ref) < 1e-5
    g = rnd(B, 174, C, seed=4)
    for a, b in zip(torch.autograd.grad(tok, [im, lm, pm], g), torch.autograd.grad(tok_ref, [img, lid, pos], g)):
        assert rel(nchw(a) if a.dim() == 4 and a.shape[-1] == C and a.shape != b.shape else a, b) < 1e-5
    # view quirk + upsample + add
    x = rnd(B, 174, C, seed=5).requires_grad_()
    io = x[:, :110, :].contiguous().view(B, -1, 5, 22)
    lo = x[:, 110:, :].contiguous().view(B, -1, 8, 8)
    ref_i = img + F.interpolate(io, size=(Hi, Wi), mode='bilinear', align_corners=False)
    ref_l = lid + F.interpolate(lo, size=(Hl, Wl), mode='bilinear', align_corners=False)
    xm = x.detach().clone().requires_grad_()
    oi, ol = ops.GptUpAddFn.apply(im, lm, xm, 5, 22, 8, 8)
    assert rel(nchw(oi), ref_i) < 1e-5 and rel(nchw(ol), ref_l) < 1e-5
    gi, gl = rnd(B, C, Hi, Wi, seed=6), rnd(B, C, Hl, Wl, seed=7)
    rg = torch.autograd.grad([ref_i, ref_l], [x, img, lid], [gi, gl])
    mg = torch.autograd.grad([oi, ol], [xm, im, lm], [nhwc(gi), nhwc(gl)])
    assert rel(mg[0], rg[0]) < 1e-5 and rel(nchw(mg[1]), rg[1]) < 1e-5 and rel(nchw(mg[2]), rg[2]) < 1e-5


@pytest.mark.parametrize('cfg', [(2, 8, 8, 64, 16, 16, False), (2, 64, 64, 3, 160, 160, True), (1, 5, 22, 64, 40, 176, False),
                                 (1, 40, 176, 8, 160, 704, False)])
def test_upsample(cfg):
    from transfuser_b200 import ops
    N, Hi, Wi, C, Ho, Wo, ac = cfg
    x = rnd(N, C, Hi, Wi, seed=1).requires_grad_()
    ref = F.interpolate(x, size=(Ho, Wo), mode='bilinear', align_corners=ac)
    xm = nhwc(x.detach()).requires_grad_()
    out = ops.upsample(xm, Ho, Wo, ac)
    assert rel(nchw(out), ref) < 1e-5
    check_grads([xm], [x], out, ref, tol=1e-5)


def test_image_prep_and_layout():
    from transfuser_b200 import ops
    img = torch.randint(0, 256, (2, 3, 16, 24), device=DEV).float()
    mean = torch.tensor([0.485, 0.456, 0.406], device=DEV).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225], device=DEV).view(1, 3, 1, 1)
    assert rel(nchw(ops.image_prep(img)), ((img / 255.0) - mean) / std) < 1e-6
    x = rnd(2, 5, 7, 9, seed=1)
    assert torch.equal(ops.nchw_to_nhwc(x), nhwc(x)) and torch.equal(ops.nhwc_to_nchw(nhwc(x)), x)


def test_losses():
    from transfuser_b200 import ops
    # weighted CE (pred_bev) and plain CE (semantic)
    logits = rnd(2, 3, 20, 24, seed=1).requires_grad_()
    tgt = torch.randint(0, 3, (2, 20, 24), device=DEV)
    w = torch.tensor([1., 1., 3.], device=DEV)
    lm = nhwc(logits.detach()).requires_grad_()
    ref, out = F.cross_entropy(logits, tgt, weight=w), ops.CrossEntropyFn.apply(lm, tgt, w, 'wsum', 1.0)
    assert abs(out.item() - ref.item()) < 1e-5 * abs(ref.item())
    (g1,), (g2,) = torch.autograd.grad(out * 0.7, lm), torch.autograd.grad(ref * 0.7, logits)
    assert rel(nchw(g1), g2) < 1e-5
    logits = rnd(2, 7, 12, 16, seed=2).requires_grad_()
    tgt = torch.randint(0, 7, (2, 12, 16), device=DEV)
    lm = nhwc(logits.detach()).requires_grad_()
    ref, out = 1.0 * F.cross_entropy(logits, tgt), ops.CrossEntropyFn.apply(lm, tgt, None, 'count', 1.0)
    assert abs(out.item() - ref.item()) < 1e-5 * abs(ref.item())
    (g1,), (g2,) = torch.autograd.grad(out, lm), torch.autograd.grad(ref, logits)
    assert rel(nchw(g1), g2) < 1e-5
    # depth: 10 * l1(sigmoid(x), t)
    x, t = rnd(2, 12, 16, seed=3).requires_grad_(), torch.rand(2, 12, 16, device=DEV)
    xm = x.detach().clone().requires_grad_()
    ref, out = 10.0 * F.l1_loss(torch.sigmoid(x), t), ops.L1Fn.apply(xm, t, True, 10.0)
    assert abs(out.item() - ref.item()) < 1e-5 * abs(ref.item())
    assert rel(torch.autograd.grad(out, xm)[0], torch.autograd.grad(ref, x)[0]) < 1e-5


@pytest.mark.parametrize('M,N,K', [(10, 37, 100), (10, 54, 216), (16, 145, 576), (3, 9, 1000), (10, 33, 1512)])
def test_small_m_gemm(M, N, K):
    """tfb_gemm_small_m (the squeeze-excite fc layers on pooled [B, C] vectors): every K-slice variant of the transB kernel (1, 2, 4, 8
    warps per output column), ragged N, both activations, and the !transB form."""
    from transfuser_b200._lib import call
    a, w, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5), rnd(N, seed=3)
    for act, fn in ((0, lambda t: t), (1, F.relu), (2, torch.sigmoid)):
        out = torch.empty(M, N, device=DEV)
        call('tfb_gemm_small_m', 1, M, N, K, a, K, w, K, out, N, b, act)
        assert rel(out, fn(a @ w.t() + b)) < TOL
    wt = w.t().contiguous()
    out = torch.empty(M, N, device=DEV)
    call('tfb_gemm_small_m', 0, M, N, K, a, K, wt, N, out, N, None, 0)
    assert rel(out, a @ wt) < TOL


def test_gru_waypoints():
    from transfuser_b200 import ops
    B = 5
    cell, outl = torch.nn.GRUCell(4, 64).to(DEV), torch.nn.Linear(64, 3).to(DEV)
    z0, tp = rnd(B, 64, seed=1).requires_grad_(), rnd(B, 2, seed=2) * 5
    z, x = z0, torch.zeros(B, 2, device=DEV)
    tpn = tp.clone(); tpn[:, 1] *= -1
    wps = []
    for _ in range(4):
        z = cell(torch.cat([x, tpn], dim=1), z)
        x = outl(z)[:, :2] + x
        wps.append(x)
    ref = torch.stack(wps, dim=1)
    ref = torch.cat((ref[:, :, :1] - 1.3, ref[:, :, 1:]), dim=2)
    zm = z0.detach().clone().requires_grad_()
    ps = [p.detach().clone().requires_grad_() for p in (cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh, outl.weight, outl.bias)]
    out = ops.GRUFn.apply(zm, tp, *ps, 4, 1.3)
    assert rel(out, ref) < 1e-5
    g = rnd(B, 4, 2, seed=3)
    rg = torch.autograd.grad(ref, [z0, cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh, outl.weight, outl.bias], g)
    mg = torch.autograd.grad(out, [zm] + ps, g)
    for a, b in zip(mg, rg):
        assert rel(a, b) < 1e-4
