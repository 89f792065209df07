"""bf16 sidecars (ops._emit16 / _as16): in the bf16 tensor-core mode the producers of a GEMM / conv operand — BatchNorm apply, SE
gating, residual add(+ReLU), LayerNorm — write the bf16 copy of their output in the same pass, and the consumer takes it instead of
launching tfb_cast_bf16. Run here on the CPU emulation of the unchanged kernel sources (tests/cuda_emul/): the sidecar must be
BIT-identical to the cast of the fp32 output (so the step's numerics do not change), the fp32 output must not change, and the
consumers must pick it up. The tensor-core GEMM itself cannot be emulated: its host wrapper is replaced by a torch stand-in that
checks the operand it was handed. Test infrastructure only."""
import pytest
import torch
import torch.nn as nn

from cuda_emul import loader


@pytest.fixture()
def lib(monkeypatch):
    from transfuser_b200 import gemm
    lib = loader.patch_product(monkeypatch)
    gemm.set_mode('bf16')          # sidecars exist in the tensor-core mode only (conftest resets the mode after the test)
    yield lib


def _rnd(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def _cast(lib, x):
    y = torch.empty(x.shape, dtype=torch.bfloat16)
    lib.call('tfb_cast_bf16', x.contiguous(), y, x.numel())
    return y


def _same_bits(a, b):
    return torch.equal(a.view(torch.int16), b.view(torch.int16))


@pytest.mark.parametrize('relu', [False, True])
@pytest.mark.parametrize('shape', [(2, 5, 7, 24), (1, 3, 3, 72)])
def test_batchnorm_sidecar(lib, shape, relu):
    from transfuser_b200 import ops
    bn = nn.BatchNorm2d(shape[-1])
    with torch.no_grad():
        bn.weight.copy_(_rnd(shape[-1], seed=1).abs() + 0.5)
        bn.bias.copy_(_rnd(shape[-1], seed=2))
    x = _rnd(*shape, seed=3).requires_grad_(True)
    y = ops.batch_norm(x, bn, relu, True, emit16=True)
    s = y._tfb16
    assert s.dtype == torch.bfloat16 and s.shape == y.shape and not s.requires_grad
    assert _same_bits(s, _cast(lib, y.detach()))
    assert torch.equal(s.float(), y.detach().bfloat16().float())     # the emulated cast itself = torch's round-to-nearest-even
    y0 = ops.batch_norm(x, _clone_bn(bn), relu, True)
    assert not hasattr(y0, '_tfb16')
    assert torch.equal(y0.detach(), y.detach())                      # the fp32 output does not depend on the sidecar
    g = _rnd(*shape, seed=4)
    gx, = torch.autograd.grad(y, x, g)                               # backward takes the extra (None) gradient slot
    assert torch.isfinite(gx).all()
    # eval mode (running statistics) emits a sidecar too
    bn.eval()
    with torch.no_grad():
        ye = ops.batch_norm(x.detach(), bn, relu, False, emit16=True)
    assert _same_bits(ye._tfb16, _cast(lib, ye))


def _clone_bn(bn):
    c = nn.BatchNorm2d(bn.num_features)
    c.load_state_dict(bn.state_dict())
    return c


@pytest.mark.parametrize('n', [64, 67, 3])
@pytest.mark.parametrize('relu', [False, True])
def test_add_sidecar(lib, n, relu):
    from transfuser_b200 import ops
    a, b = _rnd(2, n, seed=5).requires_grad_(True), _rnd(2, n, seed=6).requires_grad_(True)
    y = ops.add(a, b, relu=relu, emit16=True)
    want = (a + b).detach()
    want = want.clamp_min(0) if relu else want
    assert torch.equal(y.detach(), want)
    assert _same_bits(y._tfb16, _cast(lib, y.detach()))              # incl. the scalar tail when n % 4 != 0
    ga, gb = torch.autograd.grad(y, (a, b), torch.ones_like(y))
    assert torch.equal(ga, gb)
    assert not hasattr(ops.add(a, b, relu=relu), '_tfb16')


@pytest.mark.parametrize('C', [24, 6])
def test_se_sidecar(lib, C):
    from transfuser_b200 import ops
    N, H, W, Cr = 2, 4, 5, 8
    x = _rnd(N, H, W, C, seed=7).requires_grad_(True)
    w1, b1 = _rnd(Cr, C, 1, 1, seed=8).requires_grad_(True), _rnd(Cr, seed=9).requires_grad_(True)
    w2, b2 = _rnd(C, Cr, 1, 1, seed=10).requires_grad_(True), _rnd(C, seed=11).requires_grad_(True)
    if C % 4:
        pytest.skip('SE pooling requires C % 4 == 0 (every RegNetY width)')
    y = ops.SEFn.apply(x, w1, b1, w2, b2, True)
    y0 = ops.SEFn.apply(x, w1, b1, w2, b2)
    assert torch.equal(y.detach(), y0.detach()) and not hasattr(y0, '_tfb16')
    assert _same_bits(y._tfb16, _cast(lib, y.detach()))
    grads = torch.autograd.grad(y, (x, w1, b1, w2, b2), _rnd(N, H, W, C, seed=12))
    grads0 = torch.autograd.grad(y0, (x, w1, b1, w2, b2), _rnd(N, H, W, C, seed=12))
    for a, b in zip(grads, grads0):
        assert torch.equal(a, b)


@pytest.mark.parametrize('R,C', [(5, 72), (3, 1512), (2, 7)])
def test_layernorm_sidecar(lib, R, C):
    from transfuser_b200 import ops
    ln = nn.LayerNorm(C)
    with torch.no_grad():
        ln.weight.copy_(_rnd(C, seed=13))
        ln.bias.copy_(_rnd(C, seed=14))
    x = _rnd(R, C, seed=15).requires_grad_(True)
    y = ops.layer_norm(x, ln, emit16=True)
    y0 = ops.layer_norm(x, ln)
    assert torch.equal(y.detach(), y0.detach()) and not hasattr(y0, '_tfb16')
    assert _same_bits(y._tfb16, _cast(lib, y.detach()))
    gx, = torch.autograd.grad(y, x, _rnd(R, C, seed=16))
    gx0, = torch.autograd.grad(y0, x, _rnd(R, C, seed=16))
    assert torch.equal(gx, gx0)


def test_no_sidecar_outside_bf16_mode(lib):
    from transfuser_b200 import gemm, ops
    gemm.set_mode('simt')
    a, b = _rnd(4, 8, seed=1), _rnd(4, 8, seed=2)
    assert not hasattr(ops.add(a, b, emit16=True), '_tfb16')
    gemm.set_mode('bf16')
    ops_sidecars = ops.SIDECARS
    try:
        ops.SIDECARS = False
        assert not hasattr(ops.add(a, b, emit16=True), '_tfb16')
    finally:
        ops.SIDECARS = ops_sidecars


def test_consumers_take_the_sidecar(lib, monkeypatch):
    """linear() / conv2d() 1x1 / the attention projections hand the producer's sidecar (same storage, reshaped) to the tensor-core
    GEMM and launch no cast for the activation; a tensor without a sidecar still gets its cast."""
    from transfuser_b200 import gemm as G, ops
    seen, operands = [], []

    def fake_gemm_bf16(a, b, out, trans_a=False, trans_b=False, bias=None, relu=False, alpha=1.0, beta=0.0, splits=1):
        seen.append(a)
        operands.append(b)
        A = a.float().t() if trans_a else a.float()
        Bm = b.float().t() if trans_b else b.float()
        r = alpha * (A @ Bm)
        if bias is not None:
            r = r + bias
        if relu:
            r = r.clamp_min(0)
        out.copy_(r if beta == 0.0 else r + beta * out)
        return out
    monkeypatch.setattr(G, 'gemm_bf16', fake_gemm_bf16)
    monkeypatch.setattr(G, 'weight_bf16', lambda w: w.detach().bfloat16())
    ln = nn.LayerNorm(64)
    x = _rnd(2, 20, 64, seed=20)
    w, b = _rnd(32, 64, seed=21).requires_grad_(True), _rnd(32, seed=22).requires_grad_(True)
    h = ops.layer_norm(x, ln, emit16=True)
    lib.log.clear()
    y = ops.linear(h, w, b)
    assert 'tfb_cast_bf16' not in lib.log
    assert seen[-1].data_ptr() == h._tfb16.data_ptr() and seen[-1].shape == (40, 64)
    want = h.detach().bfloat16().float().view(40, 64) @ w.detach().bfloat16().float().t() + b.detach()
    assert torch.allclose(y.detach().view(40, 32), want, rtol=1e-5, atol=1e-5)
    # NHWC 1x1 conv after BatchNorm
    bn = nn.BatchNorm2d(64)
    f = ops.batch_norm(_rnd(2, 4, 5, 64, seed=23), bn, True, True, emit16=True)
    lib.log.clear()
    ops.conv2d(f, _rnd(32, 64, 1, 1, seed=24).requires_grad_(True))
    assert 'tfb_cast_bf16' not in lib.log and seen[-1].data_ptr() == f._tfb16.data_ptr()
    # no sidecar -> one cast, as before
    lib.log.clear()
    ops.linear(_rnd(40, 64, seed=25), w, b)
    assert lib.log.count('tfb_cast_bf16') == 1
    # backward of the sidecar-fed linear: wgrad reads the saved sidecar
    seen.clear()
    hx = _rnd(2, 20, 64, seed=26).requires_grad_(True)
    h = ops.layer_norm(hx, ln, emit16=True)
    y = ops.linear(h, w, b, relu=True)
    gw, gh = torch.autograd.grad(y, (w, hx), _rnd(2, 20, 32, seed=27))
    assert any(t.data_ptr() == h._tfb16.data_ptr() for t in operands)      # wgrad: dW = g^T xs with xs = the sidecar
    assert torch.isfinite(gw).all() and torch.isfinite(gh).all()


def test_backward_sidecar_batchnorm_to_conv(lib, monkeypatch):
    """BatchNorm backward (bwd16) writes the bf16 copy of dx in its dx pass and offers it under (address, size); the 1x1 conv in
    front of the BatchNorm (no bias, no ReLU) takes it in its backward instead of launching tfb_grad_prep over dy. The copy is
    bit-identical to the cast of dx; a gradient nobody offered, or one that was modified in place afterwards, is not matched."""
    from transfuser_b200 import gemm as G, ops
    operands = []

    def fake_gemm_bf16(a, b, out, trans_a=False, trans_b=False, bias=None, relu=False, alpha=1.0, beta=0.0, splits=1):
        operands.append((a, b))
        A = a.float().t() if trans_a else a.float()
        Bm = b.float().t() if trans_b else b.float()
        r = alpha * (A @ Bm)
        if bias is not None:
            r = r + bias
        out.copy_(r.clamp_min(0) if relu else r)
        return out
    monkeypatch.setattr(G, 'gemm_bf16', fake_gemm_bf16)
    monkeypatch.setattr(G, 'weight_bf16', lambda w: w.detach().bfloat16())
    ops._BWD16.clear()
    N, H, W, Cin, Cout = 2, 4, 5, 64, 32
    x = _rnd(N, H, W, Cin, seed=30).requires_grad_(True)
    w = (_rnd(Cout, Cin, 1, 1, seed=31) / 8).requires_grad_(True)
    bn = nn.BatchNorm2d(Cout)
    y = ops.batch_norm(ops.conv2d(x, w), bn, True, True, bwd16=True)
    g = _rnd(N, H, W, Cout, seed=32)
    lib.log.clear()
    operands.clear()
    gx, gw = torch.autograd.grad(y, (x, w), g)
    assert 'tfb_grad_prep' not in lib.log and 'tfb_cast_bf16' not in lib.log
    assert not ops._BWD16                                            # the entry was consumed
    # reference: the same chain without the sidecar (grad_prep casts dy)
    x2, w2 = x.detach().clone().requires_grad_(True), w.detach().clone().requires_grad_(True)
    y2 = ops.batch_norm(ops.conv2d(x2, w2), _clone_bn(nn.BatchNorm2d(Cout)), True, True)
    lib.log.clear()
    gx2, gw2 = torch.autograd.grad(y2, (x2, w2), g)
    assert lib.log.count('tfb_grad_prep') == 1
    assert torch.equal(gx, gx2) and torch.equal(gw, gw2)             # same bf16 operand bits -> identical stand-in GEMM results
    # direct check of the kernel output + the registry rules
    xb = _rnd(6, 8, seed=33).view(1, 2, 3, 8).requires_grad_(True)
    yb = ops.batch_norm(xb, nn.BatchNorm2d(8), False, True, bwd16=True)
    dx, = torch.autograd.grad(yb, xb, _rnd(1, 2, 3, 8, seed=34))
    (t, t16, ver), = ops._BWD16.values()
    assert t.data_ptr() == dx.data_ptr() and _same_bits(t16, _cast(lib, dx))
    assert ops._take16(_rnd(1, 2, 3, 8, seed=35)) is None            # never offered
    dx.add_(1.0)                                                     # modified after the offer: stale, must not be matched
    assert ops._take16(dx) is None and not ops._BWD16
    # tick() drops whatever was not consumed
    ops._offer16(dx, torch.empty(dx.shape, dtype=torch.bfloat16))
    ops.tick('cpu')
    assert not ops._BWD16


def test_packed_conv_weight_registry(lib, monkeypatch):
    """ops._conv_tc_run keeps one persistent packed buffer per (weight, direction); tick() re-packs all of them in one launch and the
    convs of that step skip their own pack; a weight whose version moved, invalidate_packs() (fused AdamW / graph replay), or a
    step without tick() (eval) fall back to the per-call pack, so a stale pack is never read. The tensor-core conv itself is
    intercepted (not emulable): the stand-in records the packed weights it was handed."""
    import gc
    from transfuser_b200 import ops
    seen = []
    real_call = ops.call

    def call(name, *args):
        if name == 'tfb_conv3x3_tc':
            seen.append(args[1].clone())
            args[3].zero_()
            return
        return real_call(name, *args)
    monkeypatch.setattr(ops, 'call', call)
    monkeypatch.setattr(ops, 'PACK_BATCHED', True)
    ops._PACKS.clear()
    ops._PACK_STATE.update(sig=None, table=None)
    w = nn.Parameter(_rnd(64, 64, 3, 3, seed=40))
    w2 = nn.Parameter(_rnd(72, 24, 3, 3, seed=41))
    x, x2 = torch.zeros(1, 8, 16, 64, dtype=torch.bfloat16), torch.zeros(1, 8, 16, 72, dtype=torch.bfloat16)
    plan, plan2 = ops._conv_tc_plan(64, 64, 1), ops._conv_tc_plan(72, 72, 3)

    def fresh_pack(wt, pl, mode, groups):
        out = torch.empty((pl['gblocks'], pl['nchunks'], 9, pl['NB'], pl['KC']), dtype=torch.bfloat16)
        real_call('tfb_conv3x3_pack_weights', wt.detach(), out, wt.shape[0], wt.shape[1] * groups, groups, mode, pl['NB'], pl['KC'], pl['c_step'],
                  pl['nchunks'], pl['nb_real'], pl['gblocks'])
        return out

    def run(expect_single):
        lib.log.clear()
        seen.clear()
        ops._conv_tc_run(x, w, None, plan, 0, 64, 1, False)
        ops._conv_tc_run(x2, w2, None, plan2, 1, 72, 3, False)
        assert lib.log.count('tfb_conv3x3_pack_weights') == expect_single
        assert _same_bits(seen[0], fresh_pack(w, plan, 0, 1)) and _same_bits(seen[1], fresh_pack(w2, plan2, 1, 3))

    run(2)                                   # first use: registered + packed per call
    run(2)                                   # no tick() yet (eval-style use): still per call
    ops.tick('cpu')
    assert lib.log.count('tfb_conv3x3_pack_weights_batched') == 1
    run(0)                                   # inside the step: the batched pack is reused
    run(0)
    with torch.no_grad():
        w.mul_(2.0)                          # version bump (torch optimizer / load_state_dict): only this conv re-packs
    run(1)
    ops.tick('cpu')
    run(0)
    w2.data.mul_(0.5)                        # silent update (what the fused AdamW kernel does) + the invalidation that goes with it
    ops.invalidate_packs()
    run(2)
    table = ops._PACK_STATE['table']
    ops.tick('cpu')
    assert ops._PACK_STATE['table'] is table          # unchanged registry: the descriptor table is not rebuilt (graph capture safe)
    run(0)
    monkeypatch.setattr(ops, 'PACK_BATCHED', False)
    run(2)
    monkeypatch.setattr(ops, 'PACK_BATCHED', True)
    del w2
    gc.collect()
    w2 = nn.Parameter(_rnd(72, 24, 3, 3, seed=42))    # a new weight (possibly at the old address / id): never served the old pack
    ops.tick('cpu')
    run(1)
    assert len(ops._PACKS) == 2


@pytest.mark.parametrize('ac', [False, True])
def test_upsample_sidecar(lib, ac):
    from transfuser_b200 import ops
    x = _rnd(2, 5, 7, 12, seed=50).requires_grad_(True)
    y = ops.upsample(x, 20, 21, ac, emit16=True)
    y0 = ops.upsample(x, 20, 21, ac)
    assert torch.equal(y.detach(), y0.detach()) and not hasattr(y0, '_tfb16')
    assert _same_bits(y._tfb16, _cast(lib, y.detach()))
    g = _rnd(2, 20, 21, 12, seed=51)
    assert torch.equal(torch.autograd.grad(y, x, g)[0], torch.autograd.grad(y0, x, g)[0])


def test_gpt_up_add_sidecar(lib):
    from transfuser_b200 import ops
    img, lid, tok = _rnd(2, 10, 44, 8, seed=60), _rnd(2, 16, 16, 8, seed=61), _rnd(2, 174, 8, seed=62)
    oi, ol = ops.GptUpAddFn.apply(img, lid, tok, 5, 22, 8, 8)
    assert _same_bits(oi._tfb16, _cast(lib, oi)) and _same_bits(ol._tfb16, _cast(lib, ol))
    old = ops.SIDECARS
    try:
        ops.SIDECARS = False
        pi, pl = ops.GptUpAddFn.apply(img, lid, tok, 5, 22, 8, 8)
    finally:
        ops.SIDECARS = old
    assert torch.equal(pi, oi) and torch.equal(pl, ol) and not hasattr(pi, '_tfb16')
