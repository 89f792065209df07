"""Direct check of csrc/attn_tc.cu against torch on the GPU: forward y / lse and backward dq, dk, dv separately, every head
size of the four GPT scales, plus timings (CUDA events, L2 not flushed: the operands are L2 resident in the step as well)."""
import math
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from transfuser_b200 import _lib


def rel(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-30)).item()


def run(B, T, nh, hs, p_drop=0.0, in_bf16=False, time_it=False):
    C = nh * hs
    g = torch.Generator(device='cuda').manual_seed(hs + B)
    qkv = torch.randn(B * T, 3 * C, device='cuda', generator=g)
    dy = torch.randn(B * T, C, device='cuda', generator=g)
    scale = 1.0 / math.sqrt(hs)
    seed = torch.tensor([12345], dtype=torch.int64, device='cuda')
    y = torch.full((B * T, C), float('nan'), device='cuda')
    y16 = torch.empty((B * T, C), dtype=torch.bfloat16, device='cuda')
    lse = torch.full((B, nh, T), float('nan'), device='cuda')
    qin = qkv.bfloat16() if in_bf16 else qkv
    _lib.call('tfb_attn_fwd_tc', qin, int(in_bf16), B, T, nh, hs, y, y16, lse, scale, p_drop, seed, 7)
    torch.cuda.synchronize()
    # reference on the bf16-rounded operands (what the tensor cores see), fp32 math
    qr = qkv.bfloat16().float().requires_grad_()
    q, k, v = [qr[:, i * C:(i + 1) * C].view(B, T, nh, hs).transpose(1, 2) for i in range(3)]
    s = (q @ k.transpose(-2, -1)) * scale
    P = torch.softmax(s, dim=-1)
    out = {}
    if p_drop == 0.0:
        ref = (P @ v).transpose(1, 2).reshape(B * T, C)
        out['y'] = rel(y, ref)
        out['y16'] = rel(y16, ref)
        out['lse'] = rel(lse, torch.logsumexp(s, dim=-1))
        dqkv = torch.full((B * T, 3 * C), float('nan'), device='cuda')
        d16 = torch.empty((B * T, 3 * C), dtype=torch.bfloat16, device='cuda')
        dsum = torch.empty((B, nh, T), device='cuda')
        dy16 = torch.empty((B * T, C), dtype=torch.bfloat16, device='cuda')
        yin = y16 if in_bf16 else y
        _lib.call('tfb_attn_bwd_tc', qin, int(in_bf16), dy, yin, int(in_bf16), lse, dsum, dy16, B, T, nh, hs, dqkv, d16, scale, p_drop, seed, 7)
        torch.cuda.synchronize()
        (gref,) = torch.autograd.grad(ref, qr, dy.bfloat16().float())
        for i, n in enumerate('qkv'):
            out['d' + n] = rel(dqkv[:, i * C:(i + 1) * C], gref[:, i * C:(i + 1) * C])
        out['d16'] = rel(d16, gref)
        out['nan'] = int(torch.isnan(dqkv).sum().item() + torch.isnan(y).sum().item())
    else:
        # dropout: recover the mask from a run with v = identity-like probe is overkill; check keep rate and fwd/bwd consistency:
        # y(p) must equal (P * mask / (1-p)) v for SOME mask with keep rate ~ 1-p; test through linearity in v instead:
        y2 = torch.empty_like(y)
        _lib.call('tfb_attn_fwd_tc', qin, int(in_bf16), B, T, nh, hs, y2, None, lse, scale, p_drop, seed, 7)
        out['determinism'] = rel(y2, y)
        ref = (P @ v).transpose(1, 2).reshape(B * T, C)
        out['mean_ratio'] = (y * ref).sum().item() / (ref * ref).sum().item()     # E[dropout(P)] = P  => ~1
        # backward consistency: finite-difference-free check  <dy, d y/d v [dv]> = <dv_grad, dv>: y is linear in v for a fixed mask
        dqkv = torch.empty((B * T, 3 * C), device='cuda')
        dsum = torch.empty((B, nh, T), device='cuda')
        dy16 = torch.empty((B * T, C), dtype=torch.bfloat16, device='cuda')
        _lib.call('tfb_attn_bwd_tc', qin, int(in_bf16), dy, y, 0, lse, dsum, dy16, B, T, nh, hs, dqkv, None, scale, p_drop, seed, 7)
        qkv2 = qkv.clone()
        dv = torch.randn(B * T, C, device='cuda', generator=g).bfloat16().float() * 0.5
        qkv2[:, 2 * C:] = (qkv[:, 2 * C:].bfloat16().float() + dv)
        y3 = torch.empty_like(y)
        _lib.call('tfb_attn_fwd_tc', qkv2, 0, B, T, nh, hs, y3, None, lse, scale, p_drop, seed, 7)
        lhs = (dy.bfloat16().float() * (y3 - y)).sum().item()
        rhs = (dqkv[:, 2 * C:] * dv).sum().item()
        out['dv_linearity'] = abs(lhs - rhs) / max(abs(rhs), 1e-9)
    if time_it:
        for name, fn in (('fwd', lambda: _lib.call('tfb_attn_fwd_tc', qin, int(in_bf16), B, T, nh, hs, y, y16, lse, scale, p_drop, seed, 7)),
                         ('bwd', lambda: _lib.call('tfb_attn_bwd_tc', qin, int(in_bf16), dy, y, 0, lse, dsum, dy16, B, T, nh, hs, dqkv, d16, scale, p_drop, seed, 7))):
            if name == 'bwd' and p_drop > 0:
                continue
            for _ in range(3):
                fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
            out[name + '_us'] = round(e0.elapsed_time(e1) / 20 * 1e3, 1)
    return out


if __name__ == '__main__':
    torch.manual_seed(0)
    for (B, T, nh, hs) in [(2, 174, 4, 18), (2, 174, 4, 54), (2, 174, 4, 144), (2, 174, 4, 378), (1, 100, 2, 64), (3, 192, 1, 130),
                           (10, 174, 4, 378), (10, 174, 4, 144), (10, 174, 4, 54), (10, 174, 4, 18)]:
        for in16 in (False, True):
            r = run(B, T, nh, hs, time_it=(B == 10), in_bf16=in16)
            print('B%d T%d nh%d hs%d bf16_in=%d' % (B, T, nh, hs, in16), {k: (float('%.3g' % v) if isinstance(v, float) else v) for k, v in r.items()}, flush=True)
    print('bf16 inputs', run(2, 174, 4, 54, in_bf16=True), flush=True)
    print('dropout', run(2, 174, 4, 54, p_drop=0.1), run(10, 174, 4, 378, p_drop=0.1, time_it=True), flush=True)
