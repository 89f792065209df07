"""The simple CUDA kernels added for SURVEY.md §8f (decode, geometric-fusion gather / pool, input preparation), executed
UNCHANGED on the CPU through tests/cuda_emul/cuda_emul.h (OS threads + barriers standing in for a thread block) and compared
with the oracle / plain torch ops. This is how their indexing and arithmetic were checked in a container without a GPU; the
`-m gpu` tests remain the parity tests proper. Test infrastructure only."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import bev_oracle
from oracle import pipeline_oracle as PO
from oracle import torch_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(os.path.dirname(HERE), 'transfuser_b200', 'csrc')
OUT = os.path.join(HERE, 'cuda_emul', '_build')

DRIVERS = {
    'geometric.cu': r'''
extern "C" void run_avgpool_fwd(const float* x, float* out, int N, int H, int W, int C, int gh, int gw, int vec) {
  if (vec == 4) emul_launch(dim3(3), dim3(64), [=] { avgpool_grid_fwd_kernel<4>(x, out, N, H, W, C, gh, gw); });
  else emul_launch(dim3(3), dim3(64), [=] { avgpool_grid_fwd_kernel<1>(x, out, N, H, W, C, gh, gw); });
}
extern "C" void run_avgpool_bwd(const float* dout, float* dx, int N, int H, int W, int C, int gh, int gw, int acc, int vec) {
  if (vec == 4) emul_launch(dim3(3), dim3(64), [=] { avgpool_grid_bwd_kernel<4>(dout, dx, N, H, W, C, gh, gw, acc); });
  else emul_launch(dim3(3), dim3(64), [=] { avgpool_grid_bwd_kernel<1>(dout, dx, N, H, W, C, gh, gw, acc); });
}
extern "C" void run_gather_fwd(const float* emb, const int64_t* pts, float* out, int B, int h, int w, int C, int M, int P) {
  emul_launch(dim3(2), dim3(64), [=] { gather_sum_fwd_kernel(emb, pts, out, B, h, w, C, M, P); });
}
extern "C" void run_gather_bwd(const float* dout, const int64_t* pts, float* demb, int B, int h, int w, int C, int M, int P) {
  emul_launch(dim3(2), dim3(64), [=] { gather_sum_bwd_kernel(dout, pts, demb, B, h, w, C, M, P); });
}
''',
    'decode.cu': r'''
extern "C" void run_decode(const float* preds, int B, int H, int W, int nb, int k, int npad, float ratio, float apc, float* boxes, int* labels) {
  emul_launch(dim3(B), dim3(kDecodeThreads), [=] { centernet_decode_kernel(preds, H, W, nb, k, npad, ratio, apc, boxes, labels); });
}
''',
    'input_prep.cu': r'''
extern "C" void run_draw(const double* tp, int B, float* out) {
  emul_launch(dim3(4, B), dim3(64), [=] { draw_target_point_kernel(tp, out); });
}
extern "C" void run_camera(const uint8_t* rgb, const uint8_t* depth, const uint8_t* seg, const int* shift, const uint8_t* lut, int B, int H, int W,
                           int ch, int cw, float* rgb_nchw, float* rgb_norm, float* depth_out, int64_t* seg_out) {
  emul_launch(dim3(3), dim3(64), [=] { camera_prep_kernel(rgb, depth, seg, shift, lut, B, H, W, ch, cw, rgb_nchw, rgb_norm, depth_out, seg_out); });
}
''',
    'losses.cu': r'''
extern "C" void run_targets(const float* label, int B, int K, float* tgt, int H, int W, float rw, float rh, int nb, int* count) {
  emul_launch(dim3(B), dim3(64), [=] { centernet_targets_kernel(label, K, tgt, H, W, rw, rh, nb, count); });
}
''',
    'gru_adamw.cu': r'''
extern "C" void run_adamw(float* p, float* g, float* m, float* v, int64_t n, double lr, double b1, double b2, double eps, double wd, int step,
                          const int* step_dev, float grad_scale, void* p_bf16, int zero_grad) {
  emul_launch(dim3(3), dim3(64), [=] { adamw_kernel(p, g, m, v, n, lr, b1, b2, (float)eps, wd, step, grad_scale, (__nv_bfloat16*)p_bf16, zero_grad, step_dev); });
}
extern "C" void run_gru_fwd(const float* z0, const float* tp, const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                            const float* w_out, const float* b_out, int B, int steps, float x_shift, float* wp, float* save) {
  emul_launch(dim3(B), dim3(G3), [=] { gru_fwd_kernel(z0, tp, w_ih, w_hh, b_ih, b_hh, w_out, b_out, steps, x_shift, wp, save); });
}
extern "C" void run_gru_bwd(const float* d_wp, const float* save, const float* w_ih, const float* w_hh, const float* w_out, int B, int steps,
                            float* dz0, float* dw_ih, float* dw_hh, float* db_ih, float* db_hh, float* dw_out, float* db_out) {
  emul_launch(dim3(B), dim3(G3), [=] { gru_bwd_kernel(d_wp, save, w_ih, w_hh, w_out, steps, dz0, dw_ih, dw_hh, db_ih, db_hh, dw_out, db_out); });
}
''',
    'bev_hist.cu': r'''
extern "C" void run_aligned(const float* pts, const double* T, const int* n_valid, int batch, int n_max, unsigned* counts, float* out) {
  emul_launch(dim3(2, batch), dim3(64), [=] { bev_scatter_aligned_kernel<float>(pts, T, n_valid, n_max, counts); });
  emul_launch(dim3(kGrid / 32, kGrid / 32, batch * 2), dim3(32, 8), [=] { bev_finalize_kernel(counts, out, batch); });
}
''',
}


def _emulated(cu):
    """Kernel part of csrc/<cu> (above its C-ABI entry points) + the emulation header + a driver, built with g++."""
    os.makedirs(OUT, exist_ok=True)
    text = open(os.path.join(CSRC, cu)).read()
    body = text[:text.index('}  // namespace') + len('}  // namespace')]
    body = body.replace('#include "common.cuh"', '#include "../cuda_emul.h"')
    assert '<<<' not in body
    src = os.path.join(OUT, cu.replace('.cu', '_emul.cpp'))
    lib = os.path.join(OUT, cu.replace('.cu', '_emul.so'))
    full = body + '\nusing namespace std;\n' + DRIVERS[cu]
    if not (os.path.exists(src) and open(src).read() == full and os.path.exists(lib)):
        open(src, 'w').write(full)
        r = subprocess.run(['g++', '-O1', '-std=c++17', '-ffp-contract=off', '-fPIC', '-shared', '-pthread', '-Wno-unknown-pragmas', '-Wno-attributes',
                            src, '-o', lib], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
    return ctypes.CDLL(lib)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize('shape,grid', [((2, 8, 16, 24), (4, 3)), ((1, 6, 10, 44), (5, 22)), ((2, 4, 8, 8), (8, 8))])
def test_avgpool_grid_kernels(shape, grid):
    lib = _emulated('geometric.cu')
    N, C, H, W = shape
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(1), requires_grad=True)
    want = F.adaptive_avg_pool2d(x, grid)
    xm = _nhwc(x.detach())
    out = torch.empty(N, grid[0], grid[1], C)
    vec = 4 if C % 4 == 0 else 1
    lib.run_avgpool_fwd(_p(xm), _p(out), N, H, W, C, grid[0], grid[1], vec)
    assert torch.equal(out.permute(0, 3, 1, 2), want.detach())          # same accumulation order as ATen's CPU kernel
    go = torch.randn(*want.shape, generator=torch.Generator().manual_seed(2))
    gw, = torch.autograd.grad(want, x, go)
    dx = torch.empty(N, H, W, C)
    gon = _nhwc(go)                                    # keep the operand alive across the ctypes call
    lib.run_avgpool_bwd(_p(gon), _p(dx), N, H, W, C, grid[0], grid[1], 0, vec)
    assert torch.allclose(dx.permute(0, 3, 1, 2), gw, rtol=0, atol=1e-7)
    base = torch.randn(N, H, W, C, generator=torch.Generator().manual_seed(3))
    acc = base.clone()
    lib.run_avgpool_bwd(_p(gon), _p(acc), N, H, W, C, grid[0], grid[1], 1, vec)
    assert torch.allclose(acc, base + dx, rtol=0, atol=1e-6)


@pytest.mark.parametrize('B,hw,HW,C', [(2, (5, 22), (8, 8), 16), (3, (8, 8), (5, 22), 8), (1, (4, 4), (2, 3), 4)])
def test_gather_sum_kernels(B, hw, HW, C):
    lib = _emulated('geometric.cu')
    g = torch.Generator().manual_seed(B * 7 + C)
    emb = torch.randn(B, C, *hw, generator=g, requires_grad=True)
    pts = torch.stack((torch.randint(0, hw[1], (B, *HW, 5), generator=g), torch.randint(0, hw[0], (B, *HW, 5), generator=g)), -1)
    pts[:, 0, 0] = 0
    pts[0, 0, 1, 0] = torch.tensor([-1, -1])            # python-style negative index = last column / row
    flat = pts.view(-1, 2)
    t = emb.permute(0, 2, 3, 1).contiguous()[:, flat[:, 1], flat[:, 0]].view(B, B, *HW, 5, -1)       # geometric_fusion.py:145
    want = torch.diagonal(t, 0).permute(4, 3, 0, 1, 2).contiguous().sum(-1)
    em = _nhwc(emb.detach())
    out = torch.empty(B, HW[0], HW[1], C)
    lib.run_gather_fwd(_p(em), _p(pts), _p(out), B, hw[0], hw[1], C, HW[0] * HW[1], 5)
    assert torch.allclose(out.permute(0, 3, 1, 2), want.detach(), rtol=0, atol=1e-6)
    go = torch.randn(*want.shape, generator=g)
    gw, = torch.autograd.grad(want, emb, go)
    demb = torch.zeros(B, hw[0], hw[1], C)               # the entry point memsets before the launch
    gon = _nhwc(go)
    lib.run_gather_bwd(_p(gon), _p(pts), _p(demb), B, hw[0], hw[1], C, HW[0] * HW[1], 5)
    assert torch.allclose(demb.permute(0, 3, 1, 2), gw, rtol=1e-5, atol=1e-5)
    # out-of-range correspondences are skipped
    bad = pts.clone()
    bad[0, 0, 0, 0] = torch.tensor([hw[1], 0])
    out2 = torch.empty_like(out)
    lib.run_gather_fwd(_p(em), _p(bad), _p(out2), B, hw[0], hw[1], C, HW[0] * HW[1], 5)
    assert torch.allclose(out2[0, 0, 0], out[0, 0, 0] - em[0, pts[0, 0, 0, 0, 1], pts[0, 0, 0, 0, 0]], atol=1e-5)


@pytest.mark.parametrize('case,H,W', [(0, 64, 64), (1, 64, 64), (2, 64, 64), (3, 64, 64), (0, 24, 40)])
def test_centernet_decode_kernel(case, H, W):
    lib = _emulated('decode.cu')
    g = torch.Generator().manual_seed(40 + case)
    B = 2
    heat_logit = torch.randn(B, 1, H, W, generator=g) * 2
    if case == 1:
        heat_logit = (heat_logit * 2).round() / 2
    if case == 2:
        heat_logit = heat_logit - 8 * (torch.rand(B, 1, H, W, generator=g) < 0.999)
    if case == 3:
        heat_logit[:, :, 0, :] = 30.0
    rest = [torch.randn(B, c, H, W, generator=g) for c in (2, 2, 12, 1, 1, 2)]
    want, want_labels = O.decode_heatmap([heat_logit.sigmoid()] + rest, 12, stable=True)
    raw = _nhwc(torch.cat([heat_logit] + rest, dim=1))
    boxes = torch.empty(B, 100, 8)
    labels = torch.empty(B, 100, dtype=torch.int32)
    npad = 2
    while npad < H * W:
        npad <<= 1
    lib.run_decode(_p(raw), B, H, W, 12, 100, npad, ctypes.c_float(4.0), ctypes.c_float(np.float32(2.0 * np.pi / 12)), _p(boxes), _p(labels))
    assert torch.equal(labels.long(), want_labels)
    assert torch.equal(boxes[..., 6], want[..., 6])
    assert torch.allclose(boxes, want, rtol=1e-6, atol=1e-5), (boxes - want).abs().amax(dim=(0, 1))


def test_target_point_kernel():
    lib = _emulated('input_prep.cu')
    pts = [(x, y) for x in (-16.2, -16.0, 15.9, 16.0, 16.1, 0.3) for y in (-1.4, -1.3, 30.6, 30.7, 30.8, 7.77)] + [(1e12, -1e12), (float('nan'), 0.0)]
    tp = torch.tensor(pts, dtype=torch.float64)
    out = torch.empty(len(pts), 1, 256, 256)
    lib.run_draw(_p(tp), len(pts), _p(out))
    for i, p in enumerate(pts):
        assert np.array_equal(out[i].numpy(), PO.draw_target_point(np.array(p)).astype(np.float32)), p


def test_camera_prep_kernel():
    lib = _emulated('input_prep.cu')
    conv = [0, 1, 2, 3, 4, 5, 6, 4, 3, 0, 2, 1, 5, 6, 0, 1, 2, 3, 4, 5, 6, 0, 1]
    H, W, crop = 40, 240, (32, 176)
    fs = [PO.synthetic_frame(s, H=H, W=W) for s in (0, 1, 2)]
    for f, deg in zip(fs, (5.0, -7.9, 0.0)):            # |shift| <= (240 - 176) / 2: the product's host check rejects more
        f['degree'] = deg
    shifts = [int(f['degree'] / 60 * W) for f in fs]
    B = len(fs)
    rgb, depth = (torch.from_numpy(np.stack([f[k] for f in fs])) for k in ('rgb', 'depth'))
    seg = torch.from_numpy(np.stack([f['seg'] for f in fs]))
    lut = torch.zeros(256, dtype=torch.uint8)
    lut[:len(conv)] = torch.tensor(conv, dtype=torch.uint8)
    o_rgb, o_norm = torch.empty(B, 3, *crop), torch.empty(B, *crop, 3)
    o_depth, o_seg = torch.empty(B, *crop), torch.empty(B, *crop, dtype=torch.int64)
    shift_t = torch.tensor(shifts, dtype=torch.int32)
    lib.run_camera(_p(rgb), _p(depth), _p(seg), _p(shift_t), _p(lut), B, H, W, crop[0], crop[1],
                   _p(o_rgb), _p(o_norm), _p(o_depth), _p(o_seg))
    for b, f in enumerate(fs):
        c = PO.crop_rgb(f['rgb'], crop, shifts[b])
        assert np.array_equal(o_rgb[b].numpy(), c.astype(np.float32))
        assert np.array_equal(o_norm[b].numpy(), PO.normalize_nhwc(c))
        assert np.array_equal(o_depth[b].numpy(), PO.depth_from_rgb(PO.crop_rgb(f['depth'], crop, shifts[b])).astype(np.float32))
        assert np.array_equal(o_seg[b].numpy(), PO.seg_classes(f['seg'], conv, crop, shifts[b]).astype(np.int64))


def test_aligned_histogram_kernel():
    lib = _emulated('bev_hist.cu')
    fs = [PO.synthetic_frame(s, n_points=1500) for s in (0, 1)]
    fs[1]['degree'] = 0.0
    pts = torch.from_numpy(np.stack([f['points'] for f in fs]))
    Ts = [PO.align_transform(f['ego_matrix_0'], f['ego_matrix_1'], f['degree']) for f in fs]
    T = torch.from_numpy(np.stack(Ts)).reshape(2, 16).contiguous()
    n_valid = torch.tensor([1500, 1200], dtype=torch.int32)
    counts = torch.zeros(2, 2, 256, 256, dtype=torch.int32)       # the entry point memsets before the launch
    out = torch.empty(2, 2, 256, 256)
    lib.run_aligned(_p(pts), _p(T), _p(n_valid), 2, 1500, _p(counts), _p(out))
    for b, f in enumerate(fs):
        n = int(n_valid[b])
        want = bev_oracle.lidar_to_histogram_features(PO.align_points(f['points'][:n], Ts[b]))
        assert np.array_equal(out[b].numpy(), want)


def test_input_pipeline_host_code_over_emulated_kernels(monkeypatch):
    """transfuser_b200.pipeline.InputPipeline.prepare executed end to end in this container: its three C-ABI calls are routed
    to the emulated kernels (same argument order as the entry points), everything else is the product's own host code."""
    from transfuser_b200 import pipeline
    from transfuser_b200.config import TrainConfig
    inp, bev = _emulated('input_prep.cu'), _emulated('bev_hist.cu')
    conv = [0, 1, 2, 3, 4, 5, 6, 4, 3, 0, 2, 1, 5, 6, 0, 1, 2, 3, 4, 5, 6, 0, 1]
    calls = []

    def call(name, *a):
        calls.append(name)
        if name == 'tfb_camera_prep':
            inp.run_camera(*[_p(x) if (x is None or isinstance(x, torch.Tensor)) else x for x in a])
        elif name == 'tfb_draw_target_point':
            inp.run_draw(_p(a[0]), a[1], _p(a[2]))
        elif name == 'tfb_bev_histogram_aligned':
            points, is_f64, T, n_valid, B, n_max, counts, out = a
            assert is_f64 == 0
            counts.zero_()
            bev.run_aligned(_p(points), _p(T), _p(n_valid), B, n_max, _p(counts), _p(out))
        else:
            raise AssertionError(name)

    monkeypatch.setattr(pipeline._lib, 'call', call)
    monkeypatch.setattr(pipeline.InputPipeline, '_require_cuda', lambda self: None)
    H, W, crop = 40, 240, (32, 176)
    fs = [PO.synthetic_frame(s, H=H, W=W, n_points=800) for s in (3, 4)]
    for f, deg in zip(fs, (6.0, -3.0)):
        f['degree'] = deg
    raw = dict(rgb=torch.from_numpy(np.stack([f['rgb'] for f in fs])), depth=torch.from_numpy(np.stack([f['depth'] for f in fs])),
               seg=torch.from_numpy(np.stack([f['seg'] for f in fs])),
               crop_shift=torch.tensor([pipeline.crop_shift_pixels(f['degree'], W // 3, 1) for f in fs], dtype=torch.int32),
               points=torch.from_numpy(np.stack([f['points'] for f in fs])),
               transforms=torch.from_numpy(np.stack([pipeline.align_transform(f['ego_matrix_0'], f['ego_matrix_1'], f['degree']) for f in fs])),
               target_point=torch.from_numpy(np.stack([f['target_point'] for f in fs])))
    pipe = pipeline.InputPipeline(TrainConfig(converter=conv), 'cpu', crop=crop)
    out = pipe.prepare(raw)
    assert calls == ['tfb_camera_prep', 'tfb_bev_histogram_aligned', 'tfb_draw_target_point']
    for b, f in enumerate(fs):
        shift = int(f['degree'] / 60 * (W // 3))
        T = PO.align_transform(f['ego_matrix_0'], f['ego_matrix_1'], f['degree'])
        assert np.array_equal(out['rgb'][b].numpy(), PO.crop_rgb(f['rgb'], crop, shift).astype(np.float32))
        assert np.array_equal(out['depth'][b].numpy(), PO.depth_from_rgb(PO.crop_rgb(f['depth'], crop, shift)).astype(np.float32))
        assert np.array_equal(out['semantic'][b].numpy(), PO.seg_classes(f['seg'], conv, crop, shift).astype(np.int64))
        assert np.array_equal(out['lidar'][b].numpy(), bev_oracle.lidar_to_histogram_features(PO.align_points(f['points'], T)))
        assert np.array_equal(out['target_point_image'][b].numpy(), PO.draw_target_point(f['target_point']).astype(np.float32))
        assert np.array_equal(out['target_point'][b].numpy(), f['target_point'].astype(np.float32))
    norm = pipe.prepare(raw, normalized_nhwc=True)['rgb']
    assert getattr(norm, '_tfb_nhwc_normalized', False) and norm.shape == (2, crop[0], crop[1], 3)
    assert np.array_equal(norm[0].numpy(), PO.normalize_nhwc(PO.crop_rgb(fs[0]['rgb'], crop, int(fs[0]['degree'] / 60 * (W // 3)))))
    # host-side rejection of a crop that leaves the frame
    bad = dict(raw, crop_shift=torch.tensor([40, 0], dtype=torch.int32))
    with pytest.raises(ValueError):
        pipe.prepare(bad)


def test_autograd_wrappers_over_emulated_kernels(monkeypatch):
    """ops.AvgPoolGridFn / ops.GatherSumFn / ops.centernet_decode (the product's host wrappers: shapes, saved tensors, backward
    plumbing) executed on CPU tensors with their C-ABI calls routed to the emulated kernels."""
    from transfuser_b200 import ops
    geo, dec = _emulated('geometric.cu'), _emulated('decode.cu')

    def call(name, *a):
        ptr = [_p(x) if (x is None or isinstance(x, torch.Tensor)) else x for x in a]
        if name == 'tfb_avgpool_grid_fwd':
            geo.run_avgpool_fwd(*ptr, 4 if a[5] % 4 == 0 else 1)
        elif name == 'tfb_avgpool_grid_bwd':
            geo.run_avgpool_bwd(*ptr, 4 if a[5] % 4 == 0 else 1)
        elif name == 'tfb_gather_sum_fwd':
            geo.run_gather_fwd(*ptr)
        elif name == 'tfb_gather_sum_bwd':
            a[2].zero_()
            geo.run_gather_bwd(*ptr)
        elif name == 'tfb_centernet_decode':
            preds, B, H, W, nb, k, ratio, boxes, labels = a
            npad = 2
            while npad < H * W:
                npad <<= 1
            dec.run_decode(_p(preds), B, H, W, nb, k, npad, ctypes.c_float(ratio), ctypes.c_float(np.float32(2.0 * np.pi / nb)), _p(boxes), _p(labels))
        else:
            raise AssertionError(name)

    monkeypatch.setattr(ops, 'call', call)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 8, 16, 24, generator=g, requires_grad=True)
    xm = _nhwc(x.detach()).requires_grad_()
    pts = torch.stack((torch.randint(0, 3, (2, 5, 6, 5), generator=g), torch.randint(0, 4, (2, 5, 6, 5), generator=g)), -1)
    # product composition: pool to a 4x3 grid, gather 5 correspondences per cell of a 5x6 grid
    mine = ops.gather_sum(ops.avgpool_grid(xm, 4, 3), pts)
    pooled = F.adaptive_avg_pool2d(x, (4, 3))
    want = torch.stack([pooled[b].permute(1, 2, 0)[pts[b, ..., 1], pts[b, ..., 0]].sum(2) for b in range(2)])
    assert mine.shape == want.shape == (2, 5, 6, 8) and torch.allclose(mine, want, atol=1e-6)
    go = torch.randn(*want.shape, generator=g)
    gm, = torch.autograd.grad(mine, xm, go)
    gw, = torch.autograd.grad(want, x, go)
    assert torch.allclose(gm.permute(0, 3, 1, 2), gw, atol=1e-5)
    with pytest.raises(RuntimeError):
        ops.gather_sum(xm, pts.int())
    heat = torch.randn(2, 1, 64, 64, generator=g)
    rest = [torch.randn(2, c, 64, 64, generator=g) for c in (2, 2, 12, 1, 1, 2)]
    boxes, labels = ops.centernet_decode(_nhwc(torch.cat([heat] + rest, 1)), 12, 100, 4.0)
    want_boxes, want_labels = O.decode_heatmap([heat.sigmoid()] + rest, 12, stable=True)
    assert labels.dtype == torch.int64 and torch.equal(labels, want_labels) and torch.allclose(boxes, want_boxes, rtol=1e-6, atol=1e-5)
    with pytest.raises(RuntimeError):
        ops.centernet_decode(torch.zeros(1, 64, 64, 20), 12)


def test_centernet_targets_kernel_edge_cases():
    """LidarCenterNetHead.get_targets (model.py:285-374) as rasterised by csrc/losses.cu, against the oracle (itself pinned to
    the reference): empty label sets, the maximum of 20 boxes, boxes on the map border, coincident centres (the later box
    wins the regression targets, the heatmap keeps the maximum), yaw angles on bin boundaries and outside [-pi, pi]."""
    lib = _emulated('losses.cu')
    g = torch.Generator().manual_seed(9)
    label = torch.zeros(5, 20, 7)
    r = lambda *s: torch.rand(*s, generator=g)
    label[1, :, 0:2] = r(20, 2) * 253 + 1                      # sample 1: 20 random boxes
    label[1, :, 2:4] = r(20, 2) * 32 + 8
    label[1, :, 4] = r(20) * 2 * np.pi - np.pi
    label[1, :, 5] = r(20) * 8
    label[1, :, 6] = (r(20) < 0.5).float()
    border = torch.tensor([[0.2, 0.3], [255.9, 0.1], [0.4, 255.8], [255.7, 255.9], [128.0, 0.0], [3.99, 251.9]])
    label[2, :6, 0:2] = border                                 # sample 2: centres in the first / last cells
    label[2, :6, 2:4] = torch.tensor([[60., 30.], [8., 8.], [100., 4.], [2., 2.], [16., 90.], [33., 12.]])
    label[2, :6, 4] = torch.tensor([0.0, np.pi, -np.pi, 2 * np.pi / 12, -2 * np.pi / 12 / 2, 7.5])
    label[2, :6, 5] = 1.0
    label[3, 0] = torch.tensor([100.3, 77.7, 20., 10., 0.3, 2., 1.])       # sample 3: two boxes in the same cell + a gap row
    label[3, 2] = torch.tensor([101.9, 78.1, 44., 30., -2.9, 5., 0.])
    label[4, 7] = torch.tensor([200., 20., 12., 24., 3.0, 0., 0.])         # sample 4: a single box after zero rows
    want, avg = O.centernet_targets(label, O.Cfg)
    tgt = torch.empty(5, 10, 64, 64)
    count = torch.zeros(1, dtype=torch.int32)                              # the entry point memsets it
    lib.run_targets(_p(label), 5, 20, _p(tgt), 64, 64, ctypes.c_float(64 / 256), ctypes.c_float(64 / 256), 12, _p(count))
    assert max(1, int(count)) == avg
    assert torch.allclose(tgt[:, 0:1], want['heat'], rtol=0, atol=1e-6) and torch.equal(tgt[:, 0:1] == 1, want['heat'] == 1)
    assert torch.equal(tgt[:, 1:3], want['wh']) and torch.equal(tgt[:, 3:5], want['offset'])
    assert torch.allclose(tgt[:, 5:6], want['yaw_res'], rtol=0, atol=1e-6)
    assert torch.equal(tgt[:, 6:7], want['velocity']) and torch.equal(tgt[:, 7], want['weight'][:, 0])
    assert torch.equal(tgt[:, 8].long(), want['yaw_cls']) and torch.equal(tgt[:, 9].long(), want['brake'])
    assert int((tgt[0] != 0).sum()) == 0


def test_adamw_kernel_matches_torch_adamw():
    """csrc/gru_adamw.cu's fused AdamW vs torch.optim.AdamW over 4 steps (n not a multiple of 4: vector body + scalar tail),
    the bf16 weight mirror (round-to-nearest-even of the updated fp32 value), the device-resident step counter used under
    CUDA-graph replay, gradient scaling (1/world) and the fused zero_grad."""
    lib = _emulated('gru_adamw.cu')
    n = 1003
    g0 = torch.Generator().manual_seed(0)
    p0 = torch.randn(n, generator=g0)
    ref = p0.clone().requires_grad_()
    opt = torch.optim.AdamW([ref], lr=1e-2, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    p, m, v = p0.clone(), torch.zeros(n), torch.zeros(n)
    pd, md, vd = p0.clone(), torch.zeros(n), torch.zeros(n)            # second copy driven by the device step counter
    mirror = torch.zeros(n, dtype=torch.bfloat16)
    step_dev = torch.zeros(1, dtype=torch.int32)
    f, d = ctypes.c_float, ctypes.c_double
    for step in range(1, 5):
        grad = torch.randn(n, generator=g0) * (10.0 ** (step - 3))
        ref.grad = grad.clone()
        opt.step()
        g = (grad * 4).contiguous()                                    # summed over 4 ranks, grad_scale = 1/4
        lib.run_adamw(_p(p), _p(g), _p(m), _p(v), ctypes.c_int64(n), d(1e-2), d(0.9), d(0.999), d(1e-8), d(1e-2), step, None, f(0.25), _p(mirror), 1)
        assert int((g != 0).sum()) == 0                                # zero_grad fused
        step_dev += 1                                                  # tfb_step_tick
        g2 = (grad * 4).contiguous()
        lib.run_adamw(_p(pd), _p(g2), _p(md), _p(vd), ctypes.c_int64(n), d(1e-2), d(0.9), d(0.999), d(1e-8), d(1e-2), 0, _p(step_dev), f(0.25), None, 0)
        assert torch.equal(g2, grad * 4)
        assert torch.allclose(p, ref.detach(), rtol=2e-6, atol=2e-7), (step, (p - ref.detach()).abs().max())
        assert torch.equal(p, pd) and torch.equal(m, md) and torch.equal(v, vd)
        assert torch.equal(mirror, p.bfloat16())
    st = opt.state[ref]
    # torch >= 2.0 updates exp_avg with lerp (m + (g - m)(1 - b1)); the kernel keeps the reference-era b1*m + (1-b1)*g: same value up
    # to fp32 rounding (cancellation when g ~ m). exp_avg_sq uses the same formula on both sides.
    assert torch.allclose(m, st['exp_avg'], rtol=1e-5, atol=1e-7), (m - st['exp_avg']).abs().max()
    assert torch.allclose(v, st['exp_avg_sq'], rtol=2e-6, atol=1e-14), ((v - st['exp_avg_sq']).abs() / st['exp_avg_sq']).max()


def test_gru_kernels_match_torch_gru_rollout():
    """forward_gru (model.py:611-646): 4 autoregressive GRUCell steps + Linear + cumulative sum; forward and full BPTT."""
    lib = _emulated('gru_adamw.cu')
    torch.manual_seed(5)
    B, steps = 3, 4
    cell, outl = torch.nn.GRUCell(4, 64), torch.nn.Linear(64, 3)
    z0 = torch.randn(B, 64, requires_grad=True)
    tp = torch.randn(B, 2) * 5
    z, x = z0, torch.zeros(B, 2)
    tpn = tp.clone()
    tpn[:, 1] *= -1
    wps = []
    for _ in range(steps):
        z = cell(torch.cat([x, tpn], dim=1), z)
        x = outl(z)[:, :2] + x
        wps.append(x)
    want = torch.stack(wps, dim=1)
    want = torch.cat((want[:, :, :1] - 1.3, want[:, :, 1:]), dim=2)
    ps = [t.detach().contiguous() for t in (cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh, outl.weight, outl.bias)]
    wp, save = torch.empty(B, steps, 2), torch.empty(B, steps, 5 * 64 + 4)
    z0c = z0.detach().contiguous()
    lib.run_gru_fwd(_p(z0c), _p(tp), *[_p(t) for t in ps], B, steps, ctypes.c_float(1.3), _p(wp), _p(save))
    assert torch.allclose(wp, want.detach(), rtol=1e-5, atol=1e-5)
    d = torch.randn(B, steps, 2)
    grads = torch.autograd.grad(want, [z0, cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh, outl.weight, outl.bias], d)
    dz0 = torch.empty(B, 64)
    outs = [torch.zeros_like(t) for t in ps]                           # the entry point memsets the parameter gradients
    lib.run_gru_bwd(_p(d), _p(save), _p(ps[0]), _p(ps[1]), _p(ps[4]), B, steps, _p(dz0), *[_p(t) for t in outs])
    for a, b, name in zip([dz0] + outs, grads, ('z0', 'w_ih', 'w_hh', 'b_ih', 'b_hh', 'w_out', 'b_out')):
        if name in ('w_out', 'b_out'):
            a, b = a[:2], b[:2]                                         # the third output row never reaches the waypoints
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-5), (name, (a - b).abs().max())
