"""Per-shape device time of every C-ABI entry point of one training step, each timed in its own CUDA graph (measurement tool).

One instrumented eager step records every call with its operands (transfuser_b200._lib.Profiler, keep_calls). The calls are grouped
by (entry point, shape key, CTA cap in force); one representative call per group is captured REP times back to back in a CUDA graph
and the graph is timed with CUDA events — device time per launch without the host launch path, operands warm in L2 when they fit
(so: a lower bound of the in-step time for small tensors). Printed per group: launches per step, us per launch, the time at the
binding roofline (max of FLOPs / tensor peak and algorithmic bytes / HBM peak), and the step-level excess (us - ideal) x launches,
sorted by that excess — the list of shapes worth tuning.

    python tools/shape_times.py [--top 60] [--gemm simt|bf16]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--top', type=int, default=70)
    ap.add_argument('--rep', type=int, default=10)
    args = ap.parse_args()
    import numpy as np
    import torch
    import bench
    from transfuser_b200 import _lib
    from transfuser_b200.config import TrainConfig
    from transfuser_b200.trainer import Trainer

    dev = torch.device('cuda', 0)
    torch.cuda.set_device(0)
    pk, _ = bench.peaks()
    tr = Trainer(TrainConfig(), dev, gemm_mode='bf16', lr=1e-4, seed=0, backbone='transFuser', raw_inputs=True)
    host = bench.make_host_batch(10, seed=100, torch=torch, np=np, backbone='transFuser', raw=True)
    h2d = lambda: {k: v.to(dev, non_blocking=True) for k, v in host.items()}
    for _ in range(3):
        tr.step(h2d())
    prof = _lib.Profiler(keep_calls=True)
    _lib.lib().profiler = prof
    tr.step(h2d())
    torch.cuda.synchronize()
    _lib.lib().profiler = None

    groups = {}
    cap = 0
    for name, a in prof.calls:
        if name == 'tfb_gemm_set_max_ctas':
            cap = a[0]
            continue
        if name in _lib.HOST_ONLY:
            continue
        key = prof._key(name, a)
        if key == name:      # no shape key: use the tensor sizes
            key = '%s %s' % (name, ','.join(str(t.numel()) for t in a if isinstance(t, torch.Tensor))[:60])
        if name.startswith('tfb_gemm_bf16_tc') or name.startswith('tfb_conv3x3_tc'):
            key += ' cap%d' % cap
        g = groups.setdefault(key, [name, a, cap, 0])
        g[3] += 1

    call = _lib.lib().call
    rows = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    side = torch.cuda.Stream()
    for key, (name, a, cap_g, n) in groups.items():
        try:
            call('tfb_gemm_set_max_ctas', cap_g)
            with torch.cuda.stream(side):
                call(name, *a)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(args.rep):
                    call(name, *a)
            g.replay()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(3):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (3 * args.rep)
        except Exception as e:  # noqa: BLE001
            print('skip %s: %r' % (key, e), file=sys.stderr)
            continue
        fl, by = prof._work(name, a)
        ideal = max(fl / (pk['bf16_tflops_sustained'] * 1e12), by / (pk['hbm_gbs'] * 1e9)) * 1e6
        rows.append((n * (us - ideal), n, us, ideal, key))
    call('tfb_gemm_set_max_ctas', 0)
    rows.sort(reverse=True)
    tot = sum(r[1] * r[2] for r in rows)
    print('sum over all groups: %.2f ms of device time per step (each launch timed alone, warm); ideal %.2f ms'
          % (tot / 1e3, sum(r[1] * r[3] for r in rows) / 1e3))
    print('%9s %5s %9s %9s  %s' % ('excess_us', 'n', 'us', 'ideal_us', 'entry point / shape'))
    for ex, n, us, ideal, key in rows[:args.top]:
        print('%9.0f %5d %9.2f %9.2f  %s' % (ex, n, us, ideal, key))


if __name__ == '__main__':
    main()
