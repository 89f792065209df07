"""Diagnostic: what the host CPU really offers on the GPU box (affinity, cgroup quota, OpenMP scaling of a small conv)."""
import os, time, torch
print('cpu_count', os.cpu_count(), 'affinity', len(os.sched_getaffinity(0)))
for f in ('/sys/fs/cgroup/cpu.max', '/sys/fs/cgroup/cpu/cpu.cfs_quota_us', '/sys/fs/cgroup/cpu/cpu.cfs_period_us'):
    try:
        print(f, open(f).read().strip())
    except Exception as e:
        print(f, 'n/a')
x = torch.randn(2, 72, 40, 176)
w = torch.randn(72, 72, 1, 1)
for th in (1, 8, 32, 128):
    torch.set_num_threads(th)
    torch.nn.functional.conv2d(x, w)
    t = time.time()
    for _ in range(20):
        torch.nn.functional.conv2d(x, w)
    print('threads', th, 'conv ms', (time.time() - t) / 20 * 1e3, flush=True)
