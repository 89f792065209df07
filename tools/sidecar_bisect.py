"""Root-cause tool for tests/test_bf16.py::test_bf16_sidecars_do_not_change_the_step (round-1 driver failure): run-to-run noise of
the bf16 step with sidecars off/off and on/on, then on vs off, then each forward producer's sidecar disabled in turn."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from test_model import build
from oracle import torch_oracle as O
from transfuser_b200 import gemm, ops

batch = {k: v.cuda() for k, v in O.synthetic_batch(2, seed=4).items()}
gemm.set_mode('bf16')


def run(on):
    ops.SIDECARS = on
    net = build().cuda().train()
    out = net(batch['rgb'], batch['lidar'], ego_waypoint=batch['ego_waypoint'], target_point=batch['target_point'],
              target_point_image=batch['target_point_image'], ego_vel=batch['ego_vel'], bev=batch['bev'], label=batch['label'],
              depth=batch['depth'], semantic=batch['semantic'])
    sum(out.values()).backward()
    torch.cuda.synchronize()
    return {k: v.item() for k, v in out.items()}


def diff(a, b):
    return max(abs(a[k] - b[k]) / max(abs(b[k]), 0.1) for k in a), max(a, key=lambda k: abs(a[k] - b[k]) / max(abs(b[k]), 0.1))


r = [run(False) for _ in range(3)] + [run(True) for _ in range(3)]
print('off/off noise', diff(r[0], r[1]), diff(r[0], r[2]))
print('on/on noise  ', diff(r[3], r[4]), diff(r[3], r[5]))
print('on vs off    ', diff(r[3], r[0]), diff(r[4], r[1]))
print('off', r[0])
print('on ', r[3])
