"""Parity evidence in the mode bench.py times (bf16 tensor-core mode): per-stage activations, all 11 losses and parameter
gradients of the CUDA path against the fp32 CPU oracle (B = 2, dropout off), plus the same numbers for the exact-fp32 SIMT mode.
Prints a table; tests/test_bf16.py::test_bf16_mode_vs_oracle asserts the bounds measured here."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
from test_model import Cfg, build, rel, _oracle_run
from oracle import torch_oracle as O
from transfuser_b200 import gemm


def run(mode, net0, batch, P, ref, taps):
    gemm.set_mode(mode)
    net = build().cuda().train()
    cb = {k: v.cuda() for k, v in batch.items()}
    mine = {}
    feats, grid, fused = net._model.forward_nhwc(cb['rgb'], torch.cat((cb['lidar'], cb['target_point_image']), dim=1), taps=mine)
    torch.cuda.synchronize()
    acts = {k: rel(mine[k].permute(0, 3, 1, 2), taps[k]) for k in sorted(mine)}
    acts['p2'] = rel(feats[0].permute(0, 3, 1, 2), taps['p2'])
    acts['img_grid'] = rel(grid.permute(0, 3, 1, 2), taps['img_grid'])
    acts['fused'] = rel(fused, taps['fused'])
    net = build().cuda().train()
    out = net(cb['rgb'], cb['lidar'], ego_waypoint=cb['ego_waypoint'], target_point=cb['target_point'],
              target_point_image=cb['target_point_image'], ego_vel=cb['ego_vel'], bev=cb['bev'], label=cb['label'],
              depth=cb['depth'], semantic=cb['semantic'])
    w = dict(zip(Cfg.detailed_losses, Cfg.detailed_losses_weights))
    sum(w[k] * out[k] for k in out).backward()
    torch.cuda.synchronize()
    losses = {k: abs(out[k].item() - ref[k].item()) / max(abs(ref[k].item()), 1e-12) for k in ref}
    g = {n: rel(p.grad, P[n].grad) for n, p in net.named_parameters() if not n.endswith('attn.key.bias')}
    cos = {n: torch.nn.functional.cosine_similarity(p.grad.flatten().double().cpu(), P[n].grad.flatten().double(), dim=0).item()
           for n, p in net.named_parameters() if not n.endswith('attn.key.bias')}
    return acts, losses, g, cos


if __name__ == '__main__':
    torch.manual_seed(0)
    net0 = build()
    batch = O.synthetic_batch(2, seed=3)
    P, ref, taps = _oracle_run(net0, batch, torch.float32)
    # the oracle with bf16-rounded tensor-core operands (oracle/torch_oracle.py BF16_OPERANDS): the product's rounding points on the CPU
    O.BF16_OPERANDS[0] = True
    P16, ref16, taps16 = _oracle_run(net0, batch, torch.float32)
    O.BF16_OPERANDS[0] = False
    for mode, (Pm, refm, tapsm) in (('simt', (P, ref, taps)), ('bf16', (P, ref, taps)), ('bf16 vs the bf16-operand oracle', (P16, ref16, taps16))):
        acts, losses, g, cos = run(mode.split()[0], net0, batch, Pm, refm, tapsm)
        print('==== mode', mode)
        for k, v in acts.items():
            print('activation %-10s rel err %.3e' % (k, v))
        for k, v in losses.items():
            print('loss %-22s rel err %.3e' % (k, v))
        e = np.array(list(g.values()))
        c = np.array(list(cos.values()))
        print('param grads vs fp32 oracle: rel err median %.3e p95 %.3e max %.3e | cosine median %.5f p5 %.5f min %.5f'
              % (np.median(e), np.percentile(e, 95), e.max(), np.median(c), np.percentile(c, 5), c.min()))
        for grp in ('heads|head.', 'pred_bev', 'seg_decoder', 'depth_decoder', 'join', 'decoder.', 'transformer4', 'transformer1', 'image_encoder.features.s4',
                    'image_encoder.features.s1', 'image_encoder.features.stem', 'lidar_encoder._model.s1'):
            keys = [n for n in g if any(t in n for t in grp.split('|'))]
            if keys:
                print('  %-32s n=%3d rel median %.3e  cos median %.5f min %.5f' % (grp, len(keys), np.median([g[n] for n in keys]),
                      np.median([cos[n] for n in keys]), min(cos[n] for n in keys)))
