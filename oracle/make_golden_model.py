"""ORACLE / TEST INFRASTRUCTURE ONLY. Runs the VERBATIM reference (oracle/ref_import.py) on seeded weights / inputs and
stores its outputs as the small fixture tests/golden/model_golden.npz: the 11 losses, activation statistics of the
backbone outputs and a handful of parameter-gradient norms. Weights / inputs are regenerated from seeds by
oracle/torch_oracle.py (deterministic_state, synthetic_batch), so the fixture stays a few hundred bytes."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import, torch_oracle as O  # noqa: E402

GRAD_KEYS = ['head.heatmap_head.2.weight', 'pred_bev.0.weight', 'seg_decoder.deconv3.2.weight', 'join.0.weight', 'decoder.weight_hh',
             '_model.up_conv3.weight', '_model.change_channel_conv_image.weight']
WEIGHT_SEED, BATCH_SEED, BATCH = 4, 11, 1


def reference_model():
    m = ref_import.load()
    cfg = m['config'].GlobalConfig(setting='eval')
    cfg.use_target_point_image = True
    cfg.n_layer = 4
    cfg.embd_pdrop = cfg.attn_pdrop = cfg.resid_pdrop = 0.0
    net = m['model'].LidarCenterNet(cfg, 'cpu', 'transFuser', 'regnety_032', 'regnety_032', use_velocity=False)
    names = [(n, tuple(p.shape)) for n, p in list(net.named_parameters()) + list(net.named_buffers())]
    net.load_state_dict(O.deterministic_state(names, seed=WEIGHT_SEED), strict=False)
    return net.train(), cfg


def main():
    torch.manual_seed(0)
    net, cfg = reference_model()
    b = O.synthetic_batch(BATCH, seed=BATCH_SEED)
    losses = net(b['rgb'], b['lidar'], ego_waypoint=b['ego_waypoint'], target_point=b['target_point'],
                 target_point_image=b['target_point_image'], ego_vel=b['ego_vel'], bev=b['bev'], label=b['label'],
                 depth=b['depth'], semantic=b['semantic'])
    w = dict(zip(cfg.detailed_losses, cfg.detailed_losses_weights))
    sum(w[k] * v for k, v in losses.items()).backward()
    out = {'loss_names': np.array(list(losses.keys())), 'losses': np.array([float(v) for v in losses.values()], dtype=np.float64)}
    sd = dict(net.named_parameters())
    out['grad_keys'] = np.array(GRAD_KEYS)
    out['grad_norms'] = np.array([float(sd[k].grad.double().norm()) for k in GRAD_KEYS])
    bn = net.state_dict()['_model.image_encoder.features.stem.bn.running_mean']
    out['stem_running_mean'] = bn.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'model_golden.npz'), **out)
    print('wrote model_golden.npz', dict(zip(out['loss_names'], out['losses'])))


if __name__ == '__main__':
    main()
