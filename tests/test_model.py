"""Whole hot path: LidarCenterNet.forward + backward on the GPU (CUDA kernels through the C-ABI) vs the CPU oracle
(oracle/torch_oracle.py, itself pinned to the verbatim reference) on identical weights and inputs."""
import sys
import os

import pytest
import torch

from oracle import torch_oracle as O

pytestmark = pytest.mark.gpu


class Cfg:
    """GlobalConfig fields the product reads (config.py), train.py defaults applied."""
    lidar_seq_len = 1; seq_len = 1; use_point_pillars = False; use_target_point_image = True
    gru_concat_target_point = True; pred_len = 4; lidar_pos = [1.3, 0.0, 2.5]
    n_head = 4; block_exp = 4; n_layer = 4
    embd_pdrop = attn_pdrop = resid_pdrop = 0.0
    gpt_linear_layer_init_mean = 0.0; gpt_linear_layer_init_std = 0.02; gpt_layer_norm_init_weight = 1.0
    img_vert_anchors = 5; img_horz_anchors = 22; lidar_vert_anchors = 8; lidar_horz_anchors = 8
    perception_output_features = 512; bev_features_chanels = 64; bev_upsample_factor = 2
    deconv_channel_num_1 = 128; deconv_channel_num_2 = 64; deconv_channel_num_3 = 32
    deconv_scale_factor_1 = 8; deconv_scale_factor_2 = 4
    channel = 64; num_class = 7; num_dir_bins = 12; gru_hidden_size = 64; multitask = True
    ls_seg = 1.0; ls_depth = 10.0
    bev_resolution_width = bev_resolution_height = 160
    lidar_resolution_width = lidar_resolution_height = 256
    detailed_losses = ['loss_wp', 'loss_bev', 'loss_depth', 'loss_semantic', 'loss_center_heatmap', 'loss_wh', 'loss_offset',
                       'loss_yaw_class', 'loss_yaw_res', 'loss_velocity', 'loss_brake']
    detailed_losses_weights = [1.0, 1.0, 1.0, 1.0, 0.2, 0.2, 0.2, 0.2, 0.2, 0.0, 0.0]


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def build(seed=1):
    from transfuser_b200 import LidarCenterNet
    net = LidarCenterNet(Cfg, 'cpu', 'transFuser', 'regnety_032', 'regnety_032', use_velocity=False)
    names = [(n, tuple(p.shape)) for n, p in list(net.named_parameters()) + list(net.named_buffers()) if not n.startswith('_bev')]
    net.load_state_dict(O.deterministic_state(names, seed=seed), strict=False)
    return net


def test_full_model_forward_backward_matches_oracle():
    torch.manual_seed(0)
    net = build()
    P = {k: v.clone().requires_grad_(v.dtype.is_floating_point and 'running' not in k) for k, v in net.state_dict().items()}
    batch = O.synthetic_batch(2, seed=3)

    class C(O.Cfg):
        embd_pdrop = attn_pdrop = resid_pdrop = 0.0
    taps = {}
    ref = O.forward(P, batch, C, train=True, taps=taps)
    w = dict(zip(Cfg.detailed_losses, Cfg.detailed_losses_weights))
    sum(w[k] * ref[k] for k in ref).backward()

    net = net.cuda().train()
    cb = {k: v.cuda() for k, v in batch.items()}
    out = net(cb['rgb'], cb['lidar'], ego_waypoint=cb['ego_waypoint'], target_point=cb['target_point'],
              target_point_image=cb['target_point_image'], ego_vel=cb['ego_vel'], bev=cb['bev'], label=cb['label'],
              depth=cb['depth'], semantic=cb['semantic'])
    assert list(out.keys()) == list(ref.keys())
    report = []
    for k in ref:
        e = abs(out[k].item() - ref[k].item()) / max(abs(ref[k].item()), 1e-12)
        report.append('%-22s oracle %.7f cuda %.7f rel %.2e' % (k, ref[k].item(), out[k].item(), e))
    print('\n'.join(report))
    for k in ref:
        assert abs(out[k].item() - ref[k].item()) <= 1e-3 * max(abs(ref[k].item()), 1e-6), k
    loss = sum(w[k] * out[k] for k in out)
    loss.backward()
    worst = []
    for n, p in net.named_parameters():
        g = P[n].grad
        if n.endswith('attn.key.bias'):
            continue  # true gradient is identically zero (softmax shift invariance): both sides are rounding noise
        worst.append((rel(p.grad, g), n))
    worst.sort(reverse=True)
    print('worst parameter-gradient errors:', worst[:6])
    assert worst[0][0] < 1e-3, worst[:6]
    # BatchNorm running statistics were updated identically
    sd = net.state_dict()
    bad = [(rel(sd[k], P[k]), k) for k in P if 'running' in k and ('.stem.' in k or '.s1.' in k or '.s4.' in k)]
    assert max(bad)[0] < 1e-4, max(bad)
    assert int(sd['_model.image_encoder.features.stem.bn.num_batches_tracked']) == 1
