"""One data-parallel training step of the hot path — BEV histogram, LidarCenterNet forward, weighted loss sum, backward,
gradient all-reduce, fused AdamW — as a reusable object, optionally captured into ONE CUDA graph and replayed.

This is the body of the reference's `Engine.train()` loop (train.py:304-316) with the synchronising `.item()` calls moved
out of the step; bench.py and the tests drive it."""
import torch
import torch.distributed as dist

from . import _lib, bev, gemm, ops, optim
from .model import LidarCenterNet

import os

OVERLAP_ADAMW = os.environ.get('TFB_OVERLAP_ADAMW', '1') == '1'   # AdamW per gradient span from inside backward (captured step only)

INPUT_KEYS = ('rgb', 'points', 'target_point_image', 'target_point', 'ego_vel', 'ego_waypoint', 'bev', 'semantic', 'depth', 'label')
# raw_inputs=True: what is on disk instead of the expanded tensors (uint8 frames, raw points, one pose transform, the target point);
# the GPU input pipeline (pipeline.InputPipeline, csrc/input_prep.cu) builds the model inputs inside the step — 1.6 MB instead of
# 3.8 MB per sample over PCIe
RAW_KEYS = ('rgb_u8', 'depth_u8', 'seg_u8', 'crop_shift', 'points', 'transforms', 'target_point64', 'ego_vel', 'ego_waypoint', 'bev', 'label')


class Trainer:
    def __init__(self, cfg, device, gemm_mode='bf16', lr=1e-4, seed=0, n_chunks=int(os.environ.get('TFB_GRAD_CHUNKS', '8')), backbone='transFuser',
                 raw_inputs=False):
        self.cfg, self.device = cfg, device
        # the geometric-fusion backbone consumes two more inputs per sample (train.py:279-288)
        self.raw_inputs = raw_inputs
        self.input_keys = (RAW_KEYS if raw_inputs else INPUT_KEYS) + (('bev_points', 'cam_points') if backbone == 'geometric_fusion' else ())
        if raw_inputs:
            from .pipeline import InputPipeline
            self.pipe = InputPipeline(cfg, device)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        gemm.set_mode(gemm_mode)
        ops.manual_seed(1234 + seed)
        self.net = LidarCenterNet(cfg, device, backbone, 'regnety_032', 'regnety_032', use_velocity=False).train()
        self.flat = optim.flatten(self.net)
        if self.world > 1:
            # replicas start identical whatever the caller's RNG state: rank 0's parameters and buffers win (DDP does the same at
            # construction, train.py:134)
            dist.broadcast(self.flat.flat, 0)
            for buf in self.net.buffers():
                if buf.is_floating_point() or buf.dtype == torch.int64:
                    dist.broadcast(buf, 0)
        if gemm_mode == 'bf16':
            gemm.attach_bf16_weights(self.flat)
        self.opt = optim.FusedAdamW(self.net.parameters(), lr=lr, grad_scale=1.0 / self.world)
        self.reducer = optim.GradAllReducer(self.flat, n_chunks=n_chunks, opt=self.opt if OVERLAP_ADAMW else None)
        self.weights = dict(zip(cfg.detailed_losses, cfg.detailed_losses_weights))
        self.graph = None
        self.static = None
        self.static_loss = None
        self.graph_launches = 0
        self.graph_error = None
        self._stream = None

    def step(self, d):
        """Eager step on device-resident inputs `d` (dict with INPUT_KEYS). Returns the weighted total loss (0-dim tensor).

        Never runs on the legacy default stream: autograd pins every parameter's AccumulateGrad node to the stream of its first
        backward, the gradient hooks keep those nodes alive, and a later CUDA-graph capture cannot fork the legacy stream
        (cudaErrorStreamCaptureImplicit — the round-1 "AccumulateGrad node's stream does not match" warning, fatal once the hooks
        also drive the pipelined AdamW). When called on the default stream the step runs on a stream of the trainer's own, ordered
        after the caller's work, and the caller's stream is ordered after the step."""
        dev = self.device
        if torch.device(dev).type == 'cuda' and not torch.cuda.is_current_stream_capturing():
            cur = torch.cuda.current_stream(dev)
            if cur == torch.cuda.default_stream(dev):
                if self._stream is None:
                    self._stream = torch.cuda.Stream(device=dev)
                own = self._stream
                own.wait_stream(cur)
                for t in d.values():
                    if isinstance(t, torch.Tensor) and t.is_cuda:
                        t.record_stream(own)
                with torch.cuda.stream(own):
                    loss = self._step(d)
                cur.wait_stream(own)
                loss.record_stream(cur)
                return loss
        return self._step(d)

    def _step(self, d):
        if self.raw_inputs:
            # crop + CHW / normalise, depth decode, class LUT, align + BEV histogram, target-point map: three launches on the raw batch
            p = self.pipe.prepare(dict(rgb=d['rgb_u8'], depth=d['depth_u8'], seg=d['seg_u8'], crop_shift=d['crop_shift'], points=d['points'],
                                       transforms=d['transforms'], target_point=d['target_point64']), normalized_nhwc=True)
            d = dict(d, rgb=p['rgb'], depth=p['depth'], semantic=p['semantic'], target_point=p['target_point'],
                     target_point_image=p['target_point_image'])
            lidar = p['lidar']
        else:
            lidar = bev.lidar_to_histogram_features_batched(d['points'])
        self.opt.zero_grad()
        if self.reducer.pipeline:
            self.opt.begin_step()        # AdamW spans are launched from the backward pass (optim.FusedAdamW.step_span)
        losses = self.net(d['rgb'], lidar, ego_waypoint=d['ego_waypoint'], target_point=d['target_point'],
                          target_point_image=d['target_point_image'], ego_vel=d['ego_vel'], bev=d['bev'], label=d['label'],
                          depth=d['depth'], semantic=d['semantic'], bev_points=d.get('bev_points'), cam_points=d.get('cam_points'))
        self.last_losses = losses          # the 11 scalars of this step (device tensors; static buffers under graph replay)
        loss = None
        for k, v in losses.items():
            loss = v * self.weights[k] if loss is None else loss + v * self.weights[k]
        loss.backward()
        self.opt.step(chunks=self.reducer.chunks())
        if not self.reducer.replanned and self.graph is None and not torch.cuda.is_current_stream_capturing():
            self.reducer.replan()       # once, after the first backward pass showed the order in which gradients complete
        return loss

    def capture(self, example):
        """Captures step() into a CUDA graph over static input buffers shaped like `example` (host or device tensors).
        The dropout base seed and the optimizer step count live in device memory and are advanced by kernels inside the graph,
        so every replay is a genuinely new step. Returns True on success (falls back to eager otherwise)."""
        try:
            self.static = {k: torch.empty(example[k].shape, dtype=example[k].dtype, device=self.device) for k in self.input_keys}
            self.load(example)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    self.step(self.static)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            self.opt.check_grads = False
            self.reducer.pipeline = OVERLAP_ADAMW    # every gradient is known to land in the flat buffer: update spans as they complete
            g = torch.cuda.CUDAGraph()
            l0 = _lib.lib().launches
            with torch.cuda.graph(g):
                self.static_loss = self.step(self.static)
            self.graph, self.graph_launches = g, _lib.lib().launches - l0
            return True
        except Exception as e:  # noqa: BLE001 — keep the eager path, report why
            import traceback
            self.graph, self.graph_error = None, repr(e)[:300]
            self.graph_traceback = traceback.format_exc()
            if os.environ.get('TFB_CAPTURE_DEBUG') == '1':
                print(self.graph_traceback, flush=True)
            torch.cuda.synchronize()
            return False

    def load(self, batch):
        """Copies a batch (pinned host or device tensors) into the static input buffers of the captured graph."""
        for k in self.input_keys:
            self.static[k].copy_(batch[k], non_blocking=True)

    def replay(self):
        self.opt.sync_hparams()       # lr / weight-decay edits of param_groups reach the captured AdamW kernel through device scalars
        self.graph.replay()
        ops.invalidate_packs()        # the replayed optimizer kernel changed the weights behind Python's back
        return self.static_loss
