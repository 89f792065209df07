"""Bisect tool: loss of the 3rd training step, eager (3 eager steps) vs captured graph (2 eager warm-ups + 1 replay), per feature flag."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from bench import make_host_batch
from transfuser_b200 import ops, trainer as T
from transfuser_b200.config import TrainConfig
from transfuser_b200.trainer import Trainer

dev = torch.device('cuda', 0)
host = make_host_batch(2, seed=5, torch=torch, np=np)


def run(use_graph, cfg):
    torch.manual_seed(0)
    tr = Trainer(cfg, dev, gemm_mode='bf16', seed=0)
    if use_graph:
        assert tr.capture(host), tr.graph_error
        loss = tr.replay()
        torch.cuda.synchronize()
        l3 = loss.item()
        d3 = {k: round(v.item(), 5) for k, v in tr.last_losses.items()}
        l4 = tr.replay().item()
        d4 = {k: round(v.item(), 5) for k, v in tr.last_losses.items()}
    else:
        d = {k: v.to(dev) for k, v in host.items()}
        for _ in range(3):
            loss = tr.step(d)
        l3 = loss.item()
        d3 = {k: round(v.item(), 5) for k, v in tr.last_losses.items()}
        l4 = tr.step(d).item()
        d4 = {k: round(v.item(), 5) for k, v in tr.last_losses.items()}
    torch.cuda.synchronize()
    print('   ', 'graph' if use_graph else 'eager', 'step3', d3)
    print('   ', 'graph' if use_graph else 'eager', 'step4', d4)
    return l3, l4


cases = [('default', {}), ('no dropout', {'cfg': dict(embd_pdrop=0.0, attn_pdrop=0.0, resid_pdrop=0.0)}), ('OVERLAP_ADAMW=0', {'T.OVERLAP_ADAMW': False}),
         ('WGRAD_STREAM=0', {'ops.WGRAD_STREAM': False}), ('TWO_STREAMS=0', {'ops.TWO_STREAMS': False}), ('BN_STATS_FUSED=0', {'ops.BN_STATS_FUSED': False}),
         ('DECODER_STREAMS=0', {'ops.DECODER_STREAMS': False}), ('ATTN_FUSED=0', {'ops.ATTN_FUSED': False}), ('WGRAD_MAX_CTAS=0', {'ops.WGRAD_MAX_CTAS': 0})]
only = sys.argv[1:]
for name, flags in cases:
    if only and not any(o in name for o in only):
        continue
    old = {}
    cfg = TrainConfig(**flags.get('cfg', {}))
    for k, v in flags.items():
        if k == 'cfg':
            continue
        mod, attr = k.split('.')
        m = {'ops': ops, 'T': T}[mod]
        old[k] = getattr(m, attr)
        setattr(m, attr, v)
    try:
        e3, e4 = run(False, cfg)
        g3, g4 = run(True, cfg)
        print('%-20s eager step3 %.5f step4 %.5f | graph step3 %.5f step4 %.5f | rel diff %.2e %.2e' % (name, e3, e4, g3, g4, abs(e3 - g3) / abs(e3), abs(e4 - g4) / abs(e4)), flush=True)
    finally:
        for k, v in old.items():
            mod, attr = k.split('.')
            setattr({'ops': ops, 'T': T}[mod], attr, v)
