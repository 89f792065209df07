"""Flat parameter / gradient storage + fused AdamW (csrc/gru_adamw.cu) + data-parallel gradient all-reduce.

`flatten(model)` re-homes every parameter into ONE fp32 buffer (and gives every parameter a `.grad` view into ONE
gradient buffer), which turns the optimizer step into a single HBM-bound kernel over 168 M elements and the DDP
exchange into a few large NCCL all-reduces that are issued from autograd hooks while backward is still running
(reference: torch.optim.AdamW + DistributedDataParallel, train.py:134,142,312-314)."""
import os

import torch
import torch.distributed as dist

from . import _lib
from . import ops as ops_mod

ALIGN = 64  # elements: keeps every parameter 256-byte aligned (TMA / float4 friendly)


def _qkv_packs(model):
    """[(query.p, key.p, value.p)] parameter triples (weights, then biases) of every attention module: laid out back to back in
    the flat buffer (no padding between the members) they ARE the [3C, C] weight / [3C] bias of one fused q|k|v projection
    (ops.AttentionFn: one GEMM instead of three in forward, dgrad and wgrad, one bias-gradient reduction instead of three)."""
    packs = []
    for m in model.modules():
        q, k, v = (getattr(m, n, None) for n in ('query', 'key', 'value'))
        if not all(isinstance(t, torch.nn.Linear) for t in (q, k, v)):
            continue
        for name in ('weight', 'bias'):
            trio = tuple(getattr(t, name) for t in (q, k, v))
            if all(isinstance(t, torch.nn.Parameter) for t in trio) and len({t.shape for t in trio}) == 1 \
                    and trio[0].numel() % 8 == 0 and len({id(t) for t in trio}) == 3:
                packs.append(trio)
    return packs


class FlatParams:
    def __init__(self, model):
        params, seen = [], set()
        for p in model.parameters():
            if id(p) not in seen:
                seen.add(id(p))
                params.append(p)
        # q|k|v packs: the three members take the place of the first one in parameter order, contiguous and unpadded
        pack_of = {}
        for trio in _qkv_packs(model) if isinstance(model, torch.nn.Module) else []:
            if all(id(t) in seen for t in trio) and not any(id(t) in pack_of for t in trio):
                for t in trio:
                    pack_of[id(t)] = trio
        ordered, placed = [], set()
        for p in params:
            if id(p) in placed:
                continue
            group = pack_of.get(id(p), (p,))
            for t in group:
                placed.add(id(t))
            ordered.append(group)
        dev = params[0].device
        params, offs, lens, total = [], [], [], 0
        for group in ordered:
            start = total
            for t in group:
                params.append(t)
                offs.append(total)
                lens.append(t.numel())
                total += t.numel()
            total = start + (total - start + ALIGN - 1) // ALIGN * ALIGN
            lens[-1] = total - offs[-1]                  # the padding belongs to the last member's span
        self.params, self.offsets, self.lengths, self.total = params, offs, lens, total
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        for p, o in zip(params, offs):
            n = p.numel()
            self.flat[o:o + n].copy_(p.data.reshape(-1))
            p.data = self.flat[o:o + n].view(p.shape)
            p.grad = None
            p._tfb_flat = (self, o)
        self.bf16 = None
        self.taken = set()        # offsets whose gradient span was handed to a backward kernel in the current step (ops._gbuf)

    def set_grads_to_none(self):
        """zero_grad(set_to_none=True): with .grad None the backward kernels write each gradient straight into its span of
        the flat buffer (ops._gbuf) and autograd adopts that view as p.grad — no accumulate / copy kernels."""
        for p in self.params:
            p.grad = None
        self.taken.clear()

    def gather_stragglers(self, reduced=False):
        """Before the optimizer reads the flat gradient buffer: a parameter whose .grad is not its flat view (accumulated by
        autograd into a fresh tensor) is copied in; a parameter that received no gradient is zeroed. Returns the (offset,
        padded length) spans of the parameters without a gradient: torch.optim.AdamW leaves those untouched (no weight decay,
        no moment update), so the fused optimizer skips them too."""
        base = self.grad.data_ptr()
        no_grad = []
        for p, o, ln in zip(self.params, self.offsets, self.lengths):
            g = p.grad
            if g is None:
                self.grad[o:o + p.numel()].zero_()
                no_grad.append((o, ln))
            elif g.data_ptr() != base + 4 * o:
                if reduced:
                    # the data-parallel all-reduce of this span was launched from the backward hooks and has already read (or is
                    # reading) the flat buffer: copying now would update this parameter with the local, unreduced gradient
                    raise RuntimeError('gradient of a %s parameter was not written into the flat buffer (ops._gbuf) before the '
                                       'gradient all-reduce' % (tuple(p.shape),))
                self.grad[o:o + p.numel()].copy_(g.reshape(-1))
        return no_grad


def flatten(model):
    fp = FlatParams(model)
    model._tfb_flat_params = fp
    return fp


def subtract_spans(lo, hi, holes):
    """[lo, hi) minus the (offset, length) `holes` (sorted, disjoint) -> list of (lo, hi) pieces."""
    out = []
    for o, n in holes:
        if o + n <= lo or o >= hi:
            continue
        if o > lo:
            out.append((lo, o))
        lo = max(lo, o + n)
    if lo < hi:
        out.append((lo, hi))
    return out


class FusedAdamW(torch.optim.Optimizer):
    """torch.optim.AdamW semantics (decoupled weight decay, bias correction), one fused kernel per flat buffer.
    Gradients are written in place into the flat buffer by the backward kernels (ops._gbuf), so zero_grad() only drops the
    .grad references and no buffer is cleared."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, grad_scale=1.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.grad_scale = grad_scale
        self._step = 0
        self._flat = None
        self._m = self._v = None
        self._step_dev = None
        self._hp_dev = None       # [lr, beta1, beta2, eps, weight_decay] in device memory: read by the kernel, so a schedule that edits
        self._hp_host = None      # param_groups keeps working under CUDA-graph replay (sync_hparams() before every launch / replay)
        self._opt_stream = None   # pipelined mode (step_span)
        self._spans_done = set()
        self._began = False
        self.check_grads = True   # set False once the producer set is known to be complete (saves a Python sweep per step)
        self._no_grad_spans = []  # parameters that never receive a gradient (as found by the last checked step)

    def _flat_of(self):
        if self._flat is None:
            ps = [p for g in self.param_groups for p in g['params']]
            owners = {id(getattr(p, '_tfb_flat', (None,))[0]) for p in ps}
            if len(owners) != 1 or not hasattr(ps[0], '_tfb_flat'):
                raise RuntimeError('FusedAdamW needs parameters flattened by transfuser_b200.optim.flatten(model)')
            self._flat = ps[0]._tfb_flat[0]
            self._m = torch.zeros_like(self._flat.flat)
            self._v = torch.zeros_like(self._flat.flat)
            self._step_dev = torch.zeros(1, dtype=torch.int32, device=self._flat.flat.device)
            self._hp_dev = torch.zeros(5, dtype=torch.float64, device=self._flat.flat.device)
        return self._flat

    def sync_hparams(self):
        """Uploads lr / betas / eps / weight_decay of param_groups[0] to the device scalars the kernel reads, when they changed.
        (A plain copy_ from a host tensor: legal outside graph capture only — Trainer.replay() calls this before each replay.)"""
        self._flat_of()
        g = self.param_groups[0]
        hp = (float(g['lr']), float(g['betas'][0]), float(g['betas'][1]), float(g['eps']), float(g['weight_decay']))
        if hp != self._hp_host:
            if self._hp_dev.is_cuda and torch.cuda.is_current_stream_capturing():
                raise RuntimeError('optimizer hyper-parameters changed during CUDA-graph capture')
            self._hp_dev.copy_(torch.tensor(hp, dtype=torch.float64))
            self._hp_host = hp

    def step_count(self):
        """Optimizer steps taken so far: the device-side counter (the only one that advances under CUDA-graph replay)."""
        return int(self._step_dev.item()) if self._step_dev is not None else self._step

    def zero_grad(self, set_to_none=True):
        self._flat_of().set_grads_to_none()

    # ---- pipelined mode: the update of a span of the flat buffer is launched from the backward pass as soon as the span's gradients
    # are complete (and, data-parallel, all-reduced), on a stream of its own, so the 0.9 ms HBM-bound AdamW pass of the 168 M
    # parameters overlaps the rest of backward instead of trailing it. Valid because a parameter's gradient is complete only after
    # every backward kernel that READS the parameter (its dgrad) has been enqueued, and the update stream waits for all of them.
    def begin_step(self):
        """Before backward: advances the device-side step count, uploads changed hyper-parameters, forgets the spans of the last step."""
        self._flat_of()
        self.sync_hparams()
        self._step += 1
        _lib.call('tfb_step_tick', None, self._step_dev)
        self._began = True
        self._spans_done = set()

    def _launch(self, lo, hi):
        fp, g = self._flat, self.param_groups[0]
        b1, b2 = g['betas']
        for a, b in (subtract_spans(lo, hi, self._no_grad_spans) if self._no_grad_spans else [(lo, hi)]):
            bf = fp.bf16[a:b] if fp.bf16 is not None else None
            _lib.call('tfb_adamw_step', fp.flat[a:b], fp.grad[a:b], self._m[a:b], self._v[a:b], b - a, float(g['lr']), float(b1),
                      float(b2), float(g['eps']), float(g['weight_decay']), self._step, self._step_dev, float(self.grad_scale), bf, 0, self._hp_dev)

    def step_span(self, lo, hi, work):
        """Called from a gradient hook (GradAllReducer) when span [lo, hi) is complete: update it on the optimizer stream."""
        fp = self._flat_of()
        dev = fp.flat.device
        if self._opt_stream is None:
            self._opt_stream = torch.cuda.Stream(device=dev)
        cur = torch.cuda.current_stream(dev)
        self._opt_stream.wait_stream(cur)
        for s in ops_mod.side_streams(dev):          # LiDAR trunk / decoders / weight-gradient streams produce gradients too
            self._opt_stream.wait_stream(s)
        with torch.cuda.stream(self._opt_stream):
            if work is not None:
                work.wait()
            self._launch(lo, hi)
        self._spans_done.add((lo, hi))

    @torch.no_grad()
    def step(self, closure=None, chunks=None):
        fp = self._flat_of()
        ops_mod.join_side_streams()       # weight gradients are produced on side streams (ops._OnWgradStream, the LiDAR branch)
        if self.check_grads:
            self._no_grad_spans = fp.gather_stragglers(reduced=chunks is not None and any(w is not None for _, _, w in chunks))
        if not getattr(self, '_began', False):
            self.sync_hparams()
            self._step += 1
            _lib.call('tfb_step_tick', None, self._step_dev)   # device-side step count: valid under CUDA-graph replay
            self._spans_done = set()
        self._began = False
        ops_mod.invalidate_packs()                         # the kernels rewrite the weights without bumping tensor versions
        spans = chunks or [(0, fp.total, None)]
        for lo, hi, work in spans:
            if (lo, hi) in self._spans_done:
                continue                                   # already updated from the backward pass (step_span)
            if work is not None:
                work.wait()
            self._launch(lo, hi)
        if self._opt_stream is not None:
            torch.cuda.current_stream(fp.flat.device).wait_stream(self._opt_stream)
        self._spans_done = set()


    # ---- torch.optim.AdamW-compatible (de)serialisation: the reference saves / resumes optimizer_%d.pth (train.py:183, 384)
    def state_dict(self):
        """Same structure as torch.optim.AdamW.state_dict(): per-parameter {'step', 'exp_avg', 'exp_avg_sq'} in parameter
        order, so a reference run can resume from it (and vice versa). Every parameter shares the fused step count."""
        fp = self._flat_of()
        ps = [p for g in self.param_groups for p in g['params']]
        off = {id(p): o for p, o in zip(fp.params, fp.offsets)}
        state = {}
        step = self.step_count()
        if step > 0:
            for i, p in enumerate(ps):
                o, n = off[id(p)], p.numel()
                state[i] = dict(step=torch.tensor(float(step)),
                                exp_avg=self._m[o:o + n].view(p.shape).clone(), exp_avg_sq=self._v[o:o + n].view(p.shape).clone())
        groups, start = [], 0
        for g in self.param_groups:
            d = {k: v for k, v in g.items() if k != 'params'}
            d.setdefault('amsgrad', False)
            d['params'] = list(range(start, start + len(g['params'])))
            start += len(g['params'])
            groups.append(d)
        return dict(state=state, param_groups=groups)

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        fp = self._flat_of()
        ps = [p for g in self.param_groups for p in g['params']]
        saved = [i for g in state_dict['param_groups'] for i in g['params']]
        if len(saved) != len(ps):
            raise ValueError('optimizer state has %d parameters, the model has %d' % (len(saved), len(ps)))
        off = {id(p): o for p, o in zip(fp.params, fp.offsets)}
        self._m.zero_()
        self._v.zero_()
        steps = set()
        for idx, p in zip(saved, ps):
            st = state_dict['state'].get(idx)
            if st is None:
                continue                       # never received a gradient in the saved run: moments stay zero
            if tuple(st['exp_avg'].shape) != tuple(p.shape):
                raise ValueError('optimizer state %d has shape %s, parameter has %s' % (idx, tuple(st['exp_avg'].shape), tuple(p.shape)))
            o, n = off[id(p)], p.numel()
            self._m[o:o + n].copy_(st['exp_avg'].reshape(-1))
            self._v[o:o + n].copy_(st['exp_avg_sq'].reshape(-1))
            steps.add(int(st['step']))
        if len(steps) > 1:
            raise ValueError('per-parameter step counts differ (%s): the fused optimizer keeps one step count' % sorted(steps))
        self._step = steps.pop() if steps else 0
        self._step_dev.fill_(self._step)
        for g, sg in zip(self.param_groups, state_dict['param_groups']):
            for k in ('lr', 'betas', 'eps', 'weight_decay'):
                if k in sg:
                    g[k] = tuple(sg[k]) if k == 'betas' else sg[k]
            if sg.get('amsgrad', False):
                raise ValueError('amsgrad=True is not implemented (train.py:142 uses the default)')


_SKIP_ALLREDUCE = os.environ.get('TFB_NO_ALLREDUCE', '0') == '1'
SPLIT_LATE_SPANS = os.environ.get('TFB_SPLIT_LATE_SPANS', '1') == '1'     # readiness-aware exchange spans (GradAllReducer.replan)


class GradAllReducer:
    """Data-parallel exchange: SUM all-reduce of the flat gradient buffer in `n_chunks` spans. Each span is launched (async,
    NCCL stream) from a post-accumulate-grad hook as soon as the last parameter of that span has its gradient, so the
    exchange overlaps the rest of backward; FusedAdamW waits per span and folds the 1/world_size into its update."""

    def __init__(self, fp, n_chunks=8, opt=None):
        self.fp = fp
        self.opt = opt            # FusedAdamW to update each span as soon as it is complete (pipelined mode), or None
        self.pipeline = False     # switched on by Trainer once the set of gradient producers is known to be complete
        self._launch_stream = None
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.n_chunks = n_chunks
        self._fired = []          # parameter indices in the order their gradients completed (first backward pass: see replan())
        self._record = SPLIT_LATE_SPANS
        self.replanned = False
        per = (fp.total + n_chunks - 1) // n_chunks
        per = (per + ALIGN - 1) // ALIGN * ALIGN
        self._set_spans([(lo, min(fp.total, lo + per)) for lo in range(0, fp.total, per)])
        if self.world > 1 or opt is not None:
            for i, p in enumerate(fp.params):
                p.register_post_accumulate_grad_hook(self._make_hook(i))

    def _set_spans(self, spans):
        fp = self.fp
        self.spans = spans
        starts = [lo for lo, _ in spans]
        self.counts = [0] * len(spans)
        self.param_chunks = []
        import bisect
        for p, o in zip(fp.params, fp.offsets):
            first = bisect.bisect_right(starts, o) - 1
            last = bisect.bisect_right(starts, o + max(p.numel(), 1) - 1) - 1
            ids = list(range(first, last + 1))
            self.param_chunks.append(ids)
            for c in ids:
                self.counts[c] += 1
        self.reset()

    def replan(self, late_frac=0.2):
        """Readiness-aware spans. Uniform spans in flat-buffer (registration) order complete when their LAST gradient does — and the
        flat order starts each encoder with its stem, whose gradient is the last of the whole backward pass: the spans holding
        stem / stage 1 / stage 2 (a few MB of parameters) drag tens of MB of long-finished stage-3 gradients into an exchange that can
        only start when backward ends, fully exposed. After the first backward pass the order in which gradients completed is known;
        every uniform span is cut wherever it changes between `late` parameters (the last `late_frac` of that order) and the rest, so
        the exposed tail of the step carries only the late parameters' bytes. The plan comes from rank 0 (all ranks must issue the same
        collectives). Call between steps (Trainer does, once, before capturing the step)."""
        n = len(self._fired)
        fp = self.fp
        if self.replanned or n == 0:
            return False
        rank_of = {i: r for r, i in enumerate(self._fired)}
        late = [rank_of.get(i, -1) >= (1.0 - late_frac) * n for i in range(len(fp.params))]
        cuts = set()
        prev = None
        for i, o in enumerate(fp.offsets):
            if prev is not None and late[i] != prev:
                cuts.add(o)
            prev = late[i]
        per = (fp.total + self.n_chunks - 1) // self.n_chunks
        per = (per + ALIGN - 1) // ALIGN * ALIGN
        bounds = sorted(set(range(0, fp.total, per)) | cuts | {fp.total})
        if self.world > 1:                                  # one plan for every rank
            t = torch.zeros(256, dtype=torch.int64, device=fp.grad.device)
            if dist.get_rank() == 0:
                assert len(bounds) <= 255
                t[0] = len(bounds)
                t[1:1 + len(bounds)] = torch.tensor(bounds, dtype=torch.int64)
            dist.broadcast(t, 0)
            t = t.cpu()
            bounds = [int(v) for v in t[1:1 + int(t[0])]]
        self._set_spans([(a, b) for a, b in zip(bounds[:-1], bounds[1:]) if b > a])
        self._record, self.replanned = False, True
        return True

    def _all_reduce(self, lo, hi):
        """Async SUM all-reduce of grad[lo:hi]. The gradients of the span were produced on several streams (critical chain, LiDAR
        trunk / decoders, weight-gradient stream): the collective is issued from a launch stream that waits for all of them, so the
        critical chain itself is never made to wait for the side streams here (NCCL orders its kernel after the issuing stream)."""
        g = self.fp.grad[lo:hi]
        if _SKIP_ALLREDUCE:
            return None          # measurement only (TFB_NO_ALLREDUCE=1): the step WITHOUT the exchange, to attribute its cost; results are wrong
        if not g.is_cuda:
            return dist.all_reduce(g, op=dist.ReduceOp.SUM, async_op=True)
        dev = g.device
        if self._launch_stream is None:
            self._launch_stream = torch.cuda.Stream(device=dev)
        ls = self._launch_stream
        ls.wait_stream(torch.cuda.current_stream(dev))
        for s_ in ops_mod.side_streams(dev):
            ls.wait_stream(s_)
        with torch.cuda.stream(ls):
            return dist.all_reduce(g, op=dist.ReduceOp.SUM, async_op=True)

    def reset(self):
        self.pending = list(self.counts)
        self.works = [None] * len(self.spans)

    def _make_hook(self, i):
        def hook(_p):
            if self._record:
                self._fired.append(i)
            for c in self.param_chunks[i]:
                self.pending[c] -= 1
                if self.pending[c] == 0:
                    lo, hi = self.spans[c]
                    if self.world > 1:
                        self.works[c] = self._all_reduce(lo, hi)
                    if self.pipeline and self.opt is not None and self.fp.grad.is_cuda:
                        self.opt.step_span(lo, hi, self.works[c])
        return hook

    def chunks(self):
        """(lo, hi, work) spans for FusedAdamW.step; spans whose hook never fired (unused params) are reduced here."""
        out = []
        for c, (lo, hi) in enumerate(self.spans):
            w = self.works[c]
            if w is None and self.world > 1:
                w = self._all_reduce(lo, hi)
            out.append((lo, hi, w))
        self.reset()
        return out
