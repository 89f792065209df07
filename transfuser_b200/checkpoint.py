"""Checkpoint compatibility with the reference's on-disk format.

The reference writes `model_%d.pth` = `model.state_dict()` of the (possibly DDP-wrapped, so `module.`-prefixed) module and
`optimizer_%d.pth` = `torch.optim.AdamW.state_dict()` (train.py:381-384), reads them back with `load_state_dict`
(train.py:179-183), and the driving agent strips the 7-character `module.` prefix unconditionally and loads with
strict=False (submission_agent.py:93-96). SyncBatchNorm conversion (submission_agent.py:92) does not change key names or
shapes, so the same files load here. The product modules already expose the reference's key set, including the alias
keys (`s1`/`layer1`, `stem.bn`/`bn1`, ...), so this file is only about prefixes, partial key sets and keeping the flat /
bf16 mirrors of transfuser_b200.optim coherent after a load."""
import torch


def clean_state_dict(state_dict):
    """Drops a DistributedDataParallel `module.` prefix when every key carries it (train.py:381-383 keeps it on disk)."""
    keys = list(state_dict.keys())
    if keys and all(k.startswith('module.') for k in keys):
        return {k[len('module.'):]: v for k, v in state_dict.items()}
    return dict(state_dict)


def _alias_groups(model):
    """Groups of state_dict keys that name the same tensor (the reference registers its trunk stages twice)."""
    by_ptr = {}
    for k, v in model.state_dict(keep_vars=True).items():
        by_ptr.setdefault((v.data_ptr(), tuple(v.shape), v.dtype), []).append(k)
    return [g for g in by_ptr.values() if len(g) > 1]


def load_model(model, src, strict=True, map_location=None):
    """Loads a reference checkpoint (path or state dict) into a transfuser_b200 module. A checkpoint that carries only one
    name of an aliased tensor (e.g. written from `model.module` after de-duplication) is completed from its aliases.
    Returns the (missing, unexpected) key lists of `load_state_dict`."""
    if isinstance(src, (str, bytes)) or hasattr(src, 'read'):
        src = torch.load(src, map_location=map_location or next(model.parameters()).device)
    sd = clean_state_dict(src)
    for group in _alias_groups(model):
        have = [k for k in group if k in sd]
        if have:
            for k in group:
                sd.setdefault(k, sd[have[0]])
    result = model.load_state_dict(sd, strict=strict)
    refresh_mirrors(model)
    return list(result.missing_keys), list(result.unexpected_keys)


def refresh_mirrors(model):
    """After parameters changed outside the fused optimizer: re-derive the bf16 weight mirror (tensor-core mode) and drop the
    packed conv weights of the current step."""
    from . import ops
    ops.invalidate_packs()
    fp = getattr(model, '_tfb_flat_params', None)
    if fp is not None and fp.bf16 is not None:
        from . import gemm
        gemm.attach_bf16_weights(fp)


def save_model(model, path, module_prefix=False):
    """Writes what train.py:383 writes; `module_prefix=True` reproduces a DDP-wrapped save."""
    sd = {('module.' + k if module_prefix else k): v.detach().clone().contiguous() for k, v in model.state_dict().items()}
    torch.save(sd, path)


def load_optimizer(optimizer, src, map_location=None):
    """optimizer: transfuser_b200.optim.FusedAdamW; src: path or `torch.optim.AdamW.state_dict()` (train.py:183, 384)."""
    if isinstance(src, (str, bytes)) or hasattr(src, 'read'):
        src = torch.load(src, map_location=map_location or 'cpu')
    optimizer.load_state_dict(src)


def save_optimizer(optimizer, path):
    torch.save(optimizer.state_dict(), path)
