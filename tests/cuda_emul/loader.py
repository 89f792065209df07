"""TEST INFRASTRUCTURE ONLY. ctypes binding of the emulated library (build_emul.py) with the same call convention as
transfuser_b200._lib (signatures parsed from include/tfb200.h), plus the monkeypatches that let the product's own Python
(ops / backbone / model / optim) run on CPU tensors over the emulated kernels for the duration of one test."""
import ctypes

import torch

from transfuser_b200 import _lib

from . import build_emul


class EmulLib:
    def __init__(self):
        self.cdll = ctypes.CDLL(build_emul.build())
        self.fns = {}
        for name, (ret, params) in _lib.parse_header().items():
            fn = getattr(self.cdll, name, None)
            if fn is None:
                continue                                   # tensor-core entry points: not emulated
            fn.restype = ret
            fn.argtypes = [t for t, _ in params]
            self.fns[name] = (fn, params)
        self.launches = 0
        self.profiler = None
        self.log = []

    def call(self, name, *args):
        if name not in self.fns:
            raise RuntimeError('%s is a tensor-core entry point: it has no CPU emulation' % name)
        fn, params = self.fns[name]
        takes_stream = bool(params) and params[-1][1] == 'stream'
        n_user = len(params) - (1 if takes_stream else 0)
        if len(args) != n_user:
            raise TypeError('%s expects %d arguments, got %d' % (name, n_user, len(args)))
        conv = []
        for a, (t, _) in zip(args, params):
            if t is ctypes.c_void_p:
                if isinstance(a, torch.Tensor):
                    assert not a.is_cuda
                    conv.append(a.data_ptr())
                else:
                    conv.append(None if a is None else int(a))
            else:
                conv.append(a)
        if takes_stream:
            conv.append(None)
        rc = fn(*conv)
        self.launches += 1
        self.log.append(name)
        if rc != 0:
            raise RuntimeError('%s failed with code %d: %s' % (name, rc, self.cdll.tfb_last_error().decode()))


_EMUL = None


def emul():
    global _EMUL
    if _EMUL is None:
        _EMUL = EmulLib()
    return _EMUL


def patch_product(monkeypatch):
    """Routes transfuser_b200's C-ABI calls to the emulated library and replaces the two CUDA-stream dependent helpers."""
    from transfuser_b200 import gemm, ops
    lib = emul()
    lib.log.clear()
    monkeypatch.setattr(_lib, '_LIB', lib)
    ws = {}

    def _ws(device):
        t = ws.get(str(device))
        if t is None:
            t = ws[str(device)] = torch.zeros(2 * 8192 + 8, dtype=torch.float64, device=device)
        return t
    monkeypatch.setattr(ops, '_ws', _ws)
    monkeypatch.setattr(ops, 'TWO_STREAMS', False)
    monkeypatch.setitem(ops._SEED, 'dev', {})
    gemm.set_mode('simt')
    return lib
