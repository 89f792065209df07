"""Host wrappers over the GEMM entry points of the C-ABI (csrc/gemm_simt.cu, csrc/gemm_tc.cu)."""
import os

import torch

from . import _lib

# 'tf32' (tcgen05 kind::tf32, default), 'simt' (exact fp32 CUDA cores: parity mode)
MODE = os.environ.get('TFB_GEMM', 'tf32')


def set_mode(mode):
    global MODE
    assert mode in ('tf32', 'simt')
    MODE = mode


def _ld(t):
    assert t.dim() == 2 and t.stride(1) == 1, 'row-major 2-D view required'
    return t.stride(0)


def gemm(a, b, out, trans_a=False, trans_b=False, bias=None, relu=False, alpha=1.0, beta=0.0, splits=1, mode=None):
    """out[M,N] = alpha * op(a) @ op(b) + beta*out (+bias) (relu). a, b, out: fp32 CUDA 2-D views with unit inner stride."""
    mode = mode or MODE
    M, N = out.shape
    K = a.shape[0] if trans_a else a.shape[1]
    lda, ldb, ldc = _ld(a), _ld(b), _ld(out)
    use_tc = (mode != 'simt' and lda % 4 == 0 and ldb % 4 == 0 and a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0
              and M >= 1 and N >= 8 and K >= 8)
    if use_tc:
        _lib.call('tfb_gemm_tf32_tc', int(trans_a), int(trans_b), M, N, K, a, lda, b, ldb, out, ldc, bias, int(relu),
                  float(alpha), float(beta), int(splits))
    else:
        _lib.call('tfb_gemm_f32_simt', int(trans_a), int(trans_b), M, N, K, a, lda, b, ldb, out, ldc, bias, int(relu),
                  float(alpha), float(beta), 1, 1, 0, 0, 0, 0, 0, 0)
    return out
