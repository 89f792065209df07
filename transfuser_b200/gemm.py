"""Host wrappers over the GEMM entry points of the C-ABI (csrc/gemm_simt.cu, csrc/gemm_tc.cu)."""
import os

import torch

from . import _lib

# 'tf32' (tcgen05 kind::tf32, default), 'simt' (exact fp32 CUDA cores: parity mode)
MODE = os.environ.get('TFB_GEMM', 'simt')


def set_mode(mode):
    global MODE
    assert mode in ('tf32', 'simt', 'bf16')
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


def bgemm(a, b, out, M, N, K, lda, ldb, ldc, trans_a, trans_b, batch_outer, batch_inner, sa, sb, sc, alpha=1.0, beta=0.0):
    """Two-level strided-batched fp32 GEMM on raw (tensor-with-offset) operands: sa/sb/sc = (outer stride, inner stride)."""
    _lib.call('tfb_gemm_f32_simt', int(trans_a), int(trans_b), M, N, K, a, lda, b, ldb, out, ldc, None, 0, float(alpha), float(beta),
              batch_outer, batch_inner, sa[0], sa[1], sb[0], sb[1], sc[0], sc[1])
    return out
