"""The per-op parity tests of tests/test_ops.py (forward and backward of every exact-fp32 kernel through the product's autograd
wrappers, against plain fp32 torch ops) executed in this container WITHOUT a GPU: the same test functions, with the C-ABI calls
routed to the CPU emulation of the unchanged kernel sources (tests/cuda_emul/: one OS thread per CUDA thread, barriers for
__syncthreads / warp shuffles, launches rewritten to emul_launch). The tensor-core kernels (tcgen05 / TMA) cannot be emulated;
everything in the exact-fp32 ("simt") numerics mode can. The `-m gpu` run of test_ops.py on a B200 remains the parity test proper:
this run checks the kernel source's indexing, control flow and arithmetic, not the device compiler. Test infrastructure only."""
import pytest

import test_ops as T
from cuda_emul import loader
from test_ops import (test_attention, test_batchnorm_add_relu_fused, test_batchnorm_squeeze_excite_one_node, test_batchnorm_train, test_batchnorm_with_se_pool, test_conv2d, test_gpt_tokens_and_view_quirk, test_gpt_with_velocity_embedding, test_gru_waypoints,  # noqa: F401
                      test_image_prep_and_layout, test_layernorm_linear_dropout, test_losses, test_se_add_pool, test_residual_dropout_layernorm_fused, test_se_fused_mlp_backward, test_small_m_gemm,
                      test_upsample)


@pytest.fixture(autouse=True)
def _emulated_kernels(monkeypatch):
    lib = loader.patch_product(monkeypatch)
    monkeypatch.setattr(T, 'DEV', 'cpu')
    monkeypatch.setattr(T, 'DROPOUT_N', 1 << 16)
    monkeypatch.setattr(T, 'LN_ROWS', 37)
    monkeypatch.setattr(T, 'ATT_T', 46)
    del T.SKIPPED_ON_EMULATOR[:]
    yield
    assert lib.log or T.SKIPPED_ON_EMULATOR, 'the test did not reach the emulated C-ABI'
