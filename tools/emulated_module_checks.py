"""Runs the module-level GPU tests that are too slow for the CPU test suite (whole networks: GeometricFusionBackbone training
step, latentTF training step, forward_ego inference) on the CPU emulation of the SIMT kernels (tests/cuda_emul/). By hand, in the
build container:   python tools/emulated_module_checks.py [forward_ego_late|forward_ego_tf|latent|geometric ...]
Each check is one of the functions of tests/test_widen.py, unchanged, with the device switched to 'cpu' and the C-ABI routed to the
emulated library. Expect minutes to tens of minutes per check (one OS thread per CUDA thread)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import pytest  # noqa: E402

import test_widen as W  # noqa: E402
from cuda_emul import loader  # noqa: E402

CHECKS = {
    'forward_ego_late': lambda: W.test_forward_ego_matches_oracle('late_fusion'),
    'forward_ego_tf': lambda: W.test_forward_ego_matches_oracle('transFuser'),
    'latent': W.test_latent_tf_forward_backward_matches_oracle,
    'geometric': W.test_geometric_fusion_forward_backward_matches_oracle,
}

if __name__ == '__main__':
    names = sys.argv[1:] or list(CHECKS)
    mp = pytest.MonkeyPatch()
    lib = loader.patch_product(mp)
    mp.setattr(W, 'DEV', 'cpu')
    for n in names:
        t = time.time()
        lib.log.clear()
        try:
            CHECKS[n]()
            print('%-18s PASSED  %.0f s, %d emulated C-ABI calls' % (n, time.time() - t, len(lib.log)), flush=True)
        except Exception as e:  # noqa: BLE001 — report and go on to the next check
            print('%-18s FAILED  %.0f s: %r' % (n, time.time() - t, e), flush=True)
    mp.undo()
