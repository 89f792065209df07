"""ORACLE / TEST INFRASTRUCTURE ONLY. Generates the committed fixtures under tests/golden/ by executing the REFERENCE
itself (imported verbatim from /root/reference via oracle/ref_import.py) on seeded inputs. Run in the build container:
    python oracle/make_golden.py [bev] [model]
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import bev_oracle, ref_import  # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')


def make_bev():
    ref = ref_import.load_histogram_fn()
    out = {}
    for dt in ('float32', 'float64'):
        for seed, n in ((0, 40000), (7, 2000)):
            pts = bev_oracle.synthetic_points(n, seed, np.dtype(dt).type)
            out['out_%s_%d_%d' % (dt, seed, n)] = ref(pts)
    np.savez_compressed(os.path.join(GOLD, 'bev_hist.npz'), **out)
    print('wrote bev_hist.npz', {k: float(v.sum()) for k, v in out.items()})


if __name__ == '__main__':
    what = sys.argv[1:] or ['bev', 'model']
    os.makedirs(GOLD, exist_ok=True)
    if 'bev' in what:
        make_bev()
    if 'model' in what:
        from oracle import make_golden_model
        make_golden_model.main()
