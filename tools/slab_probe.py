"""GPU probe: which UMMA descriptor convention makes the halo-slab taps read correctly?
Runs the debug conv entry with the persistent slab kernel forced, for slab_w in {10,16} x base_offset in {0,1}."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import gpu_util as G  # noqa: E402

CASES = [(2, 32, 32, 6, 32), (2, 32, 32, 32, 32), (2, 32, 32, 64, 64), (1, 32, 32, 128, 64), (1, 48, 40, 64, 32)]
os.environ['SCV_SLAB_FORCE'] = '1'
for sw in (10, 16):
    for bo in (0, 1):
        os.environ['SCV_SLAB_W'] = str(sw)
        os.environ['SCV_SLAB_BASEOFF'] = str(bo)
        for (N, H, W, Cin, Cout) in CASES:
            rng = np.random.default_rng(1)
            x = rng.standard_normal((N, H, W, Cin)).astype(np.float32)
            k = (rng.standard_normal((3, 3, Cin, Cout)) / np.sqrt(9 * Cin)).astype(np.float32)
            b = rng.standard_normal(Cout).astype(np.float32) * 0.1
            try:
                got = G.conv3x3_device(x, k, b)
                s = G.err_stats(got, G.conv3x3_ref(x, k, b))
            except Exception as exc:  # noqa: BLE001
                s = {'error': str(exc)}
            print(json.dumps({'slab_w': sw, 'base_off': bo, 'case': (N, H, W, Cin, Cout), **s}), flush=True)
# convT / 1x1 path through the slab kernel (no halo, standard SBO)
os.environ['SCV_SLAB_W'] = '10'
os.environ['SCV_SLAB_BASEOFF'] = '0'
for (N, H, W, Cin, Cout) in [(1, 32, 32, 64, 32), (2, 16, 16, 128, 64)]:
    rng = np.random.default_rng(2)
    x = rng.standard_normal((N, H, W, Cin)).astype(np.float32)
    k = (rng.standard_normal((2, 2, Cout, Cin)) / np.sqrt(Cin)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32) * 0.1
    try:
        s = G.err_stats(G.convT_device(x, k, b), G.convT_ref(x, k, b))
    except Exception as exc:  # noqa: BLE001
        s = {'error': str(exc)}
    print(json.dumps({'convT': (N, H, W, Cin, Cout), **s}), flush=True)
