"""Probe of the 8-channel (no-swizzle, tap-pair) first-layer path: one-hot weights per tap."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import gpu_util as G

rng = np.random.default_rng(0)
N, H, W, Cin, Cout = 1, 16, 8, 8, 32
x = rng.integers(1, 9, (N, H, W, Cin)).astype(np.float32)
for tap in range(9):
    k = np.zeros((3, 3, Cin, Cout), np.float32)
    for c in range(Cin):
        k[tap // 3, tap % 3, c, c] = 1.0
    b = np.zeros(Cout, np.float32)
    got = G.conv3x3_device(x, k, b, relu=False)
    ref = G.conv3x3_ref(x, k, b, relu=False)
    s = G.err_stats(got, ref)
    # which input tap does the output actually equal?
    match = []
    for t2 in range(9):
        k2 = np.zeros_like(k)
        for c in range(Cin):
            k2[t2 // 3, t2 % 3, c, c] = 1.0
        if np.array_equal(G.conv3x3_ref(x, k2, b, relu=False)[..., :8], got[..., :8]):
            match.append(t2)
    print(json.dumps({'tap': tap, 'n_bad': s['n_bad'], 'max_abs': s['max_abs'], 'matches_tap': match,
                      'got_nonzero_channels': np.nonzero(np.abs(got).sum((0, 1, 2)))[0].tolist()[:12],
                      'got[0,5,3,:4]': got[0, 5, 3, :4].tolist(), 'ref[0,5,3,:4]': ref[0, 5, 3, :4].tolist()}), flush=True)
