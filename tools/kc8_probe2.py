import os, sys, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import gpu_util as G
N, H, W, Cin, Cout = 1, 16, 8, 8, 32
x = np.ones((N, H, W, Cin), np.float32)
k = np.ones((3, 3, Cin, Cout), np.float32)
for bias in (0.0, 1.0):
    b = np.full(Cout, bias, np.float32)
    got = G.conv3x3_device(x, k, b, relu=False)
    print('bias', bias, 'got[0,5,3,:3]', got[0, 5, 3, :3], 'got[0,0,0,:3]', got[0, 0, 0, :3], 'unique', np.unique(got)[:10])
# only channel c nonzero in input
for c in (0, 7):
    x2 = np.zeros_like(x); x2[..., c] = 1
    got = G.conv3x3_device(x2, k, np.zeros(Cout, np.float32), relu=False)
    print('chan', c, 'unique', np.unique(got)[:10])
