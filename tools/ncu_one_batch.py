"""One 63-chip device batch of the BASELINE network through scv_predict_tiles (for ncu captures):
    ncu --set full --import-source on --kernel-name-base demangled -k regex:'...' python tools/ncu_one_batch.py [ntiles]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from satellite_computervision_b200 import model_tools, processing  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 63
m = model_tools.binary_unet(nchannels=6, max_batch=n, outputs='probs')
rng = np.random.default_rng(0)
x = rng.integers(0, 10000, (n, 384, 384, 6), dtype=np.uint16)
p = m.predict(x, norm=processing.scalar_spec(6, 10000.0))
print('ok', p.shape, float(p.mean()), m.times()['n_launches'])
