"""GPU parity of the siamese U-Net + ASPP (make_siamese_unet, utils/model_tools.py:533-663) through the public API /
C-ABI against the CPU oracle: probabilities max-abs <= 1e-2, classes identical wherever the oracle's decision margin
is clear; the dilated / 1x1 conv kernels on their own against torch."""
import numpy as np
import pytest

from oracle import siamese as osi
from oracle import tiling as otile
from oracle import unet as ounet
from satellite_computervision_b200 import model_tools, prediction_tools as pt

pytestmark = pytest.mark.gpu

PROB_TOL = 1e-2
MARGIN = 2.5e-3


def _mk(nch, filters, seed=0, **kw):
    specs = osi.weight_specs(nch, tuple(filters))
    w = ounet.init_weights(specs, seed=seed, randomize_bn=True, head_gain=4.0)
    m = model_tools.make_siamese_unet(nch, list(filters), [2] * len(filters), **kw)
    m.set_weights(w)
    return m, w


def _check(probs, classes, ref_p, ref_c, what):
    assert probs.shape == ref_p.shape and classes.shape == ref_c.shape and classes.dtype == np.int32
    err = np.abs(probs - ref_p).max()
    agree = (classes == ref_c).mean()
    margin = np.abs(ref_p - 0.5)
    print(what, 'max|dp|', err, 'class agreement', agree, 'frac with margin < 1e-2', (margin < 1e-2).mean(),
          'p range', ref_p.min(), ref_p.max())
    assert err <= PROB_TOL
    clear = margin >= MARGIN
    assert clear.mean() > 0.5 and np.array_equal(classes[clear], ref_c[clear])
    assert agree >= 0.99


@pytest.mark.parametrize('nch,filters,hw,N', [
    (3, (32, 64, 128), 96, 3),     # ASPP at 12 x 12: rates 3 and 6 reach real pixels, 12 only the centre tap
    (6, (32, 64), 64, 2),          # 12 stacked bands (16-channel first layer), ASPP at 16 x 16
    (3, (32, 64, 128), 384, 2),    # the reference's default filters at the 384-pixel tile: row / slab kernels at level 0 / 1
    (4, (64, 128), 128, 1),        # 64 filters at level 0: 192-channel concat tensor
])
def test_siamese_matches_oracle(nch, filters, hw, N):
    m, w = _mk(nch, filters, seed=2)
    rng = np.random.default_rng(3)
    a = rng.random((N, hw, hw, nch)).astype(np.float32)
    b = (a + 0.25 * rng.random((N, hw, hw, nch))).astype(np.float32)
    ref_p, ref_c = osi.forward(a, b, w, tuple(filters))
    probs, classes = m.predict([a, b])
    _check(probs, classes, ref_p, ref_c, f'siamese {filters} {hw}')
    # the stacked form every tiled entry point uses is the same call
    p2, c2 = m.predict(np.concatenate([a, b], -1))
    assert np.array_equal(p2, probs) and np.array_equal(c2, classes)
    # and the inputs are ordered
    p3, _ = m.predict([b, a])
    assert np.abs(p3 - probs).max() > 1e-4


def test_siamese_tiled_mosaic_matches_oracle_loop():
    """generate_chip_indices + predict_chips (utils/prediction_tools.py:87-156) over a stacked two-date raster."""
    nch, filters = 3, (32, 64, 128)
    m, w = _mk(nch, filters, seed=4, outputs='probs', max_batch=4)
    rng = np.random.default_rng(5)
    scene = rng.random((400, 528, 2 * nch)).astype(np.float32)
    idx = pt.generate_chip_indices(scene, 64, 128)
    assert len(idx) > 4
    ref = otile.predict_chips(scene, idx, np.zeros(scene.shape[:2]), osi.make_predict_fn(w, nch, filters=tuple(filters)),
                              kernel=128, buff=64)
    got = pt.predict_chips(scene, idx, np.zeros(scene.shape[:2]), m, kernel=128, buff=64)
    assert np.array_equal(got == 0, ref == 0)  # footprint bit-exact
    assert np.abs(got - ref).max() <= PROB_TOL
