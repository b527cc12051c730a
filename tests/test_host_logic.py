"""Host-side mirror of the reference interface (no GPU): index generation, chip
slicing, normaliser constants, weight bookkeeping, patch assembly."""
import os

import numpy as np
import pytest

from oracle import normalize as onorm
from oracle import tiling as otile
from oracle import unet as ounet
from satellite_computervision_b200 import _build, _lib, model_tools, prediction_tools as pt, processing


@pytest.fixture(scope='module', autouse=True)
def _built():
    _build.build()


def test_generate_chip_indices_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'chip_indices.npz'))
    for n in range(int(g['ncases'])):
        H, W, buff, kernel = (int(v) for v in g[f'case{n}_params'])
        got = np.array(pt.generate_chip_indices(np.empty((H, W, 1), np.uint8), buff, kernel), dtype=np.int64).reshape(-1, 2)
        assert np.array_equal(got, g[f'case{n}_indices'])


def test_extract_chips_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, 'predict_chips.npz'))
    buff, kernel = int(g['buff']), int(g['kernel'])
    assert np.array_equal(np.stack(pt.extract_chips(g['arr_sq'], buff, kernel)), g['chips_sq'])
    fixed = pt.extract_chips(g['arr_sq'], buff, kernel, legacy_xy_swap=False)
    want = otile.extract_chips(g['arr_sq'], buff, kernel, legacy_xy_swap=False)
    assert all(np.array_equal(a, b) for a, b in zip(fixed, want))


def test_normaliser_constants_reproduce_reference_bits(golden_dir):
    """(x - sub) / div with the spec's float32 constants is bit-identical to the reference output."""
    g = np.load(os.path.join(golden_dir, 'normalize.npz'))
    img = g['img']
    for key, spec in [
        ('rescale_mm', processing.rescale_spec(6, moments=[tuple(r) for r in g['mm']])),
        ('rescale_mm2', processing.rescale_spec(6, moments=[tuple(r) for r in g['mm2']])),
        ('normalize_mv', processing.normalize_spec(6, moments=[tuple(r) for r in g['mv']])),
        ('rescale_split', processing.rescale_spec(6, moments=[(0, 10000)] * 3, splits=[3, 3])),
        ('normalize_split', processing.normalize_spec(6, moments=[(1000.0, 250000.0)] * 2, splits=[2, 2])),
    ]:
        assert spec.mode == _lib.SCV_NORM_PER_BAND
        with np.errstate(divide='ignore', invalid='ignore'):
            got = (img - spec.sub) / spec.div
        assert got.dtype == np.float32
        assert np.array_equal(got, g[key], equal_nan=True), key
    lazy = processing.rescale_tensor(img, moments=[(0, 10000)] * 6)
    assert isinstance(lazy, processing.NormalizedTensor) and lazy.raw is not None
    assert processing.rescale_spec(6).mode == _lib.SCV_NORM_PIXEL_MINMAX          # reference default axes=[2]
    assert processing.normalize_spec(6, axes=[0, 1]).mode == _lib.SCV_NORM_TILE_ZSCORE
    assert processing.rescale_spec(6, axes=[0, 1, 2]).mode == _lib.SCV_NORM_TILE_GLOBAL_MINMAX
    assert processing.normalize_spec(6, axes=[0, 1, 2]).mode == _lib.SCV_NORM_TILE_GLOBAL_ZSCORE
    sp = processing.normalize_spec(6, splits=[2, 3])                              # data-derived, channel 5 passes through
    c = sp.to_c(6)
    assert (c.ngroups, c.group_size[0], c.group_size[1]) == (2, 2, 3)
    assert processing.band_zscore_spec().mode == _lib.SCV_NORM_PIXEL_ZSCORE_SD
    # a length-1 moments list broadcasts over the channels like the reference's numpy arrays do (:304-311)
    one = processing.rescale_spec(6, moments=[(0, 10000)])
    assert np.array_equal(one.sub, np.zeros(6, np.float32)) and one.div.shape == (6,)
    # derived / one-hot planes appended after the rescaled bands pass through, whatever the rescale mode
    wide = processing.rescale_spec(4).with_passthrough(4, 6)
    assert wide.groups == [4] and wide.mode == _lib.SCV_NORM_PIXEL_MINMAX
    widem = processing.rescale_spec(4, moments=[(0, 2)] * 4).with_passthrough(4, 6)
    assert np.array_equal(widem.sub[4:], [0, 0]) and np.array_equal(widem.div[4:], [1, 1])
    with pytest.raises(NotImplementedError):
        processing.rescale_spec(6, axes=[0])
    with pytest.raises(ValueError):
        processing.rescale_spec(6, moments=[(0, 1)] * 5)


def test_model_weight_bookkeeping():
    m = model_tools.binary_unet()
    assert m.count_params() == 31_127_361
    assert m.weight_names == [n for n, _ in ounet.weight_specs('A', 6, 1)]
    w = m.get_weights()
    assert w[0].shape == (3, 3, 6, 32) and w[0].dtype == np.float32
    assert np.all(w[2] == 1) and np.all(w[3] == 0)  # BN gamma, beta keras defaults
    lim = np.sqrt(6.0 / (9 * 6 + 9 * 32))
    assert np.abs(w[0]).max() <= lim
    with pytest.raises(ValueError):
        m.set_weights(w[:-1])
    bad = list(w)
    bad[0] = np.zeros((3, 3, 5, 32), np.float32)
    with pytest.raises(ValueError):
        m.set_weights(bad)
    b = model_tools.get_unet_model(2, 6, bias=[0.1, -0.1])
    assert b.count_params() == 18_537_474 and not b.double_conv
    assert np.allclose(b.get_weights()[-1], [0.1, -0.1])
    with pytest.raises(NotImplementedError):
        model_tools.get_unet_model(2, 6, factors=[3, 2, 2, 2, 2])
    with pytest.raises(AssertionError):
        model_tools.get_unet_model(2, 6, filters=[32, 64], factors=[2])


def test_weights_npz_roundtrip(tmp_path):
    m = model_tools.binary_unet(filters=[32, 64], nchannels=3, seed=1)
    p = str(tmp_path / 'w.npz')
    m.save_weights(p)
    m2 = model_tools.binary_unet(filters=[32, 64], nchannels=3, seed=2)
    m2.load_weights(p)
    assert all(np.array_equal(a, b) for a, b in zip(m.get_weights(), m2.get_weights()))


def test_assemble_matches_reference_np_append(golden_dir):
    g = np.load(os.path.join(golden_dir, 'patch_stitch.npz'))
    from tests.test_oracle_golden import fake_predict
    preds = fake_predict(g['patches'])
    y0, y1, x0, x1 = pt._crop(list(g['kernel_shape']), list(g['kernel_buffer']))
    got = pt._assemble(preds[:, y0:y1, x0:x1, :], int(g['cols']))
    assert np.array_equal(got, g['make_array'])


def test_foreign_models_are_rejected():
    class Keras:
        def predict(self, x):
            return x
    with pytest.raises(TypeError):
        pt.predict_chips(np.zeros((500, 500, 6)), [(64, 64)], np.zeros((500, 500)), Keras())


def test_raster_tools_chip_grid_matches_reference_golden():
    """raster_tools.generate_chip_indices (per-side buffer) == the function lifted from the reference source."""
    import os
    from oracle import tiling as otile
    from satellite_computervision_b200 import raster_tools
    z = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'raster_chip_indices.npz'))
    for n in range(int(z['ncases'])):
        H, W, buff, kernel = (int(v) for v in z[f'case{n}_params'])
        want = [tuple(int(v) for v in r) for r in z[f'case{n}_indices']]
        assert raster_tools.generate_chip_indices(H, W, buff, kernel) == want
        assert otile.raster_generate_chip_indices(H, W, buff, kernel) == want
        for y, x in want:   # every chip fits, the grid reaches the last position that does
            assert y - buff >= 0 and x - buff >= 0 and y + kernel + buff <= H and x + kernel + buff <= W
