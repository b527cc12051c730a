"""Pin the oracle (oracle/) against fixtures produced by the REAL reference
source (tests/golden/make_golden.py; prediction_tools / processing /
array_tools imported with TF, matplotlib, rasterio stubbed)."""
import os

import numpy as np
import pytest

from oracle import normalize as onorm
from oracle import tiling as otile


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + '.npz'))


def fake_predict(x, as_list=False):
    """Same element-wise stand-in model as make_golden.FakeModel."""
    x = np.asarray(x, dtype=np.float32)
    w = np.float32(0.5)
    p0 = np.zeros(x.shape[:-1], dtype=np.float32)
    for c in range(x.shape[-1]):
        p0 = p0 + x[..., c] * w
        w = w * np.float32(0.5)
    probs = np.stack([p0, np.float32(1.0) - p0], axis=-1)
    if as_list:
        return [probs, (probs[..., :1] > 0.25).astype(np.int32)]
    return probs


def test_generate_chip_indices_matches_reference(golden_dir):
    g = _load(golden_dir, 'chip_indices')
    for n in range(int(g['ncases'])):
        H, W, buff, kernel = (int(v) for v in g[f'case{n}_params'])
        got = np.array(otile.generate_chip_indices((H, W, 1), buff, kernel), dtype=np.int64).reshape(-1, 2)
        assert np.array_equal(got, g[f'case{n}_indices']), (H, W, buff, kernel)


def test_baseline_tile_counts():
    assert len(otile.generate_chip_indices((2048, 2048, 6))) == 49
    assert len(otile.generate_chip_indices((10980, 10980, 6))) == 1764
    assert otile.generate_chip_indices((384, 384, 6)) == []
    assert otile.generate_chip_indices((449, 449, 6)) == [(64, 64)]


def test_predict_chips_matches_reference(golden_dir):
    g = _load(golden_dir, 'predict_chips')
    arr, buff, kernel = g['arr'], int(g['buff']), int(g['kernel'])
    indices = otile.generate_chip_indices(arr.shape, buff, kernel)
    assert np.array_equal(np.array(indices), g['indices'])
    template = np.zeros(arr.shape[:2])
    res = otile.predict_chips(arr, indices, template, fake_predict, kernel, buff)
    assert res is template and res.dtype == np.float64
    assert np.array_equal(res, g['template'])
    t2 = np.full(arr.shape[:2], 0.5)
    res2 = otile.predict_chips(arr, [tuple(i) for i in g['indices2']], t2, fake_predict, kernel, buff)
    assert np.array_equal(res2, g['template2'])


def test_extract_chips_matches_reference_including_xy_swap(golden_dir):
    g = _load(golden_dir, 'predict_chips')
    buff, kernel = int(g['buff']), int(g['kernel'])
    chips = otile.extract_chips(g['arr'], buff, kernel, legacy_xy_swap=True)
    assert np.array_equal(np.array([c.shape for c in chips]), g['chips_shapes'])
    chips_sq = otile.extract_chips(g['arr_sq'], buff, kernel, legacy_xy_swap=True)
    assert np.array_equal(np.stack(chips_sq), g['chips_sq'])
    fixed = otile.extract_chips(g['arr_sq'], buff, kernel, legacy_xy_swap=False)
    assert not np.array_equal(np.stack(fixed), g['chips_sq'])  # the quirk is real


def test_patch_stitch_matches_reference(golden_dir):
    g = _load(golden_dir, 'patch_stitch')
    patches = g['patches']
    mixer = {'patchesPerRow': int(g['cols']), 'totalPatches': int(g['cols'] * g['rows']),
             'patchDimensions': [int(v) for v in g['kernel_shape']]}
    ks, kb = list(g['kernel_shape']), list(g['kernel_buffer'])
    preds = fake_predict(patches)
    assert np.array_equal(otile.make_array_predictions(preds, mixer, ks, kb), g['make_array'])
    assert np.array_equal(otile.make_array_predictions(fake_predict(patches, True), mixer, ks, kb), g['make_array_list'])
    assert np.array_equal(otile.callback_predictions(preds, mixer, ks, kb), g['callback'])
    assert np.array_equal(otile.callback_predictions(fake_predict(patches, True), mixer, ks, kb), g['callback_list'])
    gt = otile.geotiff_stitch(preds, mixer, kb)
    assert np.array_equal(np.transpose(gt, (2, 0, 1)), g['geotiff'])
    assert list(g['geotiff_wh']) == [gt.shape[1], gt.shape[0], 1]
    assert np.array_equal(otile.make_array_predictions(preds, mixer, ks, list(g['kernel_buffer_ns'])), g['make_array_ns'])


def test_single_column_mosaic_is_fixed():
    # reference crashes for cols == 1 (x % 1 == 1 never true); oracle implements the evident placement
    preds = np.arange(3 * 48 * 48 * 2, dtype=np.float32).reshape(3, 48, 48, 2)
    out = otile.make_array_predictions(preds, {'patchesPerRow': 1, 'totalPatches': 3}, [32, 32], [16, 16])
    assert out.shape == (96, 32, 2)
    assert np.array_equal(out[32:64], preds[1, 8:40, 8:40])


def test_normalisers_match_reference(golden_dir):
    g = _load(golden_dir, 'normalize')
    img = g['img']
    mm = [tuple(r) for r in g['mm']]
    mm2 = [tuple(r) for r in g['mm2']]
    mv = [tuple(r) for r in g['mv']]
    for got, key in [
        (onorm.rescale_tensor(img, moments=mm), 'rescale_mm'),
        (onorm.rescale_tensor(g['img_u16'], moments=mm), 'rescale_mm_u16'),
        (onorm.rescale_tensor(img, moments=mm2), 'rescale_mm2'),
        (onorm.normalize_tensor(img, moments=mv), 'normalize_mv'),
        (onorm.rescale_tensor(img, moments=[(0, 10000)] * 3, splits=[3, 3]), 'rescale_split'),
        (onorm.normalize_tensor(img, moments=[(1000.0, 250000.0)] * 2, splits=[2, 2]), 'normalize_split'),
    ]:
        assert got.dtype == g[key].dtype, key
        assert np.array_equal(got, g[key], equal_nan=True), key


def test_data_derived_normalisers_agree_with_numpy_twins(golden_dir):
    # array_tools twins use std+eps (not sqrt(var+eps)) for normalize; rescale is identical
    g = _load(golden_dir, 'normalize')
    img = g['img']
    assert np.array_equal(onorm.rescale_tensor(img, axes=(0, 1)), g['at_rescale_axes01'])
    assert np.array_equal(onorm.rescale_tensor(img, axes=(2,)), g['at_rescale_axes2'])
    np.testing.assert_allclose(onorm.normalize_tensor(img, axes=(0, 1)), g['at_normalize_axes01'], rtol=2e-5, atol=2e-5)
    np.testing.assert_allclose(onorm.normalize_tensor(img, axes=(2,)), g['at_normalize_axes2'], rtol=2e-5, atol=2e-5)


def test_overlap_chunks_geometry():
    rng = np.random.default_rng(0)
    chw = rng.random((3, 64, 96), dtype=np.float32)
    ident = lambda b: b[..., :1] * 2.0
    out = otile.predict_overlap_chunks(chw, ident, chunk=32, depth=8)
    assert np.array_equal(out, chw[0] * 2.0)
    assert otile.trim_extent(10980, 256) == 10752 and otile.trim_extent(512, 256) == 512
