"""GPU tests of the second-round surface: remaining normaliser forms (a5/a6/a8), non-square crop windows
(a15), zero-padded overlap chunks with a normaliser (a16), float64 / accumulating templates, chip-range
(tile-balanced) sharding, the streaming entry points and several engines in one process."""
import ctypes as C

import numpy as np
import pytest

from oracle import normalize as onorm
from oracle import tiling as otile
from oracle import unet as ounet
from satellite_computervision_b200 import _lib, model_tools, prediction_tools as pt, processing, sharding
from tests.gpu_util import bf16_round

pytestmark = pytest.mark.gpu


def _mk(filters=(32, 64), seed=0, nch=6, **kw):
    w = ounet.init_weights(ounet.weight_specs('A', nch, 1, tuple(filters)), seed=seed, randomize_bn=True)
    m = model_tools.binary_unet(nchannels=nch, filters=list(filters), **kw)
    m.set_weights(w)
    return m, w


def _close_bf16(got, ref, rtol=2.0 ** -7, atol=1e-6):
    """got = bf16 values from K1, ref = the oracle in fp32/fp64: one bf16 rounding + fp32-vs-oracle arithmetic."""
    ref = np.asarray(ref, np.float64)
    both_nan = np.isnan(got) & np.isnan(ref)
    d = np.abs(np.where(both_nan, 0, got) - np.where(both_nan, 0, ref))
    bad = d > rtol * np.abs(np.where(both_nan, 0, ref)) + atol
    assert not (np.isnan(got) ^ np.isnan(ref)).any(), 'NaN pattern differs'
    assert not bad.any(), f'{int(bad.sum())} of {bad.size} values differ, max |d| {d.max()}'


# ------------------------------------------------------------------ normalisers
def test_normalize_dataarray_band_zscore_with_nans():
    """a8: pc_tools.normalize_dataArray(da, 'band') -- nanmean / nanstd across bands, (x-mean)/(sd+1e-6)."""
    rng = np.random.default_rng(0)
    chw = (rng.random((6, 64, 64)) * 3000).astype(np.float64)
    chw[2, 5, 7] = np.nan
    chw[:, 9, 9] = np.nan                       # all-NaN pixel stays NaN
    chw[0, 20:24, :] = np.nan
    with np.errstate(invalid='ignore'), pytest.warns(RuntimeWarning):
        ref = np.moveaxis(onorm.normalize_data_array(chw, axis=0), 0, -1)
    lazy = processing.normalize_dataArray(chw, 'band')
    assert isinstance(lazy, processing.NormalizedTensor) and lazy.shape == (64, 64, 6)
    got = lazy.numpy()
    _close_bf16(got, ref, atol=2e-3)
    assert np.isnan(got[9, 9]).all() and np.isnan(got[5, 7, 2]) and not np.isnan(got[5, 7, 0])
    # float32 input of the same values: same answer up to the input rounding
    got32 = processing.normalize_dataArray(chw.astype(np.float32), 'band').numpy()
    _close_bf16(got32, ref, atol=2e-3)


@pytest.mark.parametrize('kind', ['rescale', 'normalize'])
def test_global_axes_and_data_derived_splits(kind):
    """a5/a6: axes=[0,1,2] (one statistic per tile) and splits= with data-derived statistics."""
    rng = np.random.default_rng(1)
    img = (rng.random((64, 64, 6)) * np.array([1, 2, 3, 4, 5, 6]) * 1000).astype(np.float32)
    fn_o = onorm.rescale_tensor if kind == 'rescale' else onorm.normalize_tensor
    fn = processing.rescale_tensor if kind == 'rescale' else processing.normalize_tensor
    for axes, splits in [([0, 1, 2], None), ([0, 1, 2], [2, 4] if kind == 'rescale' else [2, 3]),
                         ([2], [3, 3] if kind == 'rescale' else [2, 3]), ([0, 1], [3, 3] if kind == 'rescale' else [4])]:
        ref = fn_o(img, axes=tuple(axes), splits=splits)
        got = fn(img, axes=axes, splits=splits).numpy()
        _close_bf16(got, ref, rtol=2.0 ** -6, atol=2e-3)
        if kind == 'normalize' and splits and sum(splits) < 6:  # trailing channels pass through (:269-274)
            assert np.array_equal(got[..., sum(splits):], bf16_round(img[..., sum(splits):]))


def test_lazy_tensor_behaves_like_an_array():
    """Weak #12: arithmetic / np.asarray on the result of rescale_tensor materialises it through K1."""
    rng = np.random.default_rng(2)
    img = rng.integers(0, 10000, (48, 80, 6), dtype=np.uint16)       # non-square: cut into gcd squares
    mm = [(0, 10000)] * 6
    t = processing.rescale_tensor(img, moments=mm)
    want = bf16_round(onorm.rescale_tensor(img.astype(np.float32), moments=mm))
    assert np.array_equal(np.asarray(t), want)
    assert np.array_equal(t * 2.0 - 1.0, want * 2.0 - 1.0)
    assert np.array_equal(t[3:5, 7], want[3:5, 7]) and t.dtype == np.float32 and t.ndim == 3
    stack = processing.rescale_tensor(np.stack([img[:48, :48], img[:48, 32:]]), moments=mm)
    assert np.array_equal(np.asarray(stack), np.stack([want[:48, :48], want[:48, 32:]]))


def test_derived_bands_pass_through_next_to_a_data_dependent_rescale():
    m, w = _mk(nch=6)
    rng = np.random.default_rng(3)
    bands = (rng.random((2, 64, 64, 4)) * 3000).astype(np.float32)
    extra = rng.random((2, 64, 64, 2)).astype(np.float32)
    stacked = np.concatenate([bands, extra], axis=-1)
    spec = processing.rescale_spec(4, axes=[2]).with_passthrough(4, 6)
    got = processing.NormalizedTensor(stacked, spec).numpy()
    ref = np.concatenate([np.stack([onorm.rescale_tensor(b, axes=(2,)) for b in bands]), extra], axis=-1)
    _close_bf16(got, ref, atol=1e-4)
    assert np.array_equal(got[..., 4:], bf16_round(extra))
    assert m.predict(processing.NormalizedTensor(stacked, spec)).shape == (2, 64, 64, 1)


# ------------------------------------------------------------------ geometries
def test_non_square_kernel_buffer_crop(golden_dir):
    """a15: the crop window of :258-261 with kernel_buffer = [8, 16] (x / y names mixed as written)."""
    import os
    g = np.load(os.path.join(golden_dir, 'patch_stitch.npz'))
    ks, kb = [int(v) for v in g['kernel_shape']], [int(v) for v in g['kernel_buffer_ns']]
    assert g['make_array_ns'].shape[:2] == (2 * 28, 3 * 36)             # what the reference produces
    m, w = _mk(nch=6)
    rng = np.random.default_rng(4)
    patches = rng.random((6, 64, 64, 6)).astype(np.float32)
    mixer = {'patchesPerRow': 3, 'totalPatches': 6, 'patchDimensions': [48, 48]}
    kb2 = [16, 32]
    preds = m.predict(patches)
    want = otile.make_array_predictions(preds, mixer, [48, 48], kb2)
    assert want.shape == (2 * (48 + 8 - 16), 3 * (48 + 16 - 8), 1)
    prob, mask = m.predict_patches(patches, 3, [48, 48], kb2, want_mask=True)
    assert np.array_equal(prob, want[..., 0]) and np.array_equal(mask, (prob > 0.5).astype(np.uint8))
    assert np.array_equal(pt.make_array_predictions(patches, m, mixer, [48, 48], kb2), want)
    # geotiff stitcher (:496-520): its own (consistent) window
    gt, _, _ = pt.geotiff_predictions(patches, m, mixer, [16, 16])
    assert np.array_equal(gt, otile.geotiff_stitch(preds, mixer, [16, 16]))


def test_overlap_chunks_pad_after_normalisation():
    """ADVICE: map_overlap(boundary=0) pads the NORMALISED raster; with a mean/std normaliser the halo beyond
    the raster edge must be exact zeros, not (0 - mean) / std."""
    m, w = _mk(seed=6)
    rng = np.random.default_rng(5)
    chw = (rng.random((6, 128, 192)) * 3000 + 500).astype(np.float32)
    mv = [(1500.0 + 100 * c, 250000.0) for c in range(6)]
    spec = processing.normalize_spec(6, moments=mv)
    normed = np.moveaxis(onorm.normalize_tensor(np.moveaxis(chw, 0, -1), moments=mv), -1, 0)
    want = otile.predict_overlap_chunks(normed, lambda b: m.predict(b), chunk=64, depth=16)
    got = pt.predict_overlap_chunks(chw, m, chunk=64, depth=16, norm=spec)
    assert np.abs(got - want).max() <= 2e-3       # same engine, inputs equal up to fp32 normaliser rounding
    wrong = m.predict_mosaic(np.pad(np.moveaxis(chw, 0, -1), ((16, 16 + 64), (16, 16 + 64), (0, 0))), 32, 64, norm=spec,
                             want_mask=False)[0][16:16 + 128, 16:16 + 192]
    assert np.abs(wrong - want).max() > np.abs(got - want).max()   # the un-windowed call is what ADVICE flagged
    band = processing.normalize_dataArray(chw, 'band')              # lazy tensor in: same treatment
    pt.predict_overlap_chunks(np.moveaxis(band.raw, -1, 0), m, 64, 16, norm=band.norm)
    with pytest.raises(ValueError):
        pt.predict_overlap_chunks(chw, m, 64, 16, norm=processing.normalize_spec(6, axes=[0, 1]))


# ------------------------------------------------------------------ templates, sharding, streaming
def test_float64_accumulating_template_and_chip_range_sharding():
    m, w = _mk(seed=7)
    rng = np.random.default_rng(6)
    H, W, kernel, buff = 500, 613, 64, 32
    arr = rng.integers(0, 10000, (H, W, 6), dtype=np.uint16)
    spec = processing.scalar_spec(6, 10000.0)
    idx = pt.generate_chip_indices(arr, buff, kernel)
    base = rng.random((H, W))                                          # non-zero float64 template: += accumulates
    want = otile.predict_chips(arr, idx, base.copy(), lambda b: m.predict(b, norm=spec), kernel, buff)
    got = pt.predict_chips(arr, idx, base.copy(), m, kernel, buff, norm=spec)
    assert got.dtype == np.float64 and np.array_equal(got, want)
    base32 = base.astype(np.float32)
    got32 = pt.predict_chips(arr, idx, base32.copy(), m, kernel, buff, norm=spec)
    want32 = otile.predict_chips(arr, idx, base32.copy(), lambda b: m.predict(b, norm=spec), kernel, buff)
    assert got32.dtype == np.float32 and np.array_equal(got32, want32)
    # tile-balanced shards: chip ranges that start and end in the middle of tile rows
    full, fmask = m.predict_mosaic(arr, buff, kernel, norm=spec)
    world = 5
    p = np.zeros((H, W), np.float32)
    k = np.zeros((H, W), np.uint8)
    shards = [sharding.rank_shard(H, W, kernel, buff, r, world) for r in range(world)]
    assert sum(s.n_chips for s in shards) == len(idx) and max(s.n_chips for s in shards) - min(s.n_chips for s in shards) <= 1
    for s in shards:
        m.predict_mosaic(arr, buff, kernel, norm=spec, tile_range=(s.tile_begin, s.tile_end), out_prob=p, out_mask=k)
    assert np.array_equal(p, full) and np.array_equal(k, fmask)
    rects = [r for s in shards for r in sharding.shard_rects(s, H, W, buff)]
    cover = np.zeros((H, W), np.int32)
    for y0, y1, x0, x1 in rects:
        cover[y0:y1, x0:x1] += 1
    assert cover.max() == 1 and np.array_equal(cover == 1, full != 0)


def test_streamed_scenes_equal_synchronous_calls():
    """scv_stream_submit / scv_stream_wait (BASELINE configs[4]): three scenes in flight two at a time."""
    m, w = _mk(seed=8, max_batch=16)
    lib, eng = m._lib, m._ensure_engine()
    rng = np.random.default_rng(7)
    H, W = 300, 420
    spec = processing.scalar_spec(6, 10000.0).to_c(6)
    t = _lib.Tiling(64, 32)
    scenes, probs, masks, want = [], [], [], []
    for i in range(3):
        s = _lib.pinned_zeros((H, W, 6), np.uint16)
        s[...] = rng.integers(0, 10000, (H, W, 6), dtype=np.uint16)
        scenes.append(s)
        probs.append(_lib.pinned_zeros((H, W), np.float32))
        masks.append(_lib.pinned_zeros((H, W), np.uint8))
        want.append(m.predict_mosaic(np.array(s), 32, 64, norm=processing.scalar_spec(6, 10000.0)))
    o = _lib.MosaicOpts()
    tickets = []
    for i in range(3):
        tk = C.c_int(-1)
        _lib.check(lib.scv_stream_submit(eng, _lib.ptr(scenes[i]), _lib.SCV_U16, H, W, 6, C.byref(t), C.byref(spec),
                                         C.byref(o), _lib.ptr(probs[i]), _lib.ptr(masks[i]), C.byref(tk)))
        tickets.append(tk.value)
    assert tickets == [tickets[0], tickets[0] + 1, tickets[0] + 2]
    _lib.check(lib.scv_stream_wait(eng, tickets[0]))
    assert np.array_equal(probs[0], want[0][0])
    _lib.check(lib.scv_stream_wait(eng, -1))
    for i in range(3):
        assert np.array_equal(probs[i], want[i][0]) and np.array_equal(masks[i], want[i][1])
    _lib.check(lib.scv_check(eng))


def test_two_engines_on_two_devices_in_one_process():
    """ADVICE: per-device kernel attributes -- a second engine on another GPU of the same process."""
    lib = _lib.load_library()
    if lib.scv_device_count() < 2:
        pytest.skip('needs two GPUs')
    x = np.random.default_rng(9).random((2, 128, 128, 6)).astype(np.float32)
    outs = []
    for dev in (0, 1):
        m, w = _mk(seed=10, device=dev)
        outs.append(m.predict(x))
    assert np.array_equal(outs[0], outs[1])


def test_api_keeps_the_callers_current_device():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    torch.cuda.set_device(1)
    m, w = _mk(seed=11, device=0)
    m.predict(np.zeros((1, 64, 64, 6), np.float32))
    assert torch.cuda.current_device() == 1
    torch.cuda.set_device(0)
