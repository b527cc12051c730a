"""GPU parity of the whole predict path through the public API / C-ABI against the
CPU oracle (oracle/): probabilities max-abs <= 1e-2, mask agreement >= 99.9 %,
tile placement / crop / mosaic indexing bit-exact."""
import numpy as np
import pytest

from oracle import normalize as onorm
from oracle import tiling as otile
from oracle import unet as ounet
from satellite_computervision_b200 import model_tools, prediction_tools as pt, processing

pytestmark = pytest.mark.gpu

PROB_TOL = 1e-2       # north_star: bf16 compute, fp32 accumulate
MASK_AGREE = 0.999
# Random-init networks put most probabilities within ~1e-2 of the 0.5 threshold (hard part 1),
# so a class flip there measures rounding noise, not correctness.  For the small random nets the class
# check is therefore: NO disagreement at all wherever the oracle's decision margin exceeds MARGIN, and
# overall agreement reported (and >= 99 %).  The BASELINE-size network is held to the full 99.9 %.
MARGIN = 2.5e-3


def _mk(variant, nch, ncls, filters, seed=0, head_bias=None, head_gain=1.0, **kw):
    specs = ounet.weight_specs(variant, nch, ncls, tuple(filters))
    w = ounet.init_weights(specs, seed=seed, randomize_bn=True, head_bias=head_bias, head_gain=head_gain)
    if variant == 'A':
        m = model_tools.binary_unet(nchannels=nch, filters=list(filters), **kw)
    else:
        m = model_tools.get_unet_model(ncls, nch, filters=list(filters), factors=[2] * len(filters), **kw)
    m.set_weights(w)
    return m, w


@pytest.mark.parametrize('variant,ncls,filters,hw,N', [
    ('A', 1, (32, 64), 64, 3),
    ('B', 2, (32, 64), 64, 2),
    ('A', 1, (32, 64, 128), 96, 2),
    ('A', 1, (32, 64), 128, 2),                  # 128-wide level 0: row-streaming kernel (pool, store, head epilogues)
    ('B', 2, (32, 64), 256, 1),                  # two strips per row, softmax head through the generic head epilogue
    ('B', 3, (32, 64, 128, 256, 512), 96, 5),   # deep levels at 6x6 and 3x3: partial boxes
])
def test_small_models_match_oracle(variant, ncls, filters, hw, N):
    m, w = _mk(variant, 6, ncls, filters, seed=1, outputs='both') if variant == 'A' else _mk(variant, 6, ncls, filters, seed=1)
    x = np.random.default_rng(2).random((N, hw, hw, 6)).astype(np.float32)
    head = 'sigmoid' if variant == 'A' else 'softmax'
    ref_p, ref_c = ounet.forward(x, w, variant, tuple(filters), head=head)
    probs, classes = m.predict(x)
    assert probs.shape == ref_p.shape and classes.shape == ref_c.shape and classes.dtype == np.int32
    err = np.abs(probs - ref_p).max()
    agree = (classes == ref_c).mean()
    margin = np.sort(ref_p, axis=-1)
    margin = np.abs(ref_p[..., 0] - 0.5) if ncls == 1 else margin[..., -1] - margin[..., -2]
    print(variant, filters, 'max|dp|', err, 'class agreement', agree, 'frac with decision margin < 1e-2:',
          (margin < 1e-2).mean())
    assert err <= PROB_TOL
    cls_ref = ref_c[..., 0] if ncls == 1 else ref_c
    cls_got = classes[..., 0] if ncls == 1 else classes
    clear = margin >= MARGIN
    assert clear.mean() > 0.5 and np.array_equal(cls_got[clear], cls_ref[clear])
    assert agree >= 0.99


def test_full_size_tile_variant_a_matches_oracle():
    """One 384x384x6 Sentinel-2-like tile through the BASELINE network (31 M parameters)."""
    m, w = _mk('A', 6, 1, ounet.DEFAULT_FILTERS, seed=0, head_bias=0.0, head_gain=1.0, outputs='both')
    rng = np.random.default_rng(0)
    dn = rng.integers(0, 10000, (2, 384, 384, 6), dtype=np.uint16)
    mm = [(0, 10000)] * 6
    x = np.stack([onorm.rescale_tensor(t.astype(np.float32), moments=mm) for t in dn])
    ref_p, ref_c = ounet.forward(x, w, 'A')
    probs, classes = m.predict(dn, norm=processing.rescale_spec(6, moments=mm))
    d = np.abs(probs - ref_p)
    near = (np.abs(ref_p - 0.5) < 1e-2).mean()
    agree = (classes == ref_c).mean()
    print('full tile: max|dp|', d.max(), 'mean|dp|', d.mean(), 'agree', agree, 'frac within 1e-2 of thr', near,
          'p range', ref_p.min(), ref_p.max())
    assert d.max() <= PROB_TOL
    assert agree >= MASK_AGREE


def test_predict_chips_placement_is_bit_exact():
    """Mosaic path (device gather + stitch) == the reference loop of predict_chips run tile by tile
    through the same engine: same values, same places, zeros elsewhere."""
    m, w = _mk('A', 6, 1, (32, 64), seed=3)
    rng = np.random.default_rng(1)
    H, W, kernel, buff = 500, 613, 64, 32
    arr = rng.integers(0, 10000, (H, W, 6), dtype=np.uint16)
    spec = processing.scalar_spec(6, 10000.0)
    idx = pt.generate_chip_indices(arr, buff, kernel)
    assert idx == otile.generate_chip_indices(arr.shape, buff, kernel) and len(idx) == 7 * 8
    got = pt.predict_chips(arr, idx, np.zeros((H, W)), m, kernel, buff, norm=spec)
    want = otile.predict_chips(arr, idx, np.zeros((H, W)), lambda b: m.predict(b, norm=spec), kernel, buff)
    assert got.dtype == np.float64 and np.array_equal(got, want)
    assert np.all(got[:16] == 0) and np.all(got[:, :16] == 0) and np.all(got[16 + 7 * 64:] == 0)
    # subset / repeated indices take the batched-tiles path and accumulate with +=
    sub = idx[3:11] + idx[3:5]
    got2 = pt.predict_chips(arr, sub, np.full((H, W), 0.25), m, kernel, buff, norm=spec)
    want2 = otile.predict_chips(arr, sub, np.full((H, W), 0.25), lambda b: m.predict(b, norm=spec), kernel, buff)
    assert np.array_equal(got2, want2)
    # and against the fp32 oracle network: tolerance on values
    ref = otile.predict_chips(arr.astype(np.float32) / np.float32(10000.0), idx, np.zeros((H, W)),
                              ounet.make_predict_fn(w, variant='A', filters=(32, 64)), kernel, buff)
    assert np.abs(got - ref).max() <= PROB_TOL


def test_mask_and_tile_row_sharding():
    m, w = _mk('A', 6, 1, (32, 64), seed=4)
    rng = np.random.default_rng(2)
    arr = (rng.random((400, 400, 6)) * 10000).astype(np.float32)
    spec = processing.scalar_spec(6, 10000.0)
    prob, mask = m.predict_mosaic(arr, buff=32, kernel=64, norm=spec)
    assert np.array_equal(mask, (prob > 0.5).astype(np.uint8))
    # two row bands written into the same rasters == one call (multi-GPU sharding contract)
    p2 = np.zeros_like(prob)
    k2 = np.zeros_like(mask)
    m.predict_mosaic(arr, buff=32, kernel=64, norm=spec, tile_rows=(0, 2), out_prob=p2, out_mask=k2)
    assert np.all(p2[16 + 128:] == 0)
    m.predict_mosaic(arr, buff=32, kernel=64, norm=spec, tile_rows=(2, -1), out_prob=p2, out_mask=k2)
    assert np.array_equal(p2, prob) and np.array_equal(k2, mask)


def test_patch_list_geometry_matches_oracle():
    m, w = _mk('B', 6, 2, (32, 64), seed=5)
    rng = np.random.default_rng(3)
    cols, rows, k, b = 3, 2, 64, 32
    patches = rng.random((cols * rows, k + b, k + b, 6)).astype(np.float32)
    mixer = {'patchesPerRow': cols, 'totalPatches': cols * rows, 'patchDimensions': [k, k]}
    preds = m.predict(patches)
    want = otile.make_array_predictions([preds[0], preds[1]], mixer, [k, k], [b, b])
    got = pt.make_array_predictions([p[None] for p in patches], m, mixer, [k, k], [b, b])
    assert np.array_equal(got, want)
    assert np.array_equal(pt.callback_predictions(patches, m, mixer, [k, k], [b, b]),
                          otile.callback_predictions(preds, mixer, [k, k], [b, b]))
    gt, _, _ = pt.geotiff_predictions(patches, m, mixer, [b, b])
    assert np.array_equal(gt, otile.geotiff_stitch(preds, mixer, [b, b]))


def test_overlap_chunk_geometry_matches_oracle():
    m, w = _mk('A', 6, 1, (32, 64), seed=6)
    chw = np.random.default_rng(4).random((6, 128, 192)).astype(np.float32)
    got = pt.predict_overlap_chunks(chw, m, chunk=64, depth=16)
    want = otile.predict_overlap_chunks(chw, lambda b: m.predict(b), chunk=64, depth=16)
    assert np.array_equal(got, want)


def test_errors_are_loud():
    m, _ = _mk('A', 6, 1, (32, 64), seed=0)
    with pytest.raises(ValueError):
        m.predict(np.zeros((1, 50, 50, 6), np.float32))      # not a multiple of 2^levels
    with pytest.raises(ValueError):
        m.predict(np.zeros((1, 64, 64, 5), np.float32))      # wrong band count
    assert pt.predict_chips(np.zeros((96, 96, 6), np.float32), [], np.zeros((96, 96)), m, 64, 32).sum() == 0


def test_keras_h5_weight_file_round_trip(tmp_path):
    """save_weights('.h5') -> a fresh model's load_weights: same predictions bit for bit (N1)."""
    m, w = _mk('A', 6, 1, (32, 64), seed=7)
    x = np.random.default_rng(5).random((2, 64, 64, 6)).astype(np.float32)
    want = m.predict(x)
    for name in ('weights.hdf5', 'weights.npz'):
        path = str(tmp_path / name)
        m.save_weights(path)
        m2 = model_tools.binary_unet(nchannels=6, filters=[32, 64])
        assert not np.array_equal(m2.predict(x), want)
        m2.load_weights(path)
        assert np.array_equal(m2.predict(x), want)
    m3 = model_tools.binary_unet(nchannels=6, filters=[32, 64, 128])
    with pytest.raises(ValueError):
        m3.load_weights(str(tmp_path / 'weights.hdf5'))     # wrong architecture: count / shape mismatch


def test_gee_patch_files_to_geotiff(tmp_path):
    """GEE workflow end to end (N2, N3): .tfrecord.gz patches + mixer.json -> make_pred_dataset ->
    write_geotiff_predictions == the array path; prediction TFRecords hold the cropped patches."""
    import json
    from PIL import Image
    from satellite_computervision_b200 import gee_io
    m, w = _mk('A', 6, 1, (32, 64), seed=8)
    rng = np.random.default_rng(6)
    cols, rows, k, b = 3, 2, 64, 32
    feats = ['B2', 'B3', 'B4', 'B8', 'B11', 'B12']
    patches = (rng.random((cols * rows, k + b, k + b, 6)) * 3000).astype(np.float32)
    gee_io.write_patch_tfrecords(str(tmp_path / 'img00000.tfrecord.gz'), patches[:4], feats)
    gee_io.write_patch_tfrecords(str(tmp_path / 'img00001.tfrecord.gz'), patches[4:], feats)
    mixer = {'projection': {'crs': 'EPSG:32618', 'affine': {'doubleMatrix': [10.0, 0.0, 5e5, 0.0, -10.0, 4.2e6]}},
             'patchDimensions': [k, k], 'patchesPerRow': cols, 'totalPatches': cols * rows}
    jf = str(tmp_path / 'img-mixer.json')
    with open(jf, 'w') as f:
        json.dump(mixer, f)
    mm = [(0.0, 3000.0)] * 6
    files = [str(tmp_path / f'img0000{i}.tfrecord.gz') for i in (1, 0)]
    ds = pt.make_pred_dataset(files, feats, [k, k], [b, b], moments=mm)
    tif = pt.write_geotiff_predictions(ds, m, jf, 'pred', str(tmp_path), [b, b])
    want, _, _ = pt.geotiff_predictions(processing.rescale_tensor(patches, moments=mm), m, mixer, [b, b])
    with Image.open(tif) as im:
        assert np.array_equal(np.array(im), want[..., 0])
    preds = m.predict(pt.make_pred_dataset(files, feats, [k, k], [b, b], moments=mm), steps=cols * rows)
    out = pt.write_tfrecord_predictions(preds, str(tmp_path), 'pred', [k, k], [b, b])
    recs = [gee_io.parse_example(r) for r in gee_io.read_tfrecords(out, verify=True)]
    assert len(recs) == cols * rows
    stitched = np.block([[recs[r * cols + c]['b1'].reshape(k, k) for c in range(cols)] for r in range(rows)])
    assert np.array_equal(stitched, want[..., 0])


def test_raster_tools_per_side_buffer_grid():
    """Per-side-buffer grid of raster_tools (N4): engine == the reference-style loop, placement bit-exact."""
    from satellite_computervision_b200 import raster_tools
    m, w = _mk('A', 6, 1, (32, 64), seed=9)
    arr = (np.random.default_rng(7).random((300, 364, 6)) * 10000).astype(np.uint16)
    spec = processing.scalar_spec(6, 10000.0)
    buff, kernel = 16, 64
    got = raster_tools.predict_chips(arr, m, buff, kernel, norm=spec)
    idx = otile.raster_generate_chip_indices(300, 364, buff, kernel)
    assert len(idx) == 4 * 5
    want = np.zeros((300, 364))
    for y, x in idx:
        chip = arr[y - buff:y + kernel + buff, x - buff:x + kernel + buff]
        p = m.predict(chip[None], norm=spec)[0]
        want[y:y + kernel, x:x + kernel] += p[buff:buff + kernel, buff:buff + kernel, 0]
    assert np.array_equal(got, want)
