"""Tiled prediction with the reference's call surface (``utils/prediction_tools.py``),
running on the sm_100a engine: one gather+normalise kernel, the U-Net as
tcgen05 implicit GEMMs, and one head+crop+stitch kernel per device batch instead
of the reference's per-tile Python loop around ``model.predict``.
"""
from __future__ import annotations

import ctypes as C
import json

import numpy as np

from . import _lib
from .model_tools import UNetModel
from .processing import TILE_STAT_MODES, NormalizedTensor, NormSpec


def generate_chip_indices(arr, buff=128, kernel=256):
    """``utils/prediction_tools.py:87-109`` -- identical: row-major (y, x) upper-left corners of the
    kept ``kernel`` cores; ``range(buff//2, H - (buff+kernel), kernel)`` in both axes (so the top/left
    ``buff//2`` and a bottom/right margin stay unpredicted)."""
    shape = arr.shape if hasattr(arr, 'shape') else tuple(arr)
    H, W = shape[0], shape[1]
    side = buff + kernel
    x_buff = y_buff = buff // 2
    y_indices = list(range(y_buff, H - side, kernel))
    x_indices = list(range(x_buff, W - side, kernel))
    return [(y_index, x_index) for y_index in y_indices for x_index in x_indices]


def extract_chips(arr, buff=128, kernel=256, legacy_xy_swap=True):
    """``utils/prediction_tools.py:111-131``: list of (kernel+buff)^2 views.  As committed the
    reference unpacks its (y, x) tuples as ``for x, y in ...`` (``:127``), cutting every chip at the
    transposed origin; ``legacy_xy_swap=True`` (default) keeps that behaviour bit for bit,
    ``False`` is the evident intent (Appendix A3).  Pure slicing, no arithmetic."""
    raw = arr.raw if isinstance(arr, NormalizedTensor) else arr
    x_buff = y_buff = buff // 2
    chips = []
    for a, b in generate_chip_indices(raw, buff, kernel):
        x, y = (a, b) if legacy_xy_swap else (b, a)
        chips.append(raw[y - y_buff:y + kernel + y_buff, x - x_buff:x + kernel + x_buff, :])
    return chips


def _model(m):
    if not isinstance(m, UNetModel):
        raise TypeError('m must be a satellite_computervision_b200.model_tools.UNetModel '
                        '(the CUDA engine); there is no CPU / Keras fallback on this path')
    return m


def predict_chips(arr, chip_indices, template, m, kernel=256, buff=128, norm=None, channel=0):
    """``utils/prediction_tools.py:133-156``: for every (y, x) in ``chip_indices`` predict the
    buffered chip and ``template[y:y+kernel, x:x+kernel] += preds[0, b:b+kernel, b:b+kernel, 0]``.
    ``template`` (float64 zeros in the reference, ``:769``) is mutated in place and returned.

    When ``chip_indices`` is the full grid of ``generate_chip_indices`` the whole raster goes through
    the device mosaic path (gather, network and crop+stitch fused per batch); any other index list
    (subsets, repeats, overlaps -- ``+=`` then accumulates) is gathered on the host and predicted in
    device batches.  Results are identical either way."""
    m = _model(m)
    raw, norm = UNetModel._split_norm(arr, norm)
    raw = np.asarray(raw)
    chip_indices = [tuple(int(v) for v in i) for i in chip_indices]
    if len(chip_indices) < 1:
        return template
    y_buff = x_buff = buff // 2
    if chip_indices == generate_chip_indices(raw, buff, kernel):
        direct = (isinstance(template, np.ndarray) and template.shape == raw.shape[:2] and template.flags.c_contiguous
                  and template.flags.writeable and template.dtype in (np.float32, np.float64))
        if direct:
            # the stitch kernel does `template[core] += p` itself, in the template's own dtype (float64 as at :769)
            m.predict_mosaic(raw, buff=buff, kernel=kernel, norm=norm, out_channel=channel, want_mask=False,
                             out_prob=template, accumulate=True)
            return template
        prob, _ = m.predict_mosaic(raw, buff=buff, kernel=kernel, norm=norm, out_channel=channel, want_mask=False)
        ys = sorted({y for y, _ in chip_indices})
        xs = sorted({x for _, x in chip_indices})
        y0, y1, x0, x1 = ys[0], ys[-1] + kernel, xs[0], xs[-1] + kernel
        template[y0:y1, x0:x1] += prob[y0:y1, x0:x1]
        return template
    side = kernel + buff
    step = max(1, m.max_batch)
    for s in range(0, len(chip_indices), step):
        part = chip_indices[s:s + step]
        chips = np.stack([raw[y - y_buff:y + kernel + y_buff, x - x_buff:x + kernel + x_buff, :] for y, x in part])
        if chips.shape[1:3] != (side, side):
            raise ValueError('a chip index leaves the raster')
        preds = m.predict(chips, norm=norm)
        if isinstance(preds, list):  # two-output model: use the probabilities (Appendix A4)
            preds = preds[0]
        for (y, x), p in zip(part, preds):
            template[y:y + kernel, x:x + kernel] += p[y_buff:(kernel + y_buff), x_buff:(kernel + x_buff), channel]
    return template


def predict_mosaic(arr, m, buff=128, kernel=256, norm=None, channel=0):
    """The compute part of ``predict_pc_local`` / ``run_local`` (``utils/prediction_tools.py:767-776``;
    ``utils/pc_tools.py:655-664``): indices -> zero template -> predict_chips."""
    raw = arr.raw if isinstance(arr, NormalizedTensor) else np.asarray(arr)
    indices = generate_chip_indices(raw, buff, kernel)
    template = np.zeros((raw.shape[0], raw.shape[1]))
    return predict_chips(arr, indices, template, m, kernel, buff, norm=norm, channel=channel)


# ---------------------------------------------------------------- patch-list geometry
def _load_mixer(json_or_dict):
    if isinstance(json_or_dict, dict):
        return json_or_dict
    with open(json_or_dict) as f:
        return json.load(f)


def _collect(imageDataset, steps=None):
    """(N,h,w,C) ndarray / NormalizedTensor, or an iterable of (1,h,w,C) batches (``.batch(1)``)."""
    if isinstance(imageDataset, (np.ndarray, NormalizedTensor)):
        return UNetModel._split_norm(imageDataset, None)
    raws, norm = [], None
    for i, b in enumerate(imageDataset):
        if steps is not None and i >= steps:
            break
        r, n = UNetModel._split_norm(b, None)
        raws.append(np.asarray(r))
        norm = norm or n
    return np.concatenate(raws, axis=0), norm


def _crop(kernel_shape, kernel_buffer):
    # literal restatement of :258-261 / :340-343 (the reference swaps x/y names; square kernels only matter)
    x_buffer = int(kernel_buffer[0] / 2)
    y_buffer = int(kernel_buffer[1] / 2)
    x_size = kernel_shape[0] + y_buffer
    y_size = kernel_shape[1] + x_buffer
    return y_buffer, y_size, x_buffer, x_size


def _assemble(patches, cols):
    """Row-major placement of equally sized cropped patches (what the np.append loops of
    ``:269-291, :351-373`` build), single allocation.  cols == 1 works (Appendix A6)."""
    n = len(patches)
    rows = n // cols
    ph, pw = patches.shape[1:3]
    out = np.empty((rows * ph, cols * pw) + patches.shape[3:], dtype=patches.dtype)
    for i in range(rows * cols):
        r, c = divmod(i, cols)
        out[r * ph:(r + 1) * ph, c * pw:(c + 1) * pw] = patches[i]
    return out


def make_array_predictions(imageDataset, model, jsonFile, kernel_shape=[256, 256], kernel_buffer=[128, 128],
                           norm=None):
    """``utils/prediction_tools.py:293-373``: (rows*k, cols*k, channels) array; a two-output model's
    ``[probs, classes]`` are concatenated on the last axis first (``:336-338``; a 3-D ``classes`` gets
    a trailing axis, Appendix A7)."""
    model = _model(model)
    mixer = _load_mixer(jsonFile)
    patches, cols = mixer['totalPatches'], mixer['patchesPerRow']
    x, n = _collect(imageDataset, patches)
    predictions = model.predict(x, norm=norm or n)
    if isinstance(predictions, list):
        parts = [p if p.ndim == 4 else p[..., None] for p in predictions]
        predictions = np.concatenate([p.astype(np.float32) for p in parts], axis=3)
    y0, y1, x0, x1 = _crop(kernel_shape, kernel_buffer)
    return _assemble(predictions[:, y0:y1, x0:x1, :], cols)


def callback_predictions(imageDataset, model, mixer, kernel_shape=[256, 256], kernel_buffer=[128, 128], norm=None):
    """``utils/prediction_tools.py:245-291``: keeps probability channel 1 (``:267``); crop + placement run
    in the device stitch kernel."""
    model = _model(model)
    mixer = _load_mixer(mixer)
    patches, cols = mixer['totalPatches'], mixer['patchesPerRow']
    x, n = _collect(imageDataset, patches)
    prob, _ = model.predict_patches(x, cols, kernel_shape, kernel_buffer, norm=norm or n, out_channel=1)
    return prob


def geotiff_predictions(imageDataset, model, jsonFile, kernel_buffer=[128, 128], norm=None):
    """Compute part of ``write_geotiff_predictions`` (``utils/prediction_tools.py:475-520``): the
    (rows*k, cols*k, 1) float32 raster of probability channel 0 that the reference hands to rasterio,
    plus the (affine, crs) it writes with."""
    model = _model(model)
    mixer = _load_mixer(jsonFile)
    ppr, tp = mixer['patchesPerRow'], mixer['totalPatches']
    kernel_shape = mixer['patchDimensions']
    x, n = _collect(imageDataset, tp)
    # this stitcher has its own window (:503-506, :520): rows [kb[1]/2, kb[1]/2 + ks[0]), cols [kb[0]/2, kb[0]/2 + ks[1])
    x_buffer, y_buffer = int(kernel_buffer[0] / 2), int(kernel_buffer[1] / 2)
    crop = (y_buffer, y_buffer + kernel_shape[0], x_buffer, x_buffer + kernel_shape[1])
    prob, _ = model.predict_patches(x, ppr, kernel_shape, kernel_buffer, norm=norm or n, out_channel=0, crop=crop)
    proj = mixer.get('projection', {})
    return prob[..., None], proj.get('affine', {}).get('doubleMatrix'), proj.get('crs')


# ---------------------------------------------------------------- GEE patch files in / rasters out
class PatchDataset:
    """What ``make_pred_dataset`` returns: an iterable of ``(1, h, w, C)`` NormalizedTensor batches
    (``.batch(1)``, ``utils/prediction_tools.py:223-226``) that ``UNetModel.predict`` and the stitchers accept.
    Patches are decoded lazily, file by file."""

    def __init__(self, file_list, features, kernel_shape, kernel_buffer, spec_fn, one_hot, derived):
        self.file_list, self.features = sorted(file_list), list(features)
        self.kernel_shape, self.kernel_buffer = list(kernel_shape), list(kernel_buffer)
        self._spec_fn, self.one_hot, self.derived = spec_fn, one_hot, derived

    def __iter__(self):
        from . import gee_io
        for bands, extra, hot in gee_io.iter_patches(self.file_list, self.features, self.kernel_shape,
                                                     self.kernel_buffer, self.one_hot, self.derived):
            nb = bands.shape[-1]
            stacked = np.concatenate([bands] + extra + hot, axis=-1) if (extra or hot) else bands
            yield NormalizedTensor(stacked[None], self._spec_fn(nb, stacked.shape[-1]))


def make_pred_dataset(file_list, features, kernel_shape=[256, 256], kernel_buffer=[128, 128], axes=[2], splits=None,
                      moments=None, one_hot=None, **kwargs):
    """``utils/prediction_tools.py:159-226``: GZIP TFRecord patch files -> per-patch HWC stack of ``features``
    -> ``rescale_tensor(bands, axes, moments, splits)`` -> derived bands (``**kwargs`` callables) and one-hot
    planes appended un-normalised -> batches of one.  The rescale runs inside the GPU gather kernel (the
    dataset carries the normaliser spec); appended planes need ``moments`` so they can pass through as
    per-band identity."""
    from .processing import rescale_spec

    def spec_fn(nbands, ntotal):
        spec = rescale_spec(nbands, axes, 1e-8, moments, splits)
        if ntotal == nbands:
            return spec
        if spec.mode != _lib.SCV_NORM_PER_BAND:
            raise NotImplementedError('derived / one-hot bands next to a data-dependent rescale (no moments=) are '
                                      'not implemented on the GPU path')
        extra = ntotal - nbands
        return NormSpec(_lib.SCV_NORM_PER_BAND, np.concatenate([spec.sub, np.zeros(extra, np.float32)]),
                        np.concatenate([spec.div, np.ones(extra, np.float32)]), spec.eps)

    return PatchDataset(list(file_list), features, kernel_shape, kernel_buffer, spec_fn, one_hot, dict(kwargs))


def write_tfrecord_predictions(predictions, pred_path, out_image_base, kernel_shape=[256, 256], kernel_buffer=[128, 128]):
    """``utils/prediction_tools.py:375-445``: crop every prediction and write ``{pred_path}/{out_image_base}.tfrecords``
    (uncompressed; one Example per patch, FloatList features ``b1..bC``).  Returns the file name."""
    import os
    from . import gee_io
    return gee_io.write_prediction_tfrecords(predictions, os.path.join(pred_path, f'{out_image_base}.tfrecords'),
                                             kernel_shape, kernel_buffer)


def write_geotiff_prediction(image, jsonFile, aoi):
    """``utils/prediction_tools.py:447-472``: ``{aoi}.tif`` with the mixer's affine transform and CRS."""
    from . import gee_io
    mixer = _load_mixer(jsonFile)
    proj = mixer.get('projection', {})
    return gee_io.write_geotiff(f'{aoi}.tif', image, proj.get('affine', {}).get('doubleMatrix'), proj.get('crs'))


def write_geotiff_predictions(imageDataset, model, jsonFile, outImgBase, outImgPath, kernel_buffer=[128, 128], norm=None):
    """``utils/prediction_tools.py:475-536``: predict every patch, crop, place (channel 0, float32) and write
    ``{outImgPath}/{outImgBase}.tif``.  Returns the file name."""
    from . import gee_io
    out_array, transform, crs = geotiff_predictions(imageDataset, model, jsonFile, kernel_buffer, norm=norm)
    return gee_io.write_geotiff(gee_io.geotiff_path(outImgPath, outImgBase), out_array, transform, crs)


def predict_overlap_chunks(chw, m, chunk=256, depth=64, norm=None):
    """Geometry T3 -- ``predict_pc_dask`` / ``run_dask`` (``utils/prediction_tools.py:818-829``,
    ``map_overlap(depth=(0,64,64), boundary=0)``) + ``predict_chunk`` (``utils/model_tools.py:1295-1300``):
    every ``chunk``^2 block of the (C,H,W) raster is predicted with a ``depth`` halo of neighbour data,
    zeros beyond the raster edge, and the halo is trimmed.  H and W must be multiples of ``chunk``
    (``trim_dataArray``, ``utils/pc_tools.py:109-129``).  Returns (H, W) of probability channel 0.

    The reference normalises the raster first (``normalize_dataArray``, ``:757-758`` / ``:809``) and dask then pads
    the NORMALISED data with exact zeros, so with ``norm=`` (or a NormalizedTensor) the halo beyond the raster
    edge is zero AFTER normalisation: the gather kernel gets the valid window and emits zeros outside it.
    Per-tile statistics would see the padding and are rejected."""
    m = _model(m)
    chw, norm = UNetModel._split_norm(chw, norm)
    if norm is not None and norm.mode in TILE_STAT_MODES:
        raise ValueError('per-tile statistics cannot be combined with zero-padded overlap chunks')
    chw = np.asarray(chw)
    Cc, H, W = chw.shape
    if H % chunk or W % chunk:
        raise ValueError('trim the raster to a multiple of the chunk size first (trim_dataArray)')
    hwc = np.zeros((H + 2 * depth + chunk, W + 2 * depth + chunk, Cc), dtype=chw.dtype)
    hwc[depth:depth + H, depth:depth + W] = np.moveaxis(chw, 0, -1)
    # a zero-padded mosaic whose chip grid (buff = 2*depth, kernel = chunk) covers exactly [0,H)x[0,W)
    prob, _ = m.predict_mosaic(hwc, buff=2 * depth, kernel=chunk, norm=norm, want_mask=False,
                               valid=(depth, depth + H, depth, depth + W))
    return np.array(prob[depth:depth + H, depth:depth + W])


def _extract_debug(hwc, norm, device=0):
    """K1 alone on one (H,W,C) image (tests / NormalizedTensor materialisation): the bf16-rounded network
    input, widened to float32.  A square image is one chip; other shapes are cut into gcd(H, W) squares, which
    is exact for every normaliser except the per-tile statistics (those need the whole image as one tile)."""
    lib = _lib.load_library()
    a, dt = _lib.as_input(hwc)
    H, W, Cc = a.shape
    side = H if H == W else int(np.gcd(H, W))
    if side != H and norm is not None and norm.mode in TILE_STAT_MODES:
        raise ValueError('per-tile statistics can only be materialised for square images')
    t = _lib.Tiling(side, 0)
    cn = (norm or NormSpec()).to_c(Cc)
    idx = np.array([[y, x] for y in range(0, H, side) for x in range(0, W, side)], dtype=np.int32)
    cpad = C.c_int()
    out = np.empty((len(idx), side, side, _lib.SCV_MAX_BANDS), dtype=np.float32)
    _lib.check(lib.scv_debug_extract(device, _lib.ptr(a), dt, H, W, Cc, C.byref(t), C.byref(cn), _lib.ptr(idx), len(idx),
                                     _lib.ptr(out), C.byref(cpad)))
    tiles = out.reshape(-1)[:len(idx) * side * side * cpad.value].reshape(len(idx), side, side, cpad.value)[..., :Cc]
    res = np.empty((H, W, Cc), dtype=np.float32)
    for (y, x), tl in zip(idx, tiles):
        res[y:y + side, x:x + side] = tl
    return res
