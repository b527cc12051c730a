"""Oracle (test infrastructure): tile index / crop / stitch arithmetic.

Numpy restatement of the reference's integer work on the predict path.  Every
function cites the reference lines it follows (paths relative to the reference
root).  Pinned against the real reference source by
``tests/golden/make_golden.py`` (see ``oracle/__init__.py``).
"""
from __future__ import annotations

import numpy as np


def generate_chip_indices(shape, buff=128, kernel=256):
    """``utils/prediction_tools.py:87-109``.

    ``shape`` is ``arr.shape`` = (H, W, C).  Returns the row-major list of
    (y, x) upper-left corners of the *kept* ``kernel``-sized cores.  The chip
    read for index (y, x) spans ``[y-buff//2, y+kernel+buff//2)``.
    Note ``range(buff//2, H - (buff+kernel), kernel)``: the stop is exclusive,
    so a raster of exactly ``buff+kernel`` rows yields no chips.
    """
    H, W, _ = shape
    side = buff + kernel
    x_buff = y_buff = buff // 2
    y_indices = list(range(y_buff, H - side, kernel))
    x_indices = list(range(x_buff, W - side, kernel))
    return [(y, x) for y in y_indices for x in x_indices]


def extract_chips(arr, buff=128, kernel=256, legacy_xy_swap=True):
    """``utils/prediction_tools.py:111-131``.

    As committed the loop unpacks the (y, x) tuples as ``for x, y in ...``
    (``:127``), i.e. the chip for index (y, x) is cut at the transposed
    position.  ``legacy_xy_swap=True`` replicates that; ``False`` is the
    evident intent.
    """
    x_buff = y_buff = buff // 2
    chips = []
    for a, b in generate_chip_indices(arr.shape, buff, kernel):
        if legacy_xy_swap:
            x, y = a, b
        else:
            y, x = a, b
        chips.append(arr[y - y_buff:y + kernel + y_buff, x - x_buff:x + kernel + x_buff, :])
    return chips


def predict_chips(arr, chip_indices, template, predict_fn, kernel=256, buff=128):
    """``utils/prediction_tools.py:133-156`` with ``m.predict`` abstracted.

    ``predict_fn(batch)`` takes the (1, side, side, C) array the reference
    builds with ``np.array([chip])`` (``:152``) and returns (1, side, side, k).
    ``template`` is mutated in place (``+=``, ``:154``) and returned.
    """
    y_buff = x_buff = buff // 2
    if len(chip_indices) >= 1:
        for y, x in chip_indices:
            chip = arr[y - y_buff:y + kernel + y_buff, x - x_buff:x + kernel + x_buff, :]
            preds = predict_fn(np.array([chip]))
            template[y:y + kernel, x:x + kernel] += preds[0, y_buff:(kernel + y_buff), x_buff:(kernel + x_buff), 0]
    return template


def _assemble_rows(patches, cols):
    """Row-major patch assembly of ``utils/prediction_tools.py:269-291`` /
    ``:351-373`` (``np.append`` along axis 1 then axis 0).  The reference's
    ``x % cols == 1`` test never fires for ``cols == 1`` (Appendix A6); the
    oracle implements the evident placement for that case.
    """
    rows_out = None
    row = None
    x = 1
    for patch in patches:
        if cols == 1 or x % cols == 1:
            row = patch
        else:
            row = np.append(row, patch, axis=1)
        if x % cols == 0:
            if x <= cols:
                rows_out = row
            else:
                rows_out = np.append(rows_out, row, axis=0)
        x += 1
    return rows_out


def crop_window(kernel_shape=(256, 256), kernel_buffer=(128, 128)):
    """The crop slice pair of ``utils/prediction_tools.py:258-261, :340-343``.

    Note the reference mixes x/y names (``x_size = kernel_shape[0]+y_buffer``)
    -- harmless for square kernels; restated literally.
    Returns ((y0, y1), (x0, x1)) as used in ``prediction[y0:y1, x0:x1]``.
    """
    x_buffer = int(kernel_buffer[0] / 2)
    y_buffer = int(kernel_buffer[1] / 2)
    x_size = kernel_shape[0] + y_buffer
    y_size = kernel_shape[1] + x_buffer
    return (y_buffer, y_size), (x_buffer, x_size)


def make_array_predictions(predictions, mixer, kernel_shape=(256, 256), kernel_buffer=(128, 128)):
    """``utils/prediction_tools.py:293-373`` after ``model.predict``.

    ``predictions`` is the (N, h, w, k) array (or the ``[probs, classes]`` list,
    concatenated on axis 3 as at ``:336-338``; a 3-D ``classes`` gets a trailing
    axis first -- Appendix A7).  Returns (rows*kh, cols*kw, k).
    """
    if isinstance(predictions, list):
        parts = [p if p.ndim == 4 else p[..., None] for p in predictions]
        predictions = np.concatenate([p.astype(np.float32) for p in parts], axis=3)
    cols = mixer['patchesPerRow']
    (y0, y1), (x0, x1) = crop_window(kernel_shape, kernel_buffer)
    patches = [prediction[y0:y1, x0:x1, :] for prediction in predictions]
    return _assemble_rows(patches, cols)


def callback_predictions(predictions, mixer, kernel_shape=(256, 256), kernel_buffer=(128, 128)):
    """``utils/prediction_tools.py:245-291`` after ``model.predict``: keeps
    channel 1 of the probabilities (``:267``); a list output keeps element 0
    (``:254-256``).  Returns (rows*kh, cols*kw)."""
    if isinstance(predictions, list):
        predictions = predictions[0]
    cols = mixer['patchesPerRow']
    (y0, y1), (x0, x1) = crop_window(kernel_shape, kernel_buffer)
    patches = [prediction[y0:y1, x0:x1, 1] for prediction in predictions]
    return _assemble_rows(patches, cols)


def geotiff_stitch(predictions, mixer, kernel_buffer=(128, 128)):
    """Stitch part of ``utils/prediction_tools.py:475-520``: preallocated
    (rows*kh, cols*kw, 1) float32, patch i at the i-th row-major (y, x) grid
    index, channel 0 kept (``:520``)."""
    ppr = mixer['patchesPerRow']
    tp = mixer['totalPatches']
    rows = int(tp / ppr)
    kernel_shape = mixer['patchDimensions']
    H = rows * kernel_shape[0]
    W = ppr * kernel_shape[1]
    indices = [(y, x) for y in range(0, H, kernel_shape[0]) for x in range(0, W, kernel_shape[1])]
    out_array = np.zeros((H, W, 1), dtype=np.float32)
    x_buffer = int(kernel_buffer[0] / 2)
    y_buffer = int(kernel_buffer[1] / 2)
    x_size = x_buffer + kernel_shape[1]
    y_size = y_buffer + kernel_shape[0]
    if isinstance(predictions, list):
        parts = [p if p.ndim == 4 else p[..., None] for p in predictions]
        predictions = np.concatenate([p.astype(np.float32) for p in parts], axis=3)
    for i, (y, x) in enumerate(indices):
        out_array[y:y + kernel_shape[0], x:x + kernel_shape[1], 0] += predictions[i, y_buffer:y_size, x_buffer:x_size, 0]
    return out_array


def trim_extent(n, size):
    """``utils/pc_tools.py:109-129``: length kept after trimming the remainder."""
    remainder = n % size
    return n - remainder if remainder else n


def predict_overlap_chunks(chw, predict_fn, chunk=256, depth=64):
    """Geometry T3: ``utils/prediction_tools.py:818-829`` (``map_overlap(depth=
    (0,64,64), boundary=0)``) + ``utils/model_tools.py:1295-1300``.

    ``chw`` is (C, H, W) with H, W already trimmed to multiples of ``chunk``.
    Every chunk is padded with ``depth`` px of neighbour data, zeros beyond the
    raster edge; ``predict_fn`` receives (1, chunk+2*depth, chunk+2*depth, C) and
    returns (1, h, w, k); ``np.squeeze(pred[0])`` is kept and the halo trimmed.
    Returns (H, W) or (H, W, k).
    """
    C, H, W = chw.shape
    padded = np.zeros((C, H + 2 * depth, W + 2 * depth), dtype=chw.dtype)
    padded[:, depth:depth + H, depth:depth + W] = chw
    out = None
    for y in range(0, H, chunk):
        for x in range(0, W, chunk):
            block = padded[:, y:y + chunk + 2 * depth, x:x + chunk + 2 * depth]
            hwc = np.moveaxis(block, 0, -1)
            pred = predict_fn(np.expand_dims(hwc, axis=0))
            logits = np.squeeze(pred[0])
            core = logits[depth:depth + chunk, depth:depth + chunk]
            if out is None:
                out = np.zeros((H, W) + core.shape[2:], dtype=core.dtype)
            out[y:y + chunk, x:x + chunk] = core
    return out


def raster_generate_chip_indices(H, W, buff=128, kernel=256):
    """utils/raster_tools.py:23-46: per-side buffer, inclusive range end."""
    ys = list(range(buff, H - (kernel + buff) + 1, kernel))
    xs = list(range(buff, W - (kernel + buff) + 1, kernel))
    return [(y, x) for y in ys for x in xs]
