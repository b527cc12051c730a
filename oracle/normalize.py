"""Oracle (test infrastructure): the normalisers on the predict path.

Numpy restatement; TensorFlow reductions are replaced by their numpy
equivalents (``tf.nn.moments`` = mean and *population* variance;
``tf.math.reduce_min/max`` with keepdims).  Arithmetic stays in the dtype of the
input (float32 on the reference path, ``make_pred_dataset`` parses float32).
"""
from __future__ import annotations

import numpy as np


def _split_apply(img, splits, fn, passthrough_rest):
    """Channel-group handling shared by both functions."""
    if passthrough_rest:
        # normalize_tensor: first sum(splits) channels are normalised per group,
        # the rest passed through (utils/processing.py:267-275)
        split_len = sum(splits)
        to_norm = img[:, :, 0:split_len]
        dont_norm = img[:, :, split_len:]
        bounds = np.cumsum(splits)[:-1]
        parts = [fn(t) for t in np.split(to_norm, bounds, axis=2)]
        parts.append(dont_norm)
        return np.concatenate(parts, axis=2)
    # rescale_tensor: tf.split(img, splits, axis=2) -- sizes must sum to C
    # (utils/processing.py:314-318)
    assert sum(splits) == img.shape[2], 'tf.split sizes must sum to the channel count'
    bounds = np.cumsum(splits)[:-1]
    return np.concatenate([fn(t) for t in np.split(img, bounds, axis=2)], axis=2)


def rescale_tensor(img, axes=(2,), epsilon=1e-8, moments=None, splits=None):
    """``utils/processing.py:281-322``: ``(img - min) / ((max - min) + eps)``.

    ``moments`` = list of (min_c, max_c) -> float32 per-band constants
    (``:304-305``); otherwise min/max reduced over ``axes`` with keepdims
    (``:307-308``).  Default ``axes=[2]`` rescales each *pixel* across bands
    (Appendix A11).
    """
    img = np.asarray(img)

    def rescale(t):
        if moments:
            minimum = np.array([tpl[0] for tpl in moments], dtype='float32')
            maximum = np.array([tpl[1] for tpl in moments], dtype='float32')
        else:
            minimum = np.min(t, axis=tuple(axes), keepdims=True)
            maximum = np.max(t, axis=tuple(axes), keepdims=True)
        return (t - minimum) / ((maximum - minimum) + epsilon)

    if splits:
        return _split_apply(img, list(splits), rescale, passthrough_rest=False)
    return rescale(img)


def normalize_tensor(x, axes=(2,), epsilon=1e-8, moments=None, splits=None):
    """``utils/processing.py:225-279``: ``(x - mean) / sqrt(var + eps)``.

    ``moments`` = list of (mean_c, var_c) float32 (``:252-254``), else population
    mean/variance over ``axes`` with keepdims (``:257``).  The solar notebook's
    predict path is ``axes=[0, 1]`` on the 384x384 buffered tile
    (``notebooks/UNET_G4G_2019_solar.ipynb:808-820, :1541``).
    """
    x = np.asarray(x)

    def normalize(t):
        if moments:
            mean = np.array([tpl[0] for tpl in moments], dtype='float32')
            variance = np.array([tpl[1] for tpl in moments], dtype='float32')
        else:
            mean = np.mean(t, axis=tuple(axes), keepdims=True, dtype=t.dtype)
            variance = np.mean(np.square(t - mean), axis=tuple(axes), keepdims=True, dtype=t.dtype)
        return (t - mean) / np.sqrt(variance + np.asarray(epsilon, dtype=t.dtype))

    if splits:
        return _split_apply(x, list(splits), normalize, passthrough_rest=True)
    return normalize(x)


def scalar_rescale(x, rescale_val):
    """Prediction-mode scalar rescale of ``UNETDataGenerator``
    (``utils/processing.py:551-552, :601, :613``): ``x / rescale_val``
    (Sentinel-2 10000.0, NAIP 255.0)."""
    return np.asarray(x) / rescale_val


def normalize_data_array(chw, axis=0):
    """``utils/pc_tools.py:90-107`` ``normalize_dataArray(da, 'band')``: per-pixel
    z-score across bands, NaN-skipping, population std, ``(x-mean)/(sd+1e-6)``."""
    mean = np.nanmean(chw, axis=axis, keepdims=True)
    sd = np.nanstd(chw, axis=axis, keepdims=True)
    return (chw - mean) / (sd + 0.000001)
