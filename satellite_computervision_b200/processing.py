"""Normalisers of the predict path, with the reference's signatures
(``utils/processing.py:225-322``), fused into the extract kernel (K1).

The reference applies ``rescale_tensor`` / ``normalize_tensor`` to every patch in
a ``tf.data`` map before ``model.predict``.  Here they return a lazy
:class:`NormalizedTensor` -- the raw array plus a normaliser spec -- that
``UNetModel.predict``, ``prediction_tools.predict_chips`` etc. accept directly:
the arithmetic then happens on the GPU inside the gather kernel (fp32, subtract
then IEEE divide exactly as the reference writes it, one rounding to bf16).
Nothing is computed on the CPU.
"""
from __future__ import annotations

import numpy as np

from . import _lib


class NormSpec:
    """Normaliser fused into K1 (``scv_norm`` of include/scv.h)."""

    def __init__(self, mode=_lib.SCV_NORM_NONE, sub=None, div=None, eps=1e-8):
        self.mode = mode
        self.sub = None if sub is None else np.asarray(sub, dtype=np.float32)
        self.div = None if div is None else np.asarray(div, dtype=np.float32)
        self.eps = float(eps)

    def to_c(self, nbands):
        n = _lib.Norm()
        n.mode = self.mode
        n.nbands = nbands
        if self.mode == _lib.SCV_NORM_PER_BAND:
            if len(self.sub) != nbands or len(self.div) != nbands:
                raise ValueError(f'normaliser has {len(self.sub)} bands, input has {nbands}')
            for c in range(nbands):
                n.sub[c] = float(self.sub[c])
                n.div[c] = float(self.div[c])
        elif self.mode != _lib.SCV_NORM_NONE:
            n.div[0] = np.float32(self.eps)
        return n

    def __repr__(self):
        return f'NormSpec(mode={self.mode}, sub={self.sub}, div={self.div}, eps={self.eps})'


class NormalizedTensor:
    """Lazy result of ``rescale_tensor`` / ``normalize_tensor``: ``raw`` + ``norm``."""

    def __init__(self, raw, norm):
        self.raw = np.asarray(raw)
        self.norm = norm

    @property
    def shape(self):
        return self.raw.shape

    def numpy(self, device=0):
        """Materialise through the K1 kernel: the bf16-rounded values the network sees (fp32 array)."""
        from .prediction_tools import _extract_debug
        hwc = self.raw if self.raw.ndim == 3 else None
        if hwc is None:
            raise ValueError('numpy() materialises a single (H, W, C) image')
        return _extract_debug(hwc, self.norm, device)


def _tile_moments(moments, nbands, splits, what):
    m = [(float(a), float(b)) for a, b in moments]
    if splits:
        # the reference applies the same `moments` list to every channel group (:314-318, :267-275)
        for s in splits:
            if s != len(m):
                raise ValueError(f'{what}: every split must have len(moments)={len(m)} channels, got {s}')
        covered = sum(splits)
        reps = len(splits)
        m = m * reps
        return m, covered
    if len(m) != nbands:
        raise ValueError(f'{what}: {len(m)} moments for {nbands} bands')
    return m, nbands


def rescale_spec(nbands, axes=(2,), epsilon=1e-8, moments=None, splits=None):
    """NormSpec equivalent to ``rescale_tensor`` (``utils/processing.py:281-322``)."""
    if moments:
        m, covered = _tile_moments(moments, nbands, splits, 'rescale_tensor')
        if covered != nbands:
            raise ValueError('rescale_tensor: tf.split sizes must sum to the channel count')
        mn = np.array([t[0] for t in m], dtype=np.float32)
        mx = np.array([t[1] for t in m], dtype=np.float32)
        # (img - minimum)/((maximum - minimum) + epsilon), float32 arithmetic (:304-311)
        den = (mx - mn) + epsilon
        return NormSpec(_lib.SCV_NORM_PER_BAND, mn, den.astype(np.float32), epsilon)
    if splits:
        raise NotImplementedError('rescale_tensor(splits=) with data-derived min/max is not implemented on the GPU path')
    axes = tuple(sorted(int(a) for a in axes))
    if axes == (2,):
        return NormSpec(_lib.SCV_NORM_PIXEL_MINMAX, eps=epsilon)
    if axes == (0, 1):
        return NormSpec(_lib.SCV_NORM_TILE_MINMAX, eps=epsilon)
    raise NotImplementedError(f'rescale_tensor(axes={list(axes)}) is not implemented on the GPU path')


def normalize_spec(nbands, axes=(2,), epsilon=1e-8, moments=None, splits=None):
    """NormSpec equivalent to ``normalize_tensor`` (``utils/processing.py:225-279``)."""
    if moments:
        m, covered = _tile_moments(moments, nbands, splits, 'normalize_tensor')
        mean = np.array([t[0] for t in m], dtype=np.float32)
        var = np.array([t[1] for t in m], dtype=np.float32)
        den = np.sqrt(var + epsilon).astype(np.float32)  # tf.sqrt(variance + epsilon), float32 (:262)
        if covered < nbands:  # channels beyond sum(splits) pass through (:269-274)
            mean = np.concatenate([mean, np.zeros(nbands - covered, np.float32)])
            den = np.concatenate([den, np.ones(nbands - covered, np.float32)])
        return NormSpec(_lib.SCV_NORM_PER_BAND, mean, den, epsilon)
    if splits:
        raise NotImplementedError('normalize_tensor(splits=) with data-derived moments is not implemented on the GPU path')
    axes = tuple(sorted(int(a) for a in axes))
    if axes == (2,):
        return NormSpec(_lib.SCV_NORM_PIXEL_ZSCORE, eps=epsilon)
    if axes == (0, 1):
        return NormSpec(_lib.SCV_NORM_TILE_ZSCORE, eps=epsilon)
    raise NotImplementedError(f'normalize_tensor(axes={list(axes)}) is not implemented on the GPU path')


def scalar_spec(nbands, rescale_val):
    """``x / rescale_val`` of ``UNETDataGenerator`` prediction mode (``utils/processing.py:551-552,
    :601, :613``: Sentinel-2 10000.0, NAIP 255.0)."""
    return NormSpec(_lib.SCV_NORM_PER_BAND, np.zeros(nbands, np.float32), np.full(nbands, rescale_val, np.float32))


def rescale_tensor(img, axes=[2], epsilon=1e-8, moments=None, splits=None):
    """``utils/processing.py:281`` -- same signature; returns a lazy NormalizedTensor."""
    img = np.asarray(img)
    return NormalizedTensor(img, rescale_spec(img.shape[-1], axes, epsilon, moments, splits))


def normalize_tensor(x, axes=[2], epsilon=1e-8, moments=None, splits=None):
    """``utils/processing.py:225`` -- same signature; returns a lazy NormalizedTensor."""
    x = np.asarray(x)
    return NormalizedTensor(x, normalize_spec(x.shape[-1], axes, epsilon, moments, splits))
