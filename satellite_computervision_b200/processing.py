"""Normalisers of the predict path, with the reference's signatures
(``utils/processing.py:225-322``, ``utils/pc_tools.py:90-107``), fused into the extract kernel (K1).

The reference applies ``rescale_tensor`` / ``normalize_tensor`` to every patch in
a ``tf.data`` map before ``model.predict``.  Here they return a lazy
:class:`NormalizedTensor` -- the raw array plus a normaliser spec -- that
``UNetModel.predict``, ``prediction_tools.predict_chips`` etc. accept directly:
the arithmetic then happens on the GPU inside the gather kernel (fp32, subtract
then IEEE divide exactly as the reference writes it, one rounding to bf16).
Nothing is computed on the CPU.  ``np.asarray(t)`` / arithmetic on the lazy tensor
materialise it through the same kernel (the bf16-rounded values the network sees).
"""
from __future__ import annotations

import numpy as np

from . import _lib

_DATA_MODES = {
    ('rescale', (2,)): _lib.SCV_NORM_PIXEL_MINMAX,
    ('rescale', (0, 1)): _lib.SCV_NORM_TILE_MINMAX,
    ('rescale', (0, 1, 2)): _lib.SCV_NORM_TILE_GLOBAL_MINMAX,
    ('normalize', (2,)): _lib.SCV_NORM_PIXEL_ZSCORE,
    ('normalize', (0, 1)): _lib.SCV_NORM_TILE_ZSCORE,
    ('normalize', (0, 1, 2)): _lib.SCV_NORM_TILE_GLOBAL_ZSCORE,
}
TILE_STAT_MODES = (_lib.SCV_NORM_TILE_ZSCORE, _lib.SCV_NORM_TILE_MINMAX, _lib.SCV_NORM_TILE_GLOBAL_MINMAX,
                   _lib.SCV_NORM_TILE_GLOBAL_ZSCORE)


class NormSpec:
    """Normaliser fused into K1 (``scv_norm`` of include/scv.h).  ``groups``: channel-group sizes of the
    data-derived modes (``splits=``); channels beyond ``sum(groups)`` pass through."""

    def __init__(self, mode=_lib.SCV_NORM_NONE, sub=None, div=None, eps=1e-8, groups=None):
        self.mode = mode
        self.sub = None if sub is None else np.asarray(sub, dtype=np.float32)
        self.div = None if div is None else np.asarray(div, dtype=np.float32)
        self.eps = float(eps)
        self.groups = None if not groups else [int(g) for g in groups]

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
            if self.groups:
                if len(self.groups) > _lib.SCV_MAX_BANDS or sum(self.groups) > nbands:
                    raise ValueError(f'channel groups {self.groups} do not fit {nbands} bands')
                n.ngroups = len(self.groups)
                for i, g in enumerate(self.groups):
                    n.group_size[i] = g
        return n

    def with_passthrough(self, nbands, ntotal):
        """The same normaliser over the first ``nbands`` channels of a wider stack whose extra planes (derived /
        one-hot bands, ``utils/prediction_tools.py:198-215``) are appended un-normalised."""
        if ntotal == nbands or self.mode == _lib.SCV_NORM_NONE:
            return self
        extra = ntotal - nbands
        if self.mode == _lib.SCV_NORM_PER_BAND:
            return NormSpec(self.mode, np.concatenate([self.sub, np.zeros(extra, np.float32)]),
                            np.concatenate([self.div, np.ones(extra, np.float32)]), self.eps)
        return NormSpec(self.mode, eps=self.eps, groups=self.groups or [nbands])

    def __repr__(self):
        return f'NormSpec(mode={self.mode}, sub={self.sub}, div={self.div}, eps={self.eps}, groups={self.groups})'


class NormalizedTensor:
    """Lazy result of ``rescale_tensor`` / ``normalize_tensor``: ``raw`` + ``norm``.  Passing it to the
    predict functions keeps the normalisation inside the gather kernel; anything else (``np.asarray``,
    arithmetic, indexing) materialises it on the GPU first."""

    __array_priority__ = 100

    def __init__(self, raw, norm):
        self.raw = np.asarray(raw)
        self.norm = norm

    @property
    def shape(self):
        return self.raw.shape

    @property
    def ndim(self):
        return self.raw.ndim

    @property
    def dtype(self):
        return np.dtype(np.float32)

    def numpy(self, device=0):
        """Materialise through the K1 kernel: the bf16-rounded values the network sees (fp32 array), for one
        (H, W, C) image or a stack (N, H, W, C); per-tile statistics are per image."""
        from .prediction_tools import _extract_debug
        if self.raw.ndim == 3:
            return _extract_debug(self.raw, self.norm, device)
        if self.raw.ndim == 4:
            return np.stack([_extract_debug(t, self.norm, device) for t in self.raw])
        raise ValueError('numpy() materialises (H, W, C) or (N, H, W, C) data')

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a if dtype is None else a.astype(dtype, copy=False)

    def __getitem__(self, idx):
        return self.numpy()[idx]

    def __len__(self):
        return len(self.raw)

    def _bin(self, other, op, swap=False):
        a, b = self.numpy(), (other.numpy() if isinstance(other, NormalizedTensor) else other)
        return op(b, a) if swap else op(a, b)

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return self._bin(o, np.add, True)
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return self._bin(o, np.subtract, True)
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __rmul__(self, o): return self._bin(o, np.multiply, True)
    def __truediv__(self, o): return self._bin(o, np.divide)
    def __rtruediv__(self, o): return self._bin(o, np.divide, True)
    def __neg__(self): return -self.numpy()


def _expand_moments(moments, nbands, splits, what, passthrough_rest):
    """Per-channel (a, b) constants the way numpy broadcasting applies ``moments`` in the reference: the same
    list to every split (``:314-318``, ``:267-275``), a length-1 list to every channel."""
    m = [(float(a), float(b)) for a, b in moments]
    if splits:
        out = []
        for s in splits:
            if len(m) not in (1, s):
                raise ValueError(f'{what}: moments of length {len(m)} do not broadcast against a split of {s} channels')
            out += m * s if len(m) == 1 else m
        covered = sum(splits)
        if covered > nbands or (covered != nbands and not passthrough_rest):
            raise ValueError(f'{what}: split sizes {list(splits)} must sum to the channel count {nbands}')
        return out, covered
    if len(m) == 1:
        m = m * nbands
    if len(m) != nbands:
        raise ValueError(f'{what}: {len(m)} moments for {nbands} bands')
    return m, nbands


def _axes_mode(kind, axes):
    axes = tuple(sorted(int(a) % 3 for a in axes))
    try:
        return _DATA_MODES[(kind, axes)]
    except KeyError:
        raise NotImplementedError(f'{kind}_tensor(axes={list(axes)}) has no GPU form') from None


def rescale_spec(nbands, axes=(2,), epsilon=1e-8, moments=None, splits=None):
    """NormSpec equivalent to ``rescale_tensor`` (``utils/processing.py:281-322``)."""
    if moments:
        m, _ = _expand_moments(moments, nbands, splits, 'rescale_tensor', passthrough_rest=False)
        mn = np.array([t[0] for t in m], dtype=np.float32)
        mx = np.array([t[1] for t in m], dtype=np.float32)
        # (img - minimum)/((maximum - minimum) + epsilon), float32 arithmetic (:304-311)
        den = (mx - mn) + epsilon
        return NormSpec(_lib.SCV_NORM_PER_BAND, mn, den.astype(np.float32), epsilon)
    if splits and sum(splits) != nbands:
        raise ValueError('rescale_tensor: tf.split sizes must sum to the channel count')
    return NormSpec(_axes_mode('rescale', axes), eps=epsilon, groups=list(splits) if splits else None)


def normalize_spec(nbands, axes=(2,), epsilon=1e-8, moments=None, splits=None):
    """NormSpec equivalent to ``normalize_tensor`` (``utils/processing.py:225-279``)."""
    if moments:
        m, covered = _expand_moments(moments, nbands, splits, 'normalize_tensor', passthrough_rest=True)
        mean = np.array([t[0] for t in m], dtype=np.float32)
        var = np.array([t[1] for t in m], dtype=np.float32)
        den = np.sqrt(var + epsilon).astype(np.float32)  # tf.sqrt(variance + epsilon), float32 (:262)
        if covered < nbands:  # channels beyond sum(splits) pass through (:269-274)
            mean = np.concatenate([mean, np.zeros(nbands - covered, np.float32)])
            den = np.concatenate([den, np.ones(nbands - covered, np.float32)])
        return NormSpec(_lib.SCV_NORM_PER_BAND, mean, den, epsilon)
    if splits and sum(splits) > nbands:
        raise ValueError('normalize_tensor: split sizes exceed the channel count')
    return NormSpec(_axes_mode('normalize', axes), eps=epsilon, groups=list(splits) if splits else None)


def scalar_spec(nbands, rescale_val):
    """``x / rescale_val`` of ``UNETDataGenerator`` prediction mode (``utils/processing.py:551-552,
    :601, :613``: Sentinel-2 10000.0, NAIP 255.0)."""
    return NormSpec(_lib.SCV_NORM_PER_BAND, np.zeros(nbands, np.float32), np.full(nbands, rescale_val, np.float32))


def band_zscore_spec(epsilon=1e-6):
    """``pc_tools.normalize_dataArray(da, 'band')`` (``utils/pc_tools.py:90-107``; applied to the whole mosaic
    at ``utils/prediction_tools.py:757-758``): per pixel across bands, NaN-skipping, ``(x - mean) / (sd + 1e-6)``."""
    return NormSpec(_lib.SCV_NORM_PIXEL_ZSCORE_SD, eps=epsilon)


def rescale_tensor(img, axes=[2], epsilon=1e-8, moments=None, splits=None):
    """``utils/processing.py:281`` -- same signature; returns a lazy NormalizedTensor."""
    img = np.asarray(img)
    return NormalizedTensor(img, rescale_spec(img.shape[-1], axes, epsilon, moments, splits))


def normalize_tensor(x, axes=[2], epsilon=1e-8, moments=None, splits=None):
    """``utils/processing.py:225`` -- same signature; returns a lazy NormalizedTensor."""
    x = np.asarray(x)
    return NormalizedTensor(x, normalize_spec(x.shape[-1], axes, epsilon, moments, splits))


def normalize_dataArray(da, dim='band'):
    """``utils/pc_tools.py:90-107`` for the layout of the mosaic path: ``da`` is (band, y, x) like the xarray
    median composite (``utils/prediction_tools.py:750-758``) or already (y, x, band).  Returns a lazy
    NormalizedTensor over the (y, x, band) view, ready for ``predict_chips``."""
    a = np.asarray(da)
    if dim in ('band', 0):
        a = np.moveaxis(a, 0, -1)
    elif dim not in (-1, 2, 'last'):
        raise NotImplementedError("only the band dimension is normalised on the predict path (dim='band')")
    return NormalizedTensor(a, band_zscore_spec())
