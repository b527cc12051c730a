"""U-Net builders with the reference's call surface (``utils/model_tools.py:394-454``),
returning a :class:`UNetModel` that duck-types the ``keras.Model`` methods the
predict path uses (``predict`` / ``set_weights`` / ``get_weights`` /
``load_weights`` / ``count_params``) on top of the sm_100a engine (libscv.so).

Two network variants exist in the reference under one name:

* ``double_conv=False`` -- ``get_unet_model`` *as written*: ``conv_block.call``
  applies ``cba1`` twice and never builds ``cba2`` (``model_tools.py:238-239``), so
  every encoder/centre block is ONE conv-BN-ReLU; softmax head + argmax.
* ``double_conv=True`` -- the classic block of the notebooks' ``get_model``
  (``notebooks/UNET_G4G_2019_solar.ipynb:1162-1213``) and the intent of
  ``binary_unet`` (``model_tools.py:417-454``); sigmoid head, strict ``> thr``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from .processing import NormalizedTensor, NormSpec

DEFAULT_FILTERS = [32, 64, 128, 256, 512]


class UNetModel:
    """keras.Model stand-in for the predict path, backed by the CUDA engine.

    ``outputs``: 'probs' -> ``predict`` returns one array (N,H,W,k) like the
    notebooks' single-output model (what ``predict_chips`` needs,
    ``prediction_tools.py:152-154``); 'both' -> ``[probs, classes]`` like
    ``get_unet_model`` (``model_tools.py:407``): classes int32 (N,H,W) for the
    softmax/argmax head (``:406``), (N,H,W,1) for sigmoid/greater (``:445``).
    """

    def __init__(self, nclasses, nchannels, filters=None, double_conv=False, head='softmax', threshold=0.5,
                 bias=None, outputs='both', device=0, max_batch=64, seed=None, arch=_lib.SCV_ARCH_UNET):
        filters = list(DEFAULT_FILTERS if filters is None else filters)
        if head not in ('softmax', 'sigmoid'):
            raise ValueError("head must be 'softmax' or 'sigmoid'")
        self.nclasses, self.nchannels, self.filters = int(nclasses), int(nchannels), filters
        self.double_conv, self.head, self.threshold = bool(double_conv), head, float(threshold)
        self.outputs, self.device, self.max_batch = outputs, int(device), int(max_batch)
        self._cfg = _lib.Config()
        self._cfg.device = self.device
        self._cfg.double_conv = int(self.double_conv)
        self._cfg.nchannels = self.nchannels
        self._cfg.nclasses = self.nclasses
        self._cfg.nlevels = len(filters)
        if len(filters) > _lib.SCV_MAX_LEVELS:
            raise ValueError('too many levels')
        for i, f in enumerate(filters):
            self._cfg.filters[i] = int(f)
        self._cfg.head = _lib.SCV_HEAD_SIGMOID if head == 'sigmoid' else _lib.SCV_HEAD_SOFTMAX
        self._cfg.threshold = self.threshold
        self._cfg.max_batch = self.max_batch
        self._cfg.arch = int(arch)
        self._lib = _lib.load_library()
        n = self._lib.scv_num_weights(C.byref(self._cfg))
        if n < 0:
            _lib.check(n)
        self.weight_names, self.weight_shapes = [], []
        for i in range(n):
            nd = C.c_int()
            shp = (C.c_int64 * 4)()
            name = C.create_string_buffer(128)
            _lib.check(self._lib.scv_weight_shape(C.byref(self._cfg), i, C.byref(nd), shp, name, 128))
            self.weight_names.append(name.value.decode())
            self.weight_shapes.append(tuple(int(shp[d]) for d in range(nd.value)))
        self._engine = None
        self._weights_dirty = True
        self._weights = self._keras_default_init(seed, bias)

    # ------------------------------------------------------------------ weights
    def _keras_default_init(self, seed, head_bias):
        """Keras defaults: glorot_uniform kernels, zero biases, BN (1,0,0,1); the head bias is
        ``Constant(bias)`` when given (``model_tools.py:395-396, :405``)."""
        rng = np.random.default_rng(seed)
        out = []
        for name, shape in zip(self.weight_names, self.weight_shapes):
            leaf = name.rsplit('/', 1)[1]
            if leaf == 'kernel':
                rf = shape[0] * shape[1]
                fan_in, fan_out = rf * shape[2], rf * shape[3]
                limit = np.sqrt(6.0 / (fan_in + fan_out))
                w = rng.uniform(-limit, limit, shape)
            elif leaf in ('gamma', 'moving_variance'):
                w = np.ones(shape)
            elif leaf == 'bias' and name.startswith('head') and head_bias is not None:
                w = np.full(shape, head_bias)
            else:
                w = np.zeros(shape)
            out.append(np.ascontiguousarray(w, dtype=np.float32))
        return out

    def get_weights(self):
        return [w.copy() for w in self._weights]

    def set_weights(self, weights):
        weights = list(weights)
        if len(weights) != len(self.weight_shapes):
            raise ValueError(f'You called `set_weights(weights)` with a weight list of length {len(weights)}, '
                             f'but the model was expecting {len(self.weight_shapes)} weights.')
        new = []
        for w, shape, name in zip(weights, self.weight_shapes, self.weight_names):
            w = np.asarray(w)
            if tuple(w.shape) != shape:
                raise ValueError(f'weight {name}: shape {tuple(w.shape)} not compatible with {shape}')
            new.append(np.ascontiguousarray(w, dtype=np.float32))
        self._weights = new
        self._weights_dirty = True

    def load_weights(self, path, by_name=False, skip_mismatch=False):
        """``keras.Model.load_weights`` (``utils/model_tools.py:1162, :1200``): Keras HDF5 files (legacy
        ``.h5`` / ``.hdf5`` full-model or weights-only, Keras 3 ``.weights.h5`` / ``.keras``) through
        :mod:`keras_h5`, or an ``.npz`` written as ``np.savez(path, *model.get_weights())``.  Tensors are
        taken in ``get_weights()`` order and checked for count and shape."""
        if str(path).endswith('.npz'):
            with np.load(path) as z:
                keys = sorted(z.files, key=lambda k: int(k.split('_')[1]) if k.startswith('arr_') else 0)
                self.set_weights([z[k] for k in keys])
            return
        from . import keras_h5
        self.set_weights(keras_h5.read_weights(path, self))

    def save_weights(self, path):
        """``.npz`` of ``get_weights()``, or a legacy Keras HDF5 weights file for ``.h5`` / ``.hdf5``."""
        if str(path).endswith(('.h5', '.hdf5')):
            from . import keras_h5
            keras_h5.write_weights_h5(path, keras_h5.keras_layer_groups(self))
            return
        np.savez(path, *self._weights)

    def count_params(self):
        return int(sum(int(np.prod(s)) for s in self.weight_shapes))

    def summary(self):
        for n, s in zip(self.weight_names, self.weight_shapes):
            print(f'{n:40s} {s}')
        print('Total params:', self.count_params())

    # ------------------------------------------------------------------- engine
    def _ensure_engine(self):
        if self._engine is None:
            h = C.c_void_p()
            _lib.check(self._lib.scv_engine_create(C.byref(self._cfg), C.byref(h)))
            self._engine = h
            self._weights_dirty = True
        if self._weights_dirty:
            arr = (_lib.Tensor * len(self._weights))()
            for i, w in enumerate(self._weights):
                arr[i].data = w.ctypes.data_as(C.POINTER(C.c_float))
                arr[i].ndim = w.ndim
                for d in range(w.ndim):
                    arr[i].shape[d] = w.shape[d]
            _lib.check(self._lib.scv_engine_set_weights(self._engine, arr, len(self._weights)))
            self._weights_dirty = False
        return self._engine

    def set_option(self, key, value):
        _lib.check(self._lib.scv_set_option(self._ensure_engine(), key.encode(), int(value)))

    def times(self):
        t = _lib.Times()
        _lib.check(self._lib.scv_get_times(self._ensure_engine(), C.byref(t)))
        nl = t.n_layers
        return dict(total_ms=t.total_ms, extract_ms=t.extract_ms, network_ms=t.network_ms, stitch_ms=t.stitch_ms,
                    n_batches=t.n_batches, n_tiles=t.n_tiles, n_launches=t.n_launches,
                    layer_ms=list(t.layer_ms[:nl]), layer_flops=list(t.layer_flops[:nl]),
                    h2d_lead_ms=t.h2d_lead_ms, d2h_tail_ms=t.d2h_tail_ms)

    def close(self):
        if self._engine is not None:
            self._lib.scv_engine_destroy(self._engine)
            self._engine = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ predict
    @staticmethod
    def _split_norm(x, norm):
        if isinstance(x, NormalizedTensor):
            if norm is not None:
                raise ValueError('input is already a NormalizedTensor; do not pass norm= as well')
            return x.raw, x.norm
        return x, norm

    def predict(self, x, batch_size=None, verbose=0, steps=None, norm=None):
        """``keras.Model.predict``: ``x`` is (N,H,W,C) (any engine dtype; uint16 DN welcome with
        ``norm=``), a NormalizedTensor, or an iterable of such batches (a ``tf.data``-style
        dataset, ``prediction_tools.py:251, :333``; ``steps`` bounds the number of batches)."""
        if not isinstance(x, (np.ndarray, NormalizedTensor)):
            batches = []
            for i, b in enumerate(x):
                if steps is not None and i >= steps:
                    break
                batches.append(b)
            raws, norms = zip(*(self._split_norm(b, norm) for b in batches)) if batches else ((), ())
            if not raws:
                raise ValueError('empty dataset')
            x, norm = np.concatenate([np.asarray(r) for r in raws], axis=0), norms[0]
        x, norm = self._split_norm(x, norm)
        x = np.asarray(x)
        if x.ndim != 4:
            raise ValueError(f'expected (N, H, W, C) input, got shape {x.shape}')
        N, H, W, Cc = x.shape
        arr, dt = _lib.as_input(x)
        probs = np.empty((N, H, W, self.nclasses), dtype=np.float32)
        classes = np.empty((N, H, W), dtype=np.int32) if self.outputs == 'both' else None
        cn = (norm or NormSpec()).to_c(Cc)
        _lib.check(self._lib.scv_predict_tiles(self._ensure_engine(), _lib.ptr(arr), dt, N, H, W, Cc, C.byref(cn),
                                               _lib.ptr(probs), _lib.ptr(classes)))
        if self.outputs == 'both':
            if self.head == 'sigmoid':
                classes = classes[..., None]
            return [probs, classes]
        return probs

    def predict_mosaic(self, arr, buff=128, kernel=256, norm=None, out_channel=0, tile_rows=None, want_mask=True,
                       out_prob=None, out_mask=None, tile_range=None, accumulate=False, valid=None):
        """generate_chip_indices + predict_chips over a whole (H,W,C) raster on the device
        (``prediction_tools.py:767-776``).  Returns (prob (H,W), mask uint8 (H,W) or None); pixels
        outside the kept cores are left untouched (zero for the rasters allocated here), exactly the footprint
        the reference leaves unpredicted.

        ``out_prob`` may be a caller-owned float32 or float64 (H,W) raster (the reference's float64 template,
        ``:769``); ``accumulate=True`` adds into it like ``predict_chips`` (``:154``) instead of assigning.
        ``tile_rows=(r0, r1)`` / ``tile_range=(t0, t1)`` restrict the call to tile rows / to a chip range of
        the row-major chip list (multi-GPU sharding).  ``valid=(y0, y1, x0, x1)``: pixels outside this window
        reach the network as exact zeros after normalisation.  Rasters allocated here are page-locked."""
        arr, norm = self._split_norm(arr, norm)
        a, dt = _lib.as_input(arr)
        if a.ndim != 3:
            raise ValueError(f'expected (H, W, C) mosaic, got shape {a.shape}')
        H, W, Cc = a.shape
        if out_prob is None:
            out_prob = _lib.pinned_zeros((H, W), np.float32)
        if out_prob.shape != (H, W) or out_prob.dtype not in (np.float32, np.float64) or not out_prob.flags.c_contiguous:
            raise ValueError('out_prob must be a C-contiguous float32 or float64 (H, W) array')
        if want_mask and out_mask is None:
            out_mask = _lib.pinned_zeros((H, W), np.uint8)
        if want_mask and (out_mask.shape != (H, W) or out_mask.dtype != np.uint8 or not out_mask.flags.c_contiguous):
            raise ValueError('out_mask must be a C-contiguous uint8 (H, W) array')
        t = _lib.Tiling(int(kernel), int(buff))
        cn = (norm or NormSpec()).to_c(Cc)
        o = _lib.MosaicOpts()
        o.out_channel = int(out_channel)
        o.out_dtype = _lib.SCV_F64 if out_prob.dtype == np.float64 else _lib.SCV_F32
        o.accumulate = int(bool(accumulate))
        if valid is not None:
            for i, v in enumerate(valid):
                o.valid[i] = int(v)
        if tile_range is not None:
            o.tile_begin, o.tile_end = int(tile_range[0]), int(tile_range[1])
            if o.tile_end <= o.tile_begin:
                return out_prob, (out_mask if want_mask else None)
        elif tile_rows is not None:
            from .sharding import chip_grid
            ys, xs = chip_grid(H, W, int(kernel), int(buff))
            r0, r1 = tile_rows
            r1 = len(ys) if r1 < 0 else min(r1, len(ys))
            if r1 <= max(r0, 0) or not xs:
                return out_prob, (out_mask if want_mask else None)
            o.tile_begin, o.tile_end = max(r0, 0) * len(xs), r1 * len(xs)
        _lib.check(self._lib.scv_predict_mosaic_ex(self._ensure_engine(), _lib.ptr(a), dt, H, W, Cc, C.byref(t),
                                                   C.byref(cn), C.byref(o), _lib.ptr(out_prob),
                                                   _lib.ptr(out_mask) if want_mask else None))
        return out_prob, (out_mask if want_mask else None)

    def predict_patches(self, patches, cols, kernel_shape=(256, 256), kernel_buffer=(128, 128), norm=None,
                        out_channel=0, want_mask=False, crop=None):
        """Patch-list geometry (``prediction_tools.py:245-373, :475-520``): crop every patch and place patch i
        at row i//cols, col i%cols.  ``crop=(y0, y1, x0, x1)`` is the kept window of a patch; by default the
        literal window of ``:258-261`` / ``:340-343`` (which mixes the x / y names, so a non-square
        ``kernel_buffer`` keeps a non-square window -- reproduced as written).  Returns
        (prob (rows*h, cols*w) f32, mask|None)."""
        patches, norm = self._split_norm(patches, norm)
        a, dt = _lib.as_input(patches)
        N, H, W, Cc = a.shape
        if crop is None:
            x_buffer, y_buffer = int(kernel_buffer[0] / 2), int(kernel_buffer[1] / 2)
            crop = (y_buffer, kernel_shape[1] + x_buffer, x_buffer, kernel_shape[0] + y_buffer)
        y0, y1, x0, x1 = (int(v) for v in crop)
        cw = _lib.Crop(y0, x0, y1 - y0, x1 - x0)
        rows = N // cols
        if N % cols:  # the reference's append loop drops a trailing partial row (:269-291)
            a = a[:rows * cols]
            N = rows * cols
        prob = np.empty((rows * cw.h, cols * cw.w), dtype=np.float32)
        mask = np.empty((rows * cw.h, cols * cw.w), dtype=np.uint8) if want_mask else None
        cn = (norm or NormSpec()).to_c(Cc)
        _lib.check(self._lib.scv_predict_patches_ex(self._ensure_engine(), _lib.ptr(a), dt, N, H, W, Cc, C.byref(cw),
                                                    C.byref(cn), int(cols), out_channel, _lib.ptr(prob),
                                                    _lib.ptr(mask)))
        return prob, mask


class SiameseUNetModel(UNetModel):
    """``make_siamese_unet`` (``utils/model_tools.py:638-663``) on the engine: shared single-conv encoder blocks over two
    images, ``DilatedSpatialPyramidPooling`` (1x1 + 3x3 at dilation 3 / 6 / 12 -> 1x1, ``:533-574``) on both pooled
    images, decoder over ``concat([encoded_b, encoded_a, up])``, ``Conv2D(1, 1x1, sigmoid)`` + ``int32(p > class_thresh)``.

    ``predict([input_a, input_b])`` like the two-input Keras model; every tiled entry point (``predict_mosaic``,
    ``predict_patches``, ``prediction_tools.predict_chips`` ...) takes the two rasters stacked along the channel axis,
    ``np.concatenate([a, b], -1)``.  ``get_weights()`` / ``set_weights()`` use one (kernel, bias, gamma, beta, mean,
    variance) group per conv-BN unit; ``set_weights(w, order='keras2')`` accepts the list ``tf.keras`` 2 returns, which
    lists the ASPP layer's trainable weights before its moving statistics."""

    def __init__(self, n_channels, filters=(32, 64, 128), threshold=0.5, bias=None, outputs='both', **engine_kw):
        super().__init__(1, n_channels, list(filters), double_conv=False, head='sigmoid', threshold=threshold, bias=bias,
                         outputs=outputs, arch=_lib.SCV_ARCH_SIAMESE, **engine_kw)

    def keras2_permutation(self):
        """p with ``engine_order[i] = keras2_order[p[i]]`` (see the class docstring)."""
        n_enc = 6 * len(self.filters)
        p = list(range(n_enc))
        for u in range(5):
            p += [n_enc + 4 * u + j for j in range(4)] + [n_enc + 20 + 2 * u + j for j in range(2)]
        return p + list(range(n_enc + 30, len(self.weight_shapes)))

    def set_weights(self, weights, order='grouped'):
        weights = list(weights)
        if order == 'keras2':
            if len(weights) != len(self.weight_shapes):
                raise ValueError(f'expected {len(self.weight_shapes)} weights, got {len(weights)}')
            weights = [weights[i] for i in self.keras2_permutation()]
        elif order != 'grouped':
            raise ValueError("order must be 'grouped' or 'keras2'")
        super().set_weights(weights)

    def load_weights(self, path, by_name=False, skip_mismatch=False, order='grouped'):
        """``.npz`` written as ``np.savez(path, *model.get_weights())`` (``order='keras2'`` for a list taken from the
        reference's ``tf.keras`` model).  Keras HDF5 files of this architecture are not mapped: the reader's layer
        matching (``keras_h5.py``) covers the U-Net families only."""
        if not str(path).endswith('.npz'):
            raise NotImplementedError('SiameseUNetModel.load_weights: only .npz weight lists are supported '
                                      '(np.savez(path, *keras_model.get_weights()), order="keras2")')
        with np.load(path) as z:
            keys = sorted(z.files, key=lambda k: int(k.split('_')[1]) if k.startswith('arr_') else 0)
            self.set_weights([z[k] for k in keys], order=order)

    def save_weights(self, path):
        if not str(path).endswith('.npz'):
            raise NotImplementedError('SiameseUNetModel.save_weights: .npz only')
        np.savez(path, *self._weights)

    def predict(self, x, batch_size=None, verbose=0, steps=None, norm=None):
        if isinstance(x, (list, tuple)) and len(x) == 2 and all(getattr(v, 'ndim', 0) == 4 for v in x):
            a, b = (np.asarray(v) for v in x)
            if a.shape != b.shape or a.shape[-1] != self.nchannels:
                raise ValueError(f'expected two (N, H, W, {self.nchannels}) inputs, got {a.shape} and {b.shape}')
            x = np.concatenate([a, b], axis=-1)
        return super().predict(x, batch_size=batch_size, verbose=verbose, steps=steps, norm=norm)


# ---------------------------------------------------------------- reference builders
def make_siamese_unet(n_channels, filters=[32, 64, 128], factors=[2, 2, 2], bias=None, class_thresh=0.5, **engine_kw):
    """``utils/model_tools.py:638-663``: two-input change-detection U-Net, outputs ``[probs, classes]``.  ``factors``
    must all be 2 (``encoder_block`` pool size == ``decoder_block`` up size on this path)."""
    assert len(filters) == len(factors), 'filters and factors must be same length'
    if any(int(f) != 2 for f in factors):
        raise NotImplementedError('only pooling factor 2 is supported')
    return SiameseUNetModel(n_channels, filters, threshold=class_thresh, bias=bias, **engine_kw)


def get_unet_model(nclasses, nchannels, filters=[32, 64, 128, 256, 512], factors=[2, 2, 2, 2, 2], bias=None,
                   dropout=None, head_name: str = '', double_conv=False, **engine_kw):
    """``utils/model_tools.py:394-415``.  ``dropout`` layers are identity at predict time;
    ``factors`` must all be 2 (the only value the reference's models use on this path).
    ``double_conv=False`` reproduces the network the reference builds as written."""
    assert len(filters) == len(factors), 'number of filters and factors must be equal'
    if any(int(f) != 2 for f in factors):
        raise NotImplementedError('only pooling factor 2 is supported')
    return UNetModel(nclasses, nchannels, filters, double_conv=double_conv, head='softmax', bias=bias,
                     outputs='both', **engine_kw)


def binary_unet(bias=None, nchannels=6, filters=[32, 64, 128, 256, 512], threshold=0.5, outputs='probs',
                **engine_kw):
    """``utils/model_tools.py:417-454`` (broken as committed, Appendix A2) == the notebooks'
    ``get_model()``: classic double-conv U-Net, 1-channel sigmoid head, classes = int32(p > thr)
    (0.9 was used for solar, ``:444``)."""
    return UNetModel(1, nchannels, filters, double_conv=True, head='sigmoid', threshold=threshold, bias=bias,
                     outputs=outputs, **engine_kw)


def get_binary_model(nchannels=6, bias=None, **kw):
    """``utils/model_tools.py:456-...`` (inference part): the compiled wrapper around binary_unet."""
    return binary_unet(bias=bias, nchannels=nchannels, outputs='both', **kw)


def predict_chunk(data, m, norm=None):
    """``utils/model_tools.py:1271-1304`` with the model passed in (Appendix A10): (C,H,W) chunk ->
    ``np.squeeze(pred[0])``."""
    hwc = np.moveaxis(np.asarray(data), 0, -1)
    pred = m.predict(np.expand_dims(hwc, axis=0), norm=norm)
    if isinstance(pred, list):
        pred = pred[0]
    return np.squeeze(pred[0])
