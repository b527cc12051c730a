"""Generate golden vectors from the REAL reference source (build container only).

Run here (``/root/reference`` is read-only and absent on the GPU box):

    python tests/golden/make_golden.py

Imports ``/root/reference/utils/{prediction_tools,processing,array_tools}.py``
unmodified, with the third-party modules that are not installable in this
image (tensorflow, matplotlib, rasterio) replaced by inert stubs.  Only
functions whose arithmetic is pure Python/numpy are exercised, so the stubs
never compute anything except where noted:

* ``tf.sqrt`` -> ``np.sqrt``, ``tf.split`` / ``tf.concat`` -> numpy equivalents
  (structure only) for the ``moments=`` / ``splits=`` paths of
  ``normalize_tensor`` / ``rescale_tensor``.
* ``rasterio.open`` -> an object that captures what ``dst.write`` receives.

The model is a deterministic element-wise stand-in (``FakeModel``) so the
fixtures do not depend on any convolution implementation: what is pinned is the
reference's tile origins, crop windows, stitch placement and normaliser
arithmetic.  Outputs: ``tests/golden/*.npz`` (committed).
"""
import json
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference/utils'


# ------------------------------------------------------------------ stubs
class _Anything:
    def __getattr__(self, k):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__getattr__ = lambda k: _Anything()
    sys.modules[name] = m
    return m


CAPTURED = {}


class _RioWriter:
    def __init__(self, path, mode, **kw):
        self.path, self.kw = path, kw

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def write(self, arr):
        CAPTURED[os.path.basename(self.path)] = (np.array(arr), dict(self.kw))


def install_stubs():
    tf = _module('tensorflow')
    tf.sqrt = np.sqrt
    tf.split = lambda t, sizes, axis=0: np.split(t, np.cumsum(sizes)[:-1], axis=axis)
    tf.concat = lambda ts, axis=0: np.concatenate(ts, axis=axis)
    keras = types.SimpleNamespace(utils=types.SimpleNamespace(Sequence=object))
    tf.keras = keras
    _module('matplotlib', pyplot=_Anything())
    _module('matplotlib.pyplot')
    rio = _module('rasterio')
    rio.open = lambda path, mode='r', **kw: _RioWriter(path, mode, **kw)
    rio.Affine = lambda *a: tuple(a)
    _module('rasterio.crs', CRS=_Anything())
    _module('rasterio.warp', transform_bounds=_Anything())
    _module('rasterio.transform', array_bounds=_Anything())


class FakeModel:
    """Element-wise stand-in for ``keras.Model``: 2 output channels,
    p0 = sum_c w_c * x_c evaluated left to right in float32, p1 = 1 - p0."""

    def __init__(self, as_list=False):
        self.as_list = as_list

    def _one(self, x):
        x = np.asarray(x, dtype=np.float32)
        w = np.float32(0.5)
        p0 = np.zeros(x.shape[:-1], dtype=np.float32)
        for c in range(x.shape[-1]):
            p0 = p0 + x[..., c] * w
            w = w * np.float32(0.5)
        probs = np.stack([p0, np.float32(1.0) - p0], axis=-1)
        if self.as_list:
            return [probs, (probs[..., :1] > 0.25).astype(np.int32)]
        return probs

    def predict(self, x, steps=None, verbose=0):
        if isinstance(x, np.ndarray):
            return self._one(x)
        batches = []
        for i, b in enumerate(x):
            if steps is not None and i >= steps:
                break
            batches.append(self._one(b))
        if self.as_list:
            return [np.concatenate([b[k] for b in batches], axis=0) for k in range(2)]
        return np.concatenate(batches, axis=0)


class FakeDataset(list):
    """Iterable of (1,h,w,C) batches whose iterator has TF1-style ``.next()``
    (``prediction_tools.py:515``)."""

    def __iter__(self):
        it = list.__iter__(self)

        class _It:
            def next(self_inner):
                return next(it)

            __next__ = next

            def __iter__(self_inner):
                return self_inner
        return _It()


def main():
    install_stubs()
    sys.path.insert(0, REF)
    _stdout = sys.stdout
    sys.stdout = open(os.devnull, 'w')  # the reference prints at import and per tile
    try:
        import array_tools
        import prediction_tools as pt
        import processing
        out = build(pt, processing, array_tools)
    finally:
        sys.stdout = _stdout
    out['raster_chip_indices'] = raster_tools_indices()
    for name, payload in out.items():
        path = os.path.join(HERE, name + '.npz')
        np.savez_compressed(path, **payload)
        print('wrote', path, os.path.getsize(path), 'bytes')


def raster_tools_indices():
    """raster_tools.generate_chip_indices (utils/raster_tools.py:23-46, per-side buffer): the module imports
    rasterio / geopandas / GDAL at the top, so the function is lifted out of the unmodified source file with
    ``ast`` and executed on its own (it is pure Python)."""
    import ast
    src = open(os.path.join(REF, 'raster_tools.py')).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == 'generate_chip_indices')
    ns = {}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), 'raster_tools.py', 'exec'), ns)
    gen = ns['generate_chip_indices']
    out = {}
    cases = [(2048, 2048, 64, 256), (10980, 10980, 64, 256), (384, 384, 64, 256), (383, 640, 64, 256),
             (512, 512, 128, 256), (1000, 700, 32, 128), (100, 100, 0, 32), (97, 131, 3, 10), (50, 50, 64, 256)]
    for n, (H, W, buff, kernel) in enumerate(cases):
        out[f'case{n}_params'] = np.array([H, W, buff, kernel], dtype=np.int64)
        out[f'case{n}_indices'] = np.array(gen(H, W, buff, kernel), dtype=np.int64).reshape(-1, 2)
    out['ncases'] = np.array(len(cases))
    return out


def build(pt, processing, array_tools):
    out = {}
    rng = np.random.default_rng(20261017)

    # ---- a1: generate_chip_indices (prediction_tools.py:87-109)
    idx = {}
    cases = [(2048, 2048, 128, 256), (10980, 10980, 128, 256), (384, 384, 128, 256),
             (385, 385, 128, 256), (449, 640, 128, 256), (640, 900, 128, 256),
             (1000, 700, 128, 256), (2000, 2000, 256, 512), (150, 170, 16, 32),
             (100, 100, 0, 32), (97, 131, 6, 10)]
    for n, (H, W, buff, kernel) in enumerate(cases):
        arr = np.empty((H, W, 1), dtype=np.uint8)
        res = np.array(pt.generate_chip_indices(arr, buff, kernel), dtype=np.int64).reshape(-1, 2)
        idx[f'case{n}_params'] = np.array([H, W, buff, kernel], dtype=np.int64)
        idx[f'case{n}_indices'] = res
    idx['ncases'] = np.array(len(cases))
    out['chip_indices'] = idx

    # ---- a2/a3: extract_chips, predict_chips (:111-156) on a small raster
    arr = rng.random((150, 170, 3), dtype=np.float32)
    buff, kernel = 16, 32
    indices = pt.generate_chip_indices(arr, buff, kernel)
    chips = pt.extract_chips(arr, buff, kernel)
    template = np.zeros(arr.shape[:2])
    res = pt.predict_chips(arr, indices, template, FakeModel(), kernel, buff)
    assert res is template
    # overlapping / repeated indices: += accumulates
    template2 = np.full(arr.shape[:2], 0.5)
    idx2 = indices[:5] + indices[:2] + [(20, 24)]
    res2 = pt.predict_chips(arr, idx2, template2, FakeModel(), kernel, buff)
    # square raster for the transposed extract_chips quirk (:127)
    arr_sq = rng.random((120, 120, 2), dtype=np.float32)
    chips_sq = pt.extract_chips(arr_sq, buff, kernel)
    out['predict_chips'] = dict(
        arr=arr, buff=np.array(buff), kernel=np.array(kernel),
        indices=np.array(indices, dtype=np.int64),
        chips_shapes=np.array([c.shape for c in chips], dtype=np.int64),
        arr_sq=arr_sq, chips_sq=np.stack(chips_sq),
        template=res, indices2=np.array(idx2, dtype=np.int64), template2=res2)

    # ---- a15: make_array_predictions / callback_predictions / write_geotiff_predictions
    kshape, kbuf = [32, 32], [16, 16]
    cols, rows = 3, 2
    patches = [rng.random((1, 48, 48, 3), dtype=np.float32) for _ in range(cols * rows)]
    mixer = {'patchesPerRow': cols, 'totalPatches': cols * rows, 'patchDimensions': kshape,
             'projection': {'crs': 'EPSG:32618',
                            'affine': {'doubleMatrix': [10.0, 0.0, 500000.0, 0.0, -10.0, 4400000.0]}}}
    with tempfile.TemporaryDirectory() as td:
        jf = os.path.join(td, 'mixer.json')
        with open(jf, 'w') as f:
            json.dump(mixer, f)
        map_arr = pt.make_array_predictions(FakeDataset(patches), FakeModel(), jf, kshape, kbuf)
        map_list = pt.make_array_predictions(FakeDataset(patches), FakeModel(as_list=True), jf, kshape, kbuf)
        cb = pt.callback_predictions(FakeDataset(patches), FakeModel(), mixer, kshape, kbuf)
        cb_list = pt.callback_predictions(FakeDataset(patches), FakeModel(as_list=True), mixer, kshape, kbuf)
        pt.write_geotiff_predictions(FakeDataset(patches), FakeModel(), jf, 'gold', td, kbuf)
        gt, gt_kw = CAPTURED['gold.tif']
        # non-square kernel to pin the x/y naming mix at :258-261
        kshape2, kbuf2 = [32, 32], [8, 16]
        map_ns = pt.make_array_predictions(FakeDataset(patches), FakeModel(), jf, kshape2, kbuf2)
    out['patch_stitch'] = dict(
        patches=np.concatenate(patches, axis=0), cols=np.array(cols), rows=np.array(rows),
        kernel_shape=np.array(kshape), kernel_buffer=np.array(kbuf),
        make_array=map_arr, make_array_list=map_list, callback=cb, callback_list=cb_list,
        geotiff=gt, geotiff_wh=np.array([gt_kw['width'], gt_kw['height'], gt_kw['count']]),
        kernel_buffer_ns=np.array(kbuf2), make_array_ns=map_ns)

    # ---- a5/a6: rescale_tensor / normalize_tensor with moments (processing.py:225-322)
    img = (rng.random((24, 20, 6), dtype=np.float32) * np.float32(10000.0)).astype(np.float32)
    img_u16 = rng.integers(0, 10000, (24, 20, 6), dtype=np.uint16)
    mm = [(0, 10000)] * 6
    mm2 = [(100.0, 9000.0), (0.0, 8000.5), (50.0, 50.0), (1.0, 3000.0), (0.0, 10000.0), (200.0, 7000.0)]
    mv = [(1200.5, 250000.0), (1100.0, 300000.0), (900.25, 1.0), (3000.0, 900000.0), (2000.0, 0.0), (1500.0, 400000.0)]
    norm = dict(
        img=img, img_u16=img_u16,
        mm=np.array(mm, dtype=np.float64), mm2=np.array(mm2, dtype=np.float64), mv=np.array(mv, dtype=np.float64),
        rescale_mm=processing.rescale_tensor(img, moments=mm),
        rescale_mm_u16=processing.rescale_tensor(img_u16, moments=mm),
        rescale_mm2=processing.rescale_tensor(img, moments=mm2),
        normalize_mv=processing.normalize_tensor(img, moments=mv),
        # splits path (structure ops tf.split/tf.concat stubbed with numpy)
        rescale_split=processing.rescale_tensor(img, moments=[(0, 10000)] * 3, splits=[3, 3]),
        normalize_split=processing.normalize_tensor(img, moments=[(1000.0, 250000.0)] * 2, splits=[2, 2]),
        # numpy twins (array_tools.py:47-157): data-derived statistics, pure numpy
        at_rescale_axes01=array_tools.rescale_array(img, axes=(0, 1)),
        at_rescale_axes2=array_tools.rescale_array(img, axes=2),
        at_normalize_axes01=array_tools.normalize_array(img, axes=(0, 1)),
        at_normalize_axes2=array_tools.normalize_array(img, axes=(2,)),
    )
    out['normalize'] = norm
    return out


if __name__ == '__main__':
    main()
