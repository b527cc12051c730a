"""ctypes binding of include/scv.h (libscv.so).  Fails loudly when the library is missing."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SCV_MAX_LEVELS, SCV_MAX_BANDS, SCV_MAX_CLASSES, SCV_MAX_LAYERS = 8, 16, 16, 64
SCV_U8, SCV_U16, SCV_I16, SCV_F32, SCV_F64 = range(5)
SCV_HEAD_SIGMOID, SCV_HEAD_SOFTMAX = 0, 1
(SCV_NORM_NONE, SCV_NORM_PER_BAND, SCV_NORM_PIXEL_MINMAX, SCV_NORM_PIXEL_ZSCORE, SCV_NORM_TILE_ZSCORE,
 SCV_NORM_TILE_MINMAX, SCV_NORM_PIXEL_ZSCORE_SD, SCV_NORM_TILE_GLOBAL_MINMAX, SCV_NORM_TILE_GLOBAL_ZSCORE) = range(9)
SCV_OK, SCV_ERR_INVALID, SCV_ERR_CUDA, SCV_ERR_STATE, SCV_ERR_KERNEL = 0, -1, -2, -3, -4

DTYPES = {np.dtype('uint8'): SCV_U8, np.dtype('uint16'): SCV_U16, np.dtype('int16'): SCV_I16,
          np.dtype('float32'): SCV_F32, np.dtype('float64'): SCV_F64}


class ScvError(RuntimeError):
    """A failing libscv call (status code + scv_last_error text)."""

    def __init__(self, code, msg):
        super().__init__(f'libscv error {code}: {msg}')
        self.code = code


SCV_ARCH_UNET, SCV_ARCH_SIAMESE = 0, 1


class Config(C.Structure):
    _fields_ = [('device', C.c_int), ('double_conv', C.c_int), ('nchannels', C.c_int), ('nclasses', C.c_int),
                ('nlevels', C.c_int), ('filters', C.c_int * SCV_MAX_LEVELS), ('head', C.c_int),
                ('threshold', C.c_float), ('max_batch', C.c_int), ('arch', C.c_int)]


class Tensor(C.Structure):
    _fields_ = [('data', C.POINTER(C.c_float)), ('ndim', C.c_int), ('shape', C.c_int64 * 4)]


class Norm(C.Structure):
    _fields_ = [('mode', C.c_int), ('nbands', C.c_int), ('sub', C.c_float * SCV_MAX_BANDS),
                ('div', C.c_float * SCV_MAX_BANDS), ('ngroups', C.c_int), ('group_size', C.c_int * SCV_MAX_BANDS)]


class Tiling(C.Structure):
    _fields_ = [('kernel', C.c_int), ('buff', C.c_int)]


class MosaicOpts(C.Structure):
    _fields_ = [('tile_begin', C.c_int), ('tile_end', C.c_int), ('out_channel', C.c_int), ('out_dtype', C.c_int),
                ('accumulate', C.c_int), ('valid', C.c_int * 4)]


class Crop(C.Structure):
    _fields_ = [('y0', C.c_int), ('x0', C.c_int), ('h', C.c_int), ('w', C.c_int)]


class Times(C.Structure):
    _fields_ = [('total_ms', C.c_float), ('extract_ms', C.c_float), ('network_ms', C.c_float),
                ('stitch_ms', C.c_float), ('n_batches', C.c_int), ('n_tiles', C.c_int), ('n_launches', C.c_int),
                ('n_layers', C.c_int), ('layer_ms', C.c_float * SCV_MAX_LAYERS),
                ('layer_flops', C.c_double * SCV_MAX_LAYERS), ('h2d_lead_ms', C.c_float), ('d2h_tail_ms', C.c_float)]


# every symbol include/scv.h declares: name -> (restype, argtypes)
_P = C.c_void_p
PROTOTYPES = {
    'scv_version': (C.c_char_p, []),
    'scv_last_error': (C.c_char_p, []),
    'scv_device_count': (C.c_int, []),
    'scv_engine_create': (C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    'scv_engine_destroy': (None, [_P]),
    'scv_num_weights': (C.c_int, [C.POINTER(Config)]),
    'scv_weight_shape': (C.c_int, [C.POINTER(Config), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int64), C.c_char_p, C.c_int]),
    'scv_engine_set_weights': (C.c_int, [_P, C.POINTER(Tensor), C.c_int]),
    'scv_set_option': (C.c_int, [_P, C.c_char_p, C.c_int]),
    'scv_predict_tiles': (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Norm), _P, _P]),
    'scv_predict_mosaic': (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Tiling), C.POINTER(Norm),
                                     C.c_int, C.c_int, C.c_int, _P, _P]),
    'scv_predict_patches': (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Tiling),
                                      C.POINTER(Norm), C.c_int, C.c_int, _P, _P]),
    'scv_predict_mosaic_ex': (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Tiling), C.POINTER(Norm),
                                        C.POINTER(MosaicOpts), _P, _P]),
    'scv_stream_submit': (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Tiling), C.POINTER(Norm),
                                    C.POINTER(MosaicOpts), _P, _P, C.POINTER(C.c_int)]),
    'scv_stream_wait': (C.c_int, [_P, C.c_int]),
    'scv_predict_patches_ex': (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Crop),
                                         C.POINTER(Norm), C.c_int, C.c_int, _P, _P]),
    'scv_predict_mosaic_device_ex': (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Tiling),
                                               C.POINTER(Norm), C.POINTER(MosaicOpts), _P, _P, C.c_int, _P]),
    'scv_check': (C.c_int, [_P]),
    'scv_predict_mosaic_device': (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Tiling),
                                            C.POINTER(Norm), C.c_int, C.c_int, C.c_int, _P, _P, C.c_int, _P]),
    'scv_get_times': (C.c_int, [_P, C.POINTER(Times)]),
    'scv_host_alloc': (_P, [C.c_size_t]),
    'scv_host_free': (None, [_P]),
    'scv_debug_conv3x3': (C.c_int, [C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int, C.c_int, _P, _P]),
    'scv_debug_convT2x2': (C.c_int, [C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, C.c_int, C.c_int, _P]),
    'scv_debug_extract': (C.c_int, [C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Tiling),
                                    C.POINTER(Norm), _P, C.c_int, _P, C.POINTER(C.c_int)]),
}


def lib_path() -> str:
    # SCV_LIB_PATH: kernel-variant experiments only (tools/build_variant.sh); the product is the in-tree libscv.so
    return os.environ.get('SCV_LIB_PATH') or os.path.join(PKG, 'libscv.so')


def load_library():
    """Load libscv.so (built in-tree by ``_build.build()`` / ``__graft_entry__.build()``)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise ImportError(f'{path} is missing: run `python -m satellite_computervision_b200._build` '
                          '(needs nvcc). The CUDA extension is required; there is no CPU fallback.')
    lib = C.CDLL(path)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(code):
    if code != SCV_OK:
        msg = load_library().scv_last_error().decode('utf-8', 'replace')
        if code == SCV_ERR_INVALID:
            raise ValueError(f'libscv: {msg}')
        raise ScvError(code, msg)


def as_input(arr):
    """C-contiguous array of a dtype the engine reads directly (others are converted to float32)."""
    arr = np.asarray(arr)
    if arr.dtype not in DTYPES:
        arr = arr.astype(np.float32)
    arr = np.ascontiguousarray(arr)
    return arr, DTYPES[arr.dtype]


def ptr(arr):
    return arr.ctypes.data_as(C.c_void_p) if arr is not None else None


def pinned_empty(shape, dtype):
    """numpy array backed by cudaHostAlloc memory (overlap-capable H2D/D2H)."""
    lib = load_library()
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    p = lib.scv_host_alloc(max(n, 1))
    if not p:
        raise ScvError(SCV_ERR_CUDA, lib.scv_last_error().decode())
    buf = (C.c_uint8 * max(n, 1)).from_address(p)
    arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
    _PINNED[arr.ctypes.data] = p
    return arr


_PINNED = {}


class _PinnedOwner:
    def __init__(self, p):
        self.p = p

    def __del__(self):
        try:
            if self.p and _LIB is not None:
                _LIB.scv_host_free(self.p)
        except Exception:
            pass


def pinned_zeros(shape, dtype):
    """Zero-filled page-locked array whose memory is released with the array (results handed to callers)."""
    lib = load_library()
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    n = max(count * dtype.itemsize, 1)
    p = lib.scv_host_alloc(n)
    if not p:
        raise ScvError(SCV_ERR_CUDA, lib.scv_last_error().decode())
    C.memset(p, 0, n)
    buf = (C.c_uint8 * n).from_address(p)
    buf._scv_owner = _PinnedOwner(p)  # every view's .base chain ends at `buf`: the memory lives as long as any view
    return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)


def pinned_free(arr):
    p = _PINNED.pop(arr.ctypes.data, None)
    if p:
        load_library().scv_host_free(p)
