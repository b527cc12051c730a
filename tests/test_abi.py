"""The C-ABI library loads and exports every symbol include/scv.h declares; the
host-side architecture bookkeeping behind it agrees with the oracle. No GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import unet as ounet
from satellite_computervision_b200 import _build, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    _build.build()
    return _lib.load_library()


def test_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, 'include', 'scv.h')).read()
    declared = re.findall(r'SCV_API\s+[\w\s\*]+?\b(scv_\w+)\s*\(', header)
    assert len(declared) >= 25
    assert sorted(declared) == sorted(_lib.PROTOTYPES)
    for name in declared:
        assert hasattr(lib, name), name
    assert b'sm_100a' in lib.scv_version()


def test_version_string():
    lib = _lib.load_library()
    assert b'0.3' in lib.scv_version() and b'sm_100a' in lib.scv_version()


def test_struct_layouts_match_header():
    assert C.sizeof(_lib.Config) == 4 * (5 + 8 + 4)  # ... + arch (0.3)
    assert C.sizeof(_lib.Tensor) == 8 + 8 + 32
    assert C.sizeof(_lib.Norm) == 8 + 2 * 16 * 4 + 4 + 16 * 4
    assert C.sizeof(_lib.Tiling) == 8
    assert C.sizeof(_lib.MosaicOpts) == 4 * 9
    assert C.sizeof(_lib.Crop) == 16
    assert C.sizeof(_lib.Times) == 4 * 4 + 4 * 4 + 64 * 4 + 64 * 8 + 8


@pytest.mark.parametrize('double_conv,ncls,nch,filters', [(1, 1, 6, [32, 64, 128, 256, 512]), (0, 2, 6, [32, 64, 128, 256, 512]),
                                                         (1, 1, 3, [32, 64]), (0, 4, 4, [64, 128, 256])])
def test_weight_specs_match_oracle(lib, double_conv, ncls, nch, filters):
    cfg = _lib.Config()
    cfg.double_conv, cfg.nchannels, cfg.nclasses, cfg.nlevels = double_conv, nch, ncls, len(filters)
    for i, f in enumerate(filters):
        cfg.filters[i] = f
    cfg.head = _lib.SCV_HEAD_SIGMOID if ncls == 1 else _lib.SCV_HEAD_SOFTMAX
    specs = ounet.weight_specs('A' if double_conv else 'B', nch, ncls, tuple(filters))
    assert lib.scv_num_weights(C.byref(cfg)) == len(specs)
    for i, (name, shape) in enumerate(specs):
        nd, shp, buf = C.c_int(), (C.c_int64 * 4)(), C.create_string_buffer(128)
        assert lib.scv_weight_shape(C.byref(cfg), i, C.byref(nd), shp, buf, 128) == 0
        assert buf.value.decode() == name
        assert tuple(shp[d] for d in range(nd.value)) == tuple(shape)


def test_invalid_configs_are_rejected(lib):
    cfg = _lib.Config()
    cfg.nchannels, cfg.nclasses, cfg.nlevels = 6, 1, 2
    cfg.filters[0], cfg.filters[1] = 24, 64
    assert lib.scv_num_weights(C.byref(cfg)) == _lib.SCV_ERR_INVALID
    assert b'filters[0]=24' in lib.scv_last_error()
    cfg.filters[0] = 32
    cfg.nclasses = 3  # sigmoid head with 3 classes
    assert lib.scv_num_weights(C.byref(cfg)) == _lib.SCV_ERR_INVALID


def test_no_cpu_fallback_without_a_device(lib):
    if lib.scv_device_count() > 0:
        pytest.skip('a CUDA device is present')
    cfg = _lib.Config()
    cfg.nchannels, cfg.nclasses, cfg.nlevels, cfg.double_conv = 6, 1, 1, 1
    cfg.filters[0] = 32
    h = C.c_void_p()
    rc = lib.scv_engine_create(C.byref(cfg), C.byref(h))
    assert rc == _lib.SCV_ERR_CUDA and not h
    assert b'no CPU fallback' in lib.scv_last_error()
    x = np.zeros((1, 8, 16, 32), np.float32)
    k = np.zeros((3, 3, 32, 32), np.float32)
    b = np.zeros(32, np.float32)
    y = np.zeros((1, 8, 16, 32), np.float32)
    rc = lib.scv_debug_conv3x3(0, _lib.ptr(x), 1, 8, 16, 32, _lib.ptr(k), _lib.ptr(b), 32, 1, _lib.ptr(y), None)
    assert rc == _lib.SCV_ERR_CUDA
