"""GPU parity of the individual kernels through the C-ABI debug entry points:
UMMA implicit-GEMM conv / convT against a torch fp32 reference of the same op
(bf16 operands, fp32 accumulate: tolerance = 1 bf16 ulp of the result + 2e-3), and
the extract kernel bit-exactly against the oracle normalisers."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import normalize as onorm
from oracle import tiling as otile
from satellite_computervision_b200 import _lib, processing
from tests import gpu_util as G

pytestmark = pytest.mark.gpu

CONV_CASES = [
    # N, H, W, Cin, Cout            exercised path
    (2, 32, 32, 6, 32),           # KC=16 (6->16 pad), SW32, BN=32, box 16x8x1
    (1, 16, 48, 3, 64),           # KC=16, BN=64
    (2, 32, 32, 32, 32),          # KC=32, SW64
    (1, 32, 32, 32, 64),          # KC=32, BN=64
    (2, 16, 32, 64, 32),          # KC=64, SW128, BN=32
    (1, 32, 32, 64, 64),
    (1, 16, 32, 64, 128),         # BN=128
    (1, 16, 16, 128, 128),        # 2 channel chunks per tap
    (1, 16, 16, 128, 256),        # BN=256
    (1, 8, 16, 256, 512),         # 2 N tiles
    (3, 24, 24, 64, 64),          # box 8x8x2, odd N -> OOB image in the last tile
    (5, 12, 12, 128, 256),        # box 4x4x8, N tail
    (1, 20, 20, 32, 32),          # partial tiles in x and y (masked stores)
    (1, 40, 72, 64, 32),          # non-square
]


# Row-streaming tap-packed kernel (conv_rows.cuh): widths that are multiples of 128, Cout 32 / 64.
# `grid` caps the number of persistent CTAs so one CTA streams many rows: segment changes, ring wrap
# (16 row accumulators at Cout 32, 8 at Cout 64) and split UMMAs at the wrap are all exercised.
ROWS_CASES = [
    # N, H, W, Cin, Cout, grid
    (1, 8, 128, 32, 32, 0),
    (2, 6, 256, 64, 32, 0),
    (1, 10, 128, 32, 64, 0),
    (3, 4, 128, 64, 64, 0),
    (1, 4, 384, 128, 64, 0),      # two channel chunks
    (2, 40, 128, 32, 32, 1),      # one CTA: 80 rows, two segments, ring wraps 5 times
    (2, 22, 256, 64, 64, 3),      # uneven split over 3 CTAs: segments start mid-image
    (1, 36, 128, 128, 32, 2),     # two chunks per row, streaming
    (3, 18, 128, 16, 64, 2),      # KC=16 is not a row-kernel shape: must still be right (slab / tile path)
    (2, 12, 192, 64, 64, 0),      # 1.5 strips: the second strip is partial (TMA zero-fills its loads, clips its stores)
    (1, 16, 192, 32, 64, 2),
    (1, 6, 320, 64, 32, 0),       # 2.5 strips
]


@pytest.mark.parametrize('N,H,W,Cin,Cout,grid', ROWS_CASES)
def test_conv3x3_rows_kernel_matches_torch(N, H, W, Cin, Cout, grid, monkeypatch):
    monkeypatch.setenv('SCV_ROWS_PARTIAL', '1')  # widths that are not a multiple of 128 stay in the row kernel
    if grid:
        monkeypatch.setenv('SCV_DEBUG_GRID', str(grid))
    rng = np.random.default_rng(N * 1000 + H + W + Cin + Cout)
    x = rng.standard_normal((N, H, W, Cin)).astype(np.float32)
    k = (rng.standard_normal((3, 3, Cin, Cout)) / np.sqrt(9 * Cin)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32) * 0.1
    got = G.conv3x3_device(x, k, b)
    ref = G.conv3x3_ref(x, k, b)
    s = G.err_stats(got, ref)
    print('conv3x3 rows', (N, H, W, Cin, Cout, grid), s)
    assert s['nan'] == 0 and s['n_bad'] == 0, s


@pytest.mark.parametrize('N,H,W,Cin,Cout,grid', [(2, 8, 128, 32, 32, 0), (1, 12, 256, 64, 64, 0), (2, 36, 128, 32, 32, 1),
                                                 (2, 20, 256, 32, 64, 3), (2, 12, 192, 64, 64, 0), (1, 24, 192, 32, 64, 2)])
def test_rows_kernel_fused_maxpool(N, H, W, Cin, Cout, grid, monkeypatch):
    monkeypatch.setenv('SCV_ROWS_PARTIAL', '1')
    if grid:
        monkeypatch.setenv('SCV_DEBUG_GRID', str(grid))
    rng = np.random.default_rng(17)
    x = rng.standard_normal((N, H, W, Cin)).astype(np.float32)
    k = (rng.standard_normal((3, 3, Cin, Cout)) / np.sqrt(9 * Cin)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32) * 0.1
    y, p = G.conv3x3_device(x, k, b, pooled=True)
    assert G.err_stats(y, G.conv3x3_ref(x, k, b))['n_bad'] == 0
    assert np.array_equal(p, G.maxpool_ref(y))


def test_rows_kernel_equals_slab_kernel_bitwise(monkeypatch):
    """Same summation order as the 8x16-tile kernels whenever Cin fits one chunk: identical bits."""
    rng = np.random.default_rng(23)
    x = rng.standard_normal((2, 16, 128, 64)).astype(np.float32)
    k = (rng.standard_normal((3, 3, 64, 32)) / 24).astype(np.float32)
    b = rng.standard_normal(32).astype(np.float32) * 0.1
    rows = G.conv3x3_device(x, k, b)
    monkeypatch.setenv('SCV_ROWS', '0')
    other = G.conv3x3_device(x, k, b)
    assert np.array_equal(rows, other)
    # ... and with a partial last strip (192-pixel rows, the Cout = 64 layers of the 192 x 192 level)
    monkeypatch.delenv('SCV_ROWS')
    monkeypatch.setenv('SCV_ROWS_PARTIAL', '1')
    x2 = rng.standard_normal((2, 16, 192, 64)).astype(np.float32)
    k2 = (rng.standard_normal((3, 3, 64, 64)) / 24).astype(np.float32)
    b2 = rng.standard_normal(64).astype(np.float32) * 0.1
    rows2 = G.conv3x3_device(x2, k2, b2)
    monkeypatch.setenv('SCV_ROWS', '0')
    assert np.array_equal(rows2, G.conv3x3_device(x2, k2, b2))


@pytest.mark.parametrize('N,H,W,Cin,Cout', CONV_CASES)
def test_conv3x3_matches_torch(N, H, W, Cin, Cout):
    rng = np.random.default_rng(N * 1000 + H + W + Cin + Cout)
    x = rng.standard_normal((N, H, W, Cin)).astype(np.float32)
    k = (rng.standard_normal((3, 3, Cin, Cout)) / np.sqrt(9 * Cin)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32) * 0.1
    got = G.conv3x3_device(x, k, b)
    ref = G.conv3x3_ref(x, k, b)
    s = G.err_stats(got, ref)
    print('conv3x3', (N, H, W, Cin, Cout), s)
    assert s['nan'] == 0 and s['n_bad'] == 0, s


@pytest.mark.parametrize('N,H,W,Cin,Cout,grid', [
    (1, 16, 16, 128, 256, 1),     # one CTA walks all tiles, 2 channel chunks
    (1, 8, 16, 256, 512, 2),      # 2 N tiles: each CTA keeps its N tile
    (5, 12, 12, 128, 256, 3),     # box 4x4x8, image tail, odd tile count over 3 CTAs
    (3, 24, 24, 64, 64, 2),       # box 8x8x2
    (2, 20, 20, 32, 128, 4),      # partial tiles (masked stores), KC=32
    (1, 16, 48, 3, 64, 1),        # KC=16
])
def test_persistent_tile_kernel_equals_tile_kernel(N, H, W, Cin, Cout, grid, monkeypatch):
    """conv_ptile_kernel (persistent, two TMEM accumulators) against torch and, bit for bit, against the
    one-tile-per-CTA kernel it replaces for large launches."""
    rng = np.random.default_rng(31 + N + H + Cin)
    x = rng.standard_normal((N, H, W, Cin)).astype(np.float32)
    k = (rng.standard_normal((3, 3, Cin, Cout)) / np.sqrt(9 * Cin)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32) * 0.1
    monkeypatch.setenv('SCV_ROWS', '0')
    monkeypatch.setenv('SCV_SLAB', '0')
    monkeypatch.setenv('SCV_PTILE', '0')
    ref_kernel = G.conv3x3_device(x, k, b)
    monkeypatch.setenv('SCV_PTILE', '2')
    monkeypatch.setenv('SCV_DEBUG_GRID', str(grid))
    got = G.conv3x3_device(x, k, b)
    s = G.err_stats(got, G.conv3x3_ref(x, k, b))
    assert s['nan'] == 0 and s['n_bad'] == 0, s
    assert np.array_equal(got, ref_kernel)
    if H % 2 == 0 and W % 2 == 0:
        y, p = G.conv3x3_device(x, k, b, pooled=True)
        assert np.array_equal(y, ref_kernel) and np.array_equal(p, G.maxpool_ref(y))


@pytest.mark.parametrize('N,H,W,Cin,Cout,grid', [
    (2, 32, 32, 128, 128, 0),     # 2 chunks, 16 tiles over 16 CTAs (single-tile passes)
    (1, 48, 96, 256, 128, 3),     # 4 chunks, 36 tiles over 3 CTAs: 6 two-tile passes each, rings wrap
    (3, 16, 24, 128, 128, 2),     # 9 tiles over 2 CTAs: odd tile counts -> a one-tile tail pass
    (1, 32, 16, 64, 128, 1),      # one chunk
    (2, 32, 48, 128, 64, 3),      # Cout = 64 with two chunks (decoder_1/conv0 at 192x192): N = 64 weight tiles, 10-deep ring
    (1, 48, 32, 256, 64, 2),
])
def test_weight_streaming_slab_kernel(N, H, W, Cin, Cout, grid, monkeypatch):
    """conv_slabw_kernel (halo slabs + streamed weights shared by two M tiles) against torch and, bit for bit,
    against the one-tile-per-CTA kernel (same (chunk, tap, k) summation order); pooled epilogue too."""
    rng = np.random.default_rng(51 + N + H + Cin)
    x = rng.standard_normal((N, H, W, Cin)).astype(np.float32)
    k = (rng.standard_normal((3, 3, Cin, Cout)) / np.sqrt(9 * Cin)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32) * 0.1
    for name in ('SCV_ROWS', 'SCV_SLAB', 'SCV_PTILE', 'SCV_SLABW'):
        monkeypatch.setenv(name, '0')
    ref_kernel = G.conv3x3_device(x, k, b)
    monkeypatch.setenv('SCV_SLABW', '2')
    monkeypatch.setenv('SCV_SLABW_MIN_CIN', '64')
    if grid:
        monkeypatch.setenv('SCV_DEBUG_GRID', str(grid))
    got = G.conv3x3_device(x, k, b)
    s = G.err_stats(got, G.conv3x3_ref(x, k, b))
    assert s['nan'] == 0 and s['n_bad'] == 0, s
    assert np.array_equal(got, ref_kernel)
    y, p = G.conv3x3_device(x, k, b, pooled=True)
    assert np.array_equal(y, ref_kernel) and np.array_equal(p, G.maxpool_ref(y))


@pytest.mark.parametrize('N,H,W,Cin,Cout,grid', [
    (2, 32, 32, 64, 64, 0),       # one chunk, 16 tiles = 8 pairs, single pass
    (1, 48, 96, 128, 64, 4),      # two chunks (decoder_1/conv0 shape), 36 tiles over 2 pairs: 9 rounds, rings wrap
    (3, 32, 16, 32, 64, 2),       # KC = 32 (encoder_1/conv0 shape), 12 tiles over one pair, two issuers alternate
    (2, 16, 48, 64, 64, 6),       # 12 tiles over 3 pairs
    (1, 64, 64, 192, 64, 2),      # three chunks: three slabs, single issuer, one pair
])
def test_cta_pair_slab_kernel(N, H, W, Cin, Cout, grid, monkeypatch):
    """conv_slab2_kernel (tcgen05.mma.cta_group::2: M = 256 over a CTA pair, each CTA holds half of the weights)
    against torch and, bit for bit, against the one-tile-per-CTA kernel (same (chunk, tap, k) summation order);
    pooled + skip epilogue too."""
    rng = np.random.default_rng(77 + N + H + Cin)
    x = rng.standard_normal((N, H, W, Cin)).astype(np.float32)
    k = (rng.standard_normal((3, 3, Cin, Cout)) / np.sqrt(9 * Cin)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32) * 0.1
    for name in ('SCV_ROWS', 'SCV_SLAB', 'SCV_PTILE', 'SCV_SLABW', 'SCV_SLAB2'):
        monkeypatch.setenv(name, '0')
    ref_kernel = G.conv3x3_device(x, k, b)
    monkeypatch.setenv('SCV_SLAB2', '2')
    if grid:
        monkeypatch.setenv('SCV_DEBUG_GRID', str(grid))
    got = G.conv3x3_device(x, k, b)
    s = G.err_stats(got, G.conv3x3_ref(x, k, b))
    assert s['nan'] == 0 and s['n_bad'] == 0, s
    assert np.array_equal(got, ref_kernel)
    y, p = G.conv3x3_device(x, k, b, pooled=True)
    assert np.array_equal(y, ref_kernel) and np.array_equal(p, G.maxpool_ref(y))


def test_persistent_tile_kernel_convT(monkeypatch):
    rng = np.random.default_rng(41)
    x = rng.standard_normal((2, 6, 6, 1024)).astype(np.float32)
    k = (rng.standard_normal((2, 2, 512, 1024)) / 32).astype(np.float32)
    b = rng.standard_normal(512).astype(np.float32) * 0.1
    monkeypatch.setenv('SCV_PTILE', '0')
    ref_kernel = G.convT_device(x, k, b)
    monkeypatch.setenv('SCV_PTILE', '2')
    monkeypatch.setenv('SCV_DEBUG_GRID', '8')
    got = G.convT_device(x, k, b)
    assert G.err_stats(got, G.convT_ref(x, k, b))['n_bad'] == 0 and np.array_equal(got, ref_kernel)


def test_conv3x3_without_relu_keeps_negatives():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((1, 16, 16, 32)).astype(np.float32)
    k = (rng.standard_normal((3, 3, 32, 32)) / 17).astype(np.float32)
    b = np.zeros(32, np.float32)
    got = G.conv3x3_device(x, k, b, relu=False)
    s = G.err_stats(got, G.conv3x3_ref(x, k, b, relu=False))
    assert s['n_bad'] == 0 and (got < 0).mean() > 0.3


@pytest.mark.parametrize('N,H,W,Cin,Cout', [(2, 32, 32, 32, 32), (2, 24, 24, 64, 64), (8, 4, 4, 64, 32), (1, 16, 48, 6, 32)])
def test_fused_maxpool_epilogue(N, H, W, Cin, Cout):
    rng = np.random.default_rng(7)
    x = rng.standard_normal((N, H, W, Cin)).astype(np.float32)
    k = (rng.standard_normal((3, 3, Cin, Cout)) / np.sqrt(9 * Cin)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32) * 0.1
    y, p = G.conv3x3_device(x, k, b, pooled=True)
    ref = G.conv3x3_ref(x, k, b)
    assert G.err_stats(y, ref)['n_bad'] == 0
    # the pooled tensor must be exactly the 2x2 max of the full-resolution output the same kernel wrote
    assert np.array_equal(p, G.maxpool_ref(y))


@pytest.mark.parametrize('N,H,W,Cin,Cout', [(1, 16, 16, 64, 32), (2, 12, 12, 128, 64), (2, 6, 6, 1024, 512), (1, 8, 24, 256, 128)])
def test_convT2x2_matches_torch(N, H, W, Cin, Cout):
    rng = np.random.default_rng(11)
    x = rng.standard_normal((N, H, W, Cin)).astype(np.float32)
    k = (rng.standard_normal((2, 2, Cout, Cin)) / np.sqrt(Cin)).astype(np.float32)
    b = rng.standard_normal(Cout).astype(np.float32) * 0.1
    got = G.convT_device(x, k, b)
    s = G.err_stats(got, G.convT_ref(x, k, b))
    print('convT', (N, H, W, Cin, Cout), s)
    assert s['nan'] == 0 and s['n_bad'] == 0, s


def _extract(hwc, tiling, norm, indices):
    lib = _lib.load_library()
    a, dt = _lib.as_input(hwc)
    H, W, Cc = a.shape
    t = _lib.Tiling(*tiling)
    side = tiling[0] + tiling[1]
    idx = np.ascontiguousarray(np.array(indices, dtype=np.int32).reshape(-1, 2))
    out = np.empty((len(idx) * side * side * 16,), np.float32)
    cpad = C.c_int()
    _lib.check(lib.scv_debug_extract(0, _lib.ptr(a), dt, H, W, Cc, C.byref(t), C.byref(norm.to_c(Cc)), _lib.ptr(idx),
                                     len(idx), _lib.ptr(out), C.byref(cpad)))
    assert cpad.value == (8 if Cc <= 8 else 16)  # one pixel = one or two 128-bit stores
    return out[:len(idx) * side * side * cpad.value].reshape(len(idx), side, side, cpad.value)


@pytest.mark.parametrize('dtype', ['uint16', 'float32', 'float64', 'uint8', 'int16'])
def test_extract_per_band_bit_exact(dtype):
    """Chip placement bit-exact and values == bf16(reference float32 normaliser output)."""
    rng = np.random.default_rng(3)
    H, W, Cc = 300, 333, 6  # odd width: rows start at every 16-byte misalignment
    if dtype == 'uint16':
        arr = rng.integers(0, 10000, (H, W, Cc), dtype=np.uint16)
    elif dtype == 'uint8':
        arr = rng.integers(0, 255, (H, W, Cc), dtype=np.uint8)
    elif dtype == 'int16':
        arr = rng.integers(-2000, 10000, (H, W, Cc)).astype(np.int16)
    else:
        arr = (rng.random((H, W, Cc)) * 10000).astype(dtype)
    kernel, buff = 64, 32
    idx = otile.generate_chip_indices(arr.shape, buff, kernel)
    assert len(idx) == 12
    mm = [(0, 10000), (10.0, 9000.0), (0.0, 8000.5), (1.0, 3000.0), (0.0, 10000.0), (200.0, 7000.0)]
    spec = processing.rescale_spec(Cc, moments=mm)
    got = _extract(arr, (kernel, buff), spec, idx)
    for i, (y, x) in enumerate(idx):
        chip = arr[y - 16:y + 80, x - 16:x + 80, :]
        want = G.bf16_round(onorm.rescale_tensor(chip.astype(np.float32), moments=mm))
        assert np.array_equal(got[i, :, :, :Cc], want), (dtype, i)
        assert np.all(got[i, :, :, Cc:] == 0)


def test_extract_u16_fast_path_bit_exact(monkeypatch):
    """uint16 x 6 bands x per-band constants runs a dedicated kernel (with and without a subtract): the same
    bits as the generic kernel and as bf16(reference normaliser), including every value that lands within a
    few fp32 ulps of a bf16 rounding boundary (all 65536 digital numbers are present)."""
    rng = np.random.default_rng(9)
    H, W = 160, 431
    arr = rng.integers(0, 65536, (H, W, 6), dtype=np.uint16)
    arr.reshape(-1)[:65536 * 6] = np.repeat(np.arange(65536, dtype=np.uint16), 6)  # every DN in every band
    kernel, buff = 64, 32
    idx = otile.generate_chip_indices(arr.shape, buff, kernel)
    for spec, mm in ((processing.scalar_spec(6, 10000.0), None),
                     (processing.rescale_spec(6, moments=[(0, 10000), (3.0, 9000.5), (0.0, 8000.5), (1.0, 3000.0), (0.0, 65535.0), (200.0, 7000.0)]),
                      [(0, 10000), (3.0, 9000.5), (0.0, 8000.5), (1.0, 3000.0), (0.0, 65535.0), (200.0, 7000.0)])):
        got = _extract(arr, (kernel, buff), spec, idx)
        monkeypatch.setenv('SCV_K1_GENERIC', '1')
        generic = _extract(arr, (kernel, buff), spec, idx)
        monkeypatch.delenv('SCV_K1_GENERIC')
        assert np.array_equal(got, generic)
        for i, (y, x) in enumerate(idx):
            chip = arr[y - 16:y + 80, x - 16:x + 80, :].astype(np.float32)
            want = onorm.rescale_tensor(chip, moments=mm) if mm else chip / np.float32(10000.0)
            assert np.array_equal(got[i, :, :, :6], G.bf16_round(want)), i
            assert np.all(got[i, :, :, 6:] == 0)


def test_extract_pixel_and_tile_modes():
    rng = np.random.default_rng(4)
    arr = (rng.random((200, 216, 6)) * 5000 + 100).astype(np.float32)
    kernel, buff = 64, 32
    idx = otile.generate_chip_indices(arr.shape, buff, kernel)
    chips = [arr[y - 16:y + 80, x - 16:x + 80, :] for y, x in idx]
    got = _extract(arr, (kernel, buff), processing.rescale_spec(6), idx)  # reference default axes=[2]
    for i, c in enumerate(chips):
        assert np.array_equal(got[i, ..., :6], G.bf16_round(onorm.rescale_tensor(c)))
    got = _extract(arr, (kernel, buff), processing.normalize_spec(6), idx)
    for i, c in enumerate(chips):
        np.testing.assert_allclose(got[i, ..., :6], onorm.normalize_tensor(c), rtol=2 ** -7, atol=1e-5)
    got = _extract(arr, (kernel, buff), processing.normalize_spec(6, axes=[0, 1]), idx)  # solar notebook form
    for i, c in enumerate(chips):
        np.testing.assert_allclose(got[i, ..., :6], onorm.normalize_tensor(c, axes=(0, 1)), rtol=2 ** -7, atol=2e-3)
    got = _extract(arr, (kernel, buff), processing.rescale_spec(6, axes=[0, 1]), idx)
    for i, c in enumerate(chips):
        np.testing.assert_allclose(got[i, ..., :6], onorm.rescale_tensor(c, axes=(0, 1)), rtol=2 ** -7, atol=1e-5)
