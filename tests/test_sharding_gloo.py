"""N > 1 host logic on CPU: world_size-2 (and 3) gloo process groups exercise the row-band sharding, the
optional mosaic gather and the max-over-ranks reduction that bench.py uses -- with a CPU stand-in for the
per-band device call (the stand-in is NOT a product path; it only fills each band's output rows)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import tiling as otile
from satellite_computervision_b200 import sharding


def test_split_rows_partitions_exactly():
    assert sharding.split_rows(42, 8) == [(0, 6), (6, 12), (12, 17), (17, 22), (22, 27), (27, 32), (32, 37), (37, 42)]
    for n in (0, 1, 5, 42, 43):
        for w in (1, 2, 3, 8):
            parts = sharding.split_rows(n, w)
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            assert max(b - a for a, b in parts) - min(b - a for a, b in parts) <= 1


def test_bands_match_reference_chip_grid():
    H = W = 10980
    idx = otile.generate_chip_indices((H, W, 6), 128, 256)
    bands = [sharding.rank_band(H, W, 256, 128, r, 8) for r in range(8)]
    assert sum(b.n_chips for b in bands) == len(idx) == 1764
    rows = sorted({y for y, _ in idx})
    for b in bands:
        mine = rows[b.tile_row_begin:b.tile_row_end]
        assert b.src_row0 == mine[0] - 64 and b.src_row1 == mine[-1] + 320      # rows the chips read
        assert b.dst_row0 == mine[0] and b.dst_row1 == mine[-1] + 256            # rows the cores write
        assert (b.dst_col0, b.dst_col1) == (64, 64 + 42 * 256)
    assert all(a.dst_row1 == b.dst_row0 for a, b in zip(bands, bands[1:]))       # disjoint, contiguous
    assert sharding.rank_band_from(bands[3], 5, H, W) == bands[5]
    empty = sharding.rank_band(500, 500, 256, 128, 1, 2)  # one tile row only: rank 1 has nothing
    assert empty.n_chips == 0 and sharding.rank_band(500, 500, 256, 128, 0, 2).n_chips == 1


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, H, W, kernel, buff, out_path):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)  # every rank generates the same scene and keeps its band
        scene = rng.random((H, W, 3)).astype(np.float32)
        band = sharding.rank_band(H, W, kernel, buff, rank, world)
        local = scene[band.src_row0:band.src_row1]
        # stand-in for the device call: per-pixel mean over bands, written to the kept cores only
        rows = np.zeros((band.dst_row1 - band.dst_row0, W), np.float32)
        off = band.dst_row0 - band.src_row0
        rows[:, band.dst_col0:band.dst_col1] = local[off:off + rows.shape[0], band.dst_col0:band.dst_col1].mean(-1)
        full = sharding.gather_mosaic(torch.from_numpy(rows), band, H, W, dst=0)
        worst = sharding.max_over_ranks(10.0 * (rank + 1))
        assert worst == 10.0 * world
        if rank == 0:
            np.save(out_path, full.numpy())
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_gloo_row_band_sharding_and_gather(tmp_path, world):
    H, W, kernel, buff = 420, 300, 32, 16
    out = str(tmp_path / 'full.npy')
    mp.spawn(_worker, args=(world, _free_port(), H, W, kernel, buff, out), nprocs=world, join=True)
    full = np.load(out)
    scene = np.random.default_rng(0).random((H, W, 3)).astype(np.float32)
    idx = otile.generate_chip_indices((H, W, 3), buff, kernel)
    want = otile.predict_chips(scene, idx, np.zeros((H, W)), lambda b: b.mean(-1, keepdims=True), kernel, buff)
    assert np.array_equal(full, want.astype(np.float32))


def test_tile_balanced_shards_partition_the_chip_list():
    """SURVEY 7 hard part 7: 1764 chips over 8 ranks = 220 / 221 each (whole tile rows would give 252 / 210)."""
    H = W = 10980
    idx = otile.generate_chip_indices((H, W, 6), 128, 256)
    shards = [sharding.rank_shard(H, W, 256, 128, r, 8) for r in range(8)]
    assert [s.n_chips for s in shards] == [220, 221, 220, 221, 220, 221, 220, 221]
    assert shards[0].tile_begin == 0 and shards[-1].tile_end == len(idx)
    assert all(a.tile_end == b.tile_begin for a, b in zip(shards, shards[1:]))
    cover = np.zeros((H // 4, W // 4), np.int8)  # rectangles are multiples of 64: check on a 4x coarser grid
    for s in shards:
        chips = idx[s.tile_begin:s.tile_end]
        assert s.src_row0 == chips[0][0] - 64 and s.src_row1 == chips[-1][0] + 320
        assert s.dst_row0 == chips[0][0] and s.dst_row1 == chips[-1][0] + 256
        rects = sharding.shard_rects(s, H, W, 128)
        assert sum((y1 - y0) * (x1 - x0) for y0, y1, x0, x1 in rects) == s.n_chips * 256 * 256
        for y0, y1, x0, x1 in rects:
            cover[y0 // 4:y1 // 4, x0 // 4:x1 // 4] += 1
    assert cover.max() == 1 and cover[16:16 + 42 * 64, 16:16 + 42 * 64].min() == 1 and cover.sum() == (42 * 64) ** 2
    one = sharding.rank_shard(500, 500, 256, 128, 1, 2)  # a single chip: rank 1 has nothing
    assert one.n_chips + sharding.rank_shard(500, 500, 256, 128, 0, 2).n_chips == 1


def _shard_worker(rank, world, port, H, W, kernel, buff, out_path):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        scene = np.random.default_rng(0).random((H, W, 3)).astype(np.float32)
        sh = sharding.rank_shard(H, W, kernel, buff, rank, world)
        rows = np.zeros((sh.dst_row1 - sh.dst_row0, W), np.float32)
        # stand-in for the device call (NOT a product path): per-pixel band mean written to this shard's cores only
        for y0, y1, x0, x1 in sharding.shard_rects(sh, H, W, buff):
            rows[y0 - sh.dst_row0:y1 - sh.dst_row0, x0:x1] = scene[y0:y1, x0:x1].mean(-1)
        full = sharding.gather_shards(torch.from_numpy(rows), sh, H, W, buff, dst=0)
        if rank == 0:
            np.save(out_path, full.numpy())
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_gloo_tile_balanced_sharding_and_batched_gather(tmp_path, world):
    H, W, kernel, buff = 420, 300, 32, 16
    out = str(tmp_path / 'full.npy')
    mp.spawn(_shard_worker, args=(world, _free_port(), H, W, kernel, buff, out), nprocs=world, join=True)
    full = np.load(out)
    scene = np.random.default_rng(0).random((H, W, 3)).astype(np.float32)
    idx = otile.generate_chip_indices((H, W, 3), buff, kernel)
    want = otile.predict_chips(scene, idx, np.zeros((H, W)), lambda b: b.mean(-1, keepdims=True), kernel, buff)
    assert np.array_equal(full, want.astype(np.float32))
