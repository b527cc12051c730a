"""Row-band sharding of the chip grid across ranks: tiles are independent units with read-only
overlap, so rank r takes a contiguous range of TILE ROWS, needs only the mosaic rows those chips read (a
buff/2-row halo each side, taken from the host scene, never from peers) and writes a disjoint range of output
rows.  No collective on the data path; `gather_mosaic` is the one optional collective (bands -> rank 0).

Pure integer logic + thin torch.distributed helpers (torch is imported lazily: the core package does not need it).
"""
from __future__ import annotations

from collections import namedtuple

Band = namedtuple('Band', 'rank world tile_row_begin tile_row_end n_tile_cols src_row0 src_row1 dst_row0 dst_row1 '
                          'dst_col0 dst_col1 n_chips')
# tile-balanced shard: chips [tile_begin, tile_end) of the row-major chip list; the tile rows it touches may be partial
Shard = namedtuple('Shard', 'rank world tile_begin tile_end n_tile_cols tile_row_begin tile_row_end src_row0 src_row1 '
                            'dst_row0 dst_row1 n_chips kernel')


def chip_grid(H, W, kernel=256, buff=128):
    """y and x coordinates of generate_chip_indices (utils/prediction_tools.py:105-106)."""
    side, half = kernel + buff, buff // 2
    return list(range(half, H - side, kernel)), list(range(half, W - side, kernel))


def split_rows(n_rows, world):
    """Contiguous, as-even-as-possible ranges: the first n_rows % world ranks get one extra row."""
    base, extra = divmod(n_rows, world)
    out, r = [], 0
    for i in range(world):
        n = base + (1 if i < extra else 0)
        out.append((r, r + n))
        r += n
    return out


def rank_band(H, W, kernel, buff, rank, world):
    """Everything rank `rank` needs to know about its share of an (H, W) scene."""
    ys, xs = chip_grid(H, W, kernel, buff)
    r0, r1 = split_rows(len(ys), world)[rank]
    half, side = buff // 2, kernel + buff
    if r1 > r0 and xs:
        src0, src1 = ys[r0] - half, ys[r1 - 1] - half + side
        dst0, dst1 = ys[r0], ys[r1 - 1] + kernel
        c0, c1 = xs[0], xs[-1] + kernel
    else:
        src0 = src1 = dst0 = dst1 = c0 = c1 = 0
    return Band(rank, world, r0, r1, len(xs), src0, src1, dst0, dst1, c0, c1, (r1 - r0) * len(xs))


def rank_shard(H, W, kernel, buff, rank, world):
    """Tile-balanced sharding (hard part 7): rank r takes chips [r*n/world, (r+1)*n/world) of the
    row-major chip list, so every rank gets n/world chips +-1 (1764 chips over 8 ranks: 220 / 221) instead of
    whole tile rows (6,6,5,5,5,5,5,5 rows = 252 vs 210 chips).  src / dst rows are the mosaic rows read and the
    output rows written by the tile rows the range touches (the first and last may be shared with a neighbour,
    each rank writing only its own cores)."""
    ys, xs = chip_grid(H, W, kernel, buff)
    n, ncols = len(ys) * len(xs), len(xs)
    t0, t1 = (n * rank) // world, (n * (rank + 1)) // world
    half, side = buff // 2, kernel + buff
    if t1 > t0:
        r0, r1 = t0 // ncols, (t1 - 1) // ncols + 1
        src0, src1 = ys[r0] - half, ys[r1 - 1] - half + side
        dst0, dst1 = ys[r0], ys[r1 - 1] + kernel
    else:
        r0 = r1 = src0 = src1 = dst0 = dst1 = 0
    return Shard(rank, world, t0, t1, ncols, r0, r1, src0, src1, dst0, dst1, t1 - t0, kernel)


def shard_rects(shard, H, W, buff):
    """Output rectangles (y0, y1, x0, x1) a shard writes: one per tile row it touches (cores only)."""
    ys, xs = chip_grid(H, W, shard.kernel, buff)
    out = []
    for r in range(shard.tile_row_begin, shard.tile_row_end):
        c0 = shard.tile_begin - r * shard.n_tile_cols if r == shard.tile_row_begin else 0
        c1 = shard.tile_end - r * shard.n_tile_cols if r == shard.tile_row_end - 1 else shard.n_tile_cols
        c0, c1 = max(c0, 0), min(c1, shard.n_tile_cols)
        if c1 > c0:
            out.append((ys[r], ys[r] + shard.kernel, xs[c0], xs[c1 - 1] + shard.kernel))
    return out


def gather_shards(rows, shard, H, W, buff, dst=0, group=None):
    """Optional final collective for tile-balanced shards: every rank's output rows [dst_row0, dst_row1) (a
    (rows, W) tensor holding its cores) are gathered on rank `dst` and its rectangles pasted into the (H, W)
    raster.  All transfers are posted as ONE batch of point-to-point operations (`batch_isend_irecv` =
    ncclGroupStart / ncclSend / ncclRecv / ncclGroupEnd on NCCL; NVSwitch gives every peer full bandwidth, so a
    flat gather is optimal -- SURVEY 8(e))."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if rank != dst:
        if shard.n_chips > 0:
            for w in dist.batch_isend_irecv([dist.P2POp(dist.isend, rows.contiguous(), dst, group)]):
                w.wait()
        return None
    full = torch.zeros((H, W), dtype=rows.dtype, device=rows.device)
    bufs, ops = {}, []
    for r in range(world):
        sh = rank_shard(H, W, shard.kernel, buff, r, world)
        if sh.n_chips == 0:
            continue
        if r == dst:
            bufs[r] = (sh, rows)
            continue
        buf = torch.empty((sh.dst_row1 - sh.dst_row0, W), dtype=rows.dtype, device=rows.device)
        bufs[r] = (sh, buf)
        ops.append(dist.P2POp(dist.irecv, buf, r, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    for r, (sh, buf) in bufs.items():
        for y0, y1, x0, x1 in shard_rects(sh, H, W, buff):
            full[y0:y1, x0:x1] = buf[y0 - sh.dst_row0:y1 - sh.dst_row0, x0:x1]
    return full


def gather_mosaic(band_rows, band, H, W, dst=0, group=None):
    """Optional final collective: every rank sends its output rows [dst_row0, dst_row1) (a (rows, W) tensor)
    to rank `dst`, which returns the assembled (H, W) raster (zeros outside all bands); other ranks get None.
    Works with any torch.distributed backend (NCCL on GPU tensors, gloo on CPU tensors)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if rank != dst:
        if band.dst_row1 > band.dst_row0:
            dist.send(band_rows.contiguous(), dst, group=group)
        return None
    full = torch.zeros((H, W), dtype=band_rows.dtype, device=band_rows.device)
    full[band.dst_row0:band.dst_row1] = band_rows
    for r in range(world):
        if r == dst:
            continue
        b = rank_band_from(band, r, H, W)
        if b.dst_row1 > b.dst_row0:
            buf = torch.empty((b.dst_row1 - b.dst_row0, W), dtype=band_rows.dtype, device=band_rows.device)
            dist.recv(buf, r, group=group)
            full[b.dst_row0:b.dst_row1] = buf
    return full


def rank_band_from(band, rank, H, W):
    """The band of another rank of the same scene/tiling, derived from this rank's band geometry."""
    n_rows_here = band.tile_row_end - band.tile_row_begin
    if n_rows_here <= 0:
        raise ValueError('cannot infer the tiling from an empty band')
    kernel = (band.dst_row1 - band.dst_row0) // n_rows_here
    buff = (band.src_row1 - band.src_row0) - (band.dst_row1 - band.dst_row0)
    return rank_band(H, W, kernel, buff, rank, band.world)


def max_over_ranks(value, device=None, group=None):
    """max over ranks of a per-rank scalar (device time in ms); identity without a process group."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
