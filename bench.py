#!/usr/bin/env python
"""bench.py -- megapixels/sec of tiled U-Net inference on a synthetic 10980x10980x6 Sentinel-2 scene.

    python bench.py --gpus N --steps K --warmup W            # this repo's engine (CUDA, sm_100a)
    python bench.py --impl reference --gpus N ...            # the reference algorithm on the host CPUs

One "step" = one pass of the hot path over the whole scene: buffered chip gather + per-band
normalise (K1) -> U-Net (tcgen05 implicit-GEMM convs) -> sigmoid/threshold + crop + stitch (K4).
`value` is timed with the scene band already resident in HBM (CUDA events on the launching stream,
max over ranks); `e2e` goes through the public API with pinned HOST buffers (H2D of the scene and D2H
of the stitched probability + mask rasters inside the timed region).  Multi-GPU: one process per GPU,
tile rows sharded across ranks, no collective on the data path ("weak" is not claimed: total work is
fixed, scaling = "strong").

The reference has no published throughput -> vs_baseline is null.  TensorFlow is not
installable in this image, so the reference arm times the oracle port (oracle/: torch-CPU fp32
restatement of model.predict + predict_chips) on all host cores, on a bounded sample of the same tiles.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCENE = 10980
BANDS = 6
KERNEL, BUFF = 256, 128
METRIC = 'megapixels/sec tiled U-Net inference'
UNIT = 'MP/s'


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return dict(hbm=d['hbm_gbs'], tc_burst=d['bf16_tflops'], tc_sustained=d.get('bf16_tflops_sustained', d['bf16_tflops']),
                    source='measured (MEASURED_PEAKS.json)')
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, source='fallback (B200_PROFILING.md)')


def make_scene(h, w, seed=1):
    """Synthetic Sentinel-2 L2A digital numbers: uniform noise modulated by a
    low-frequency field so the masks are not trivial."""
    rng = np.random.default_rng(seed)
    dn = rng.integers(0, 10000, (h, w, BANDS), dtype=np.uint16)
    yy = np.linspace(0, 6 * np.pi, h, dtype=np.float32)[:, None]
    xx = np.linspace(0, 4 * np.pi, w, dtype=np.float32)[None, :]
    field = (0.6 + 0.4 * np.sin(yy) * np.cos(xx)).astype(np.float32)
    return (dn * field[..., None]).astype(np.uint16)


def random_weights(model, seed=0):
    """Random-init weights of the BASELINE architecture: keras glorot-uniform kernels (the model's own
    init) with randomised BatchNorm statistics so the BN fold is exercised."""
    rng = np.random.default_rng(seed)
    out = []
    for name, w in zip(model.weight_names, model.get_weights()):
        leaf = name.rsplit('/', 1)[1]
        if leaf == 'gamma':
            w = rng.uniform(0.5, 1.5, w.shape)
        elif leaf in ('beta', 'moving_mean'):
            w = rng.normal(0, 0.1, w.shape)
        elif leaf == 'moving_variance':
            w = rng.uniform(0.5, 1.5, w.shape)
        elif leaf == 'bias':
            w = rng.normal(0, 0.05, w.shape)
        out.append(np.asarray(w, np.float32))
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index):
        self.gpu, self.lines, self.proc = gpu_index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                          '-lms', '100', '-i', str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for ln in self.lines:
            f = [s.strip() for s in ln.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': float(max(mx)), 'reasons': sorted(reasons),
                'samples': len(sm)}


def cpu_reference_sample(n_tiles, seed=1, variant='A'):
    """The reference algorithm (generate_chip_indices + per-tile batch-1 predict + crop/stitch) on the
    host cores via the oracle port; returns (MP/s in scene-equivalent pixels, description, cores)."""
    import torch

    from oracle import normalize as onorm
    from oracle import tiling as otile
    from oracle import unet as ounet
    specs = ounet.weight_specs(variant, BANDS, 1)
    w = ounet.init_weights(specs, seed=0)
    cols = max(1, n_tiles)
    width = 64 + KERNEL * cols + 192  # exactly `cols` chips in one tile row
    arr = make_scene(KERNEL + BUFF + 1 + 64, width, seed)[:KERNEL + BUFF + 65]
    x = onorm.rescale_tensor(arr.astype(np.float32), moments=[(0, 10000)] * BANDS)
    idx = otile.generate_chip_indices(x.shape, BUFF, KERNEL)
    idx = idx[:n_tiles]
    fn = ounet.make_predict_fn(w, variant=variant)
    fn(x[None, :384, :384])  # warm-up (thread pool, oneDNN primitives)
    t0 = time.perf_counter()
    otile.predict_chips(x, idx, np.zeros(x.shape[:2]), fn, KERNEL, BUFF)
    dt = time.perf_counter() - t0
    # scene-equivalent pixels: the full scene has 1764 chips for 120.56 MP
    px_per_chip = SCENE * SCENE / 1764.0
    mps = len(idx) * px_per_chip / dt / 1e6
    return mps, f'{len(idx)} of 1764 chips (384x384x6, batch-1 predict + crop/stitch), {dt:.2f} s', torch.get_num_threads()


def run_reference(args, rank):
    if rank != 0:
        return
    vals = []
    for _ in range(args.warmup):
        cpu_reference_sample(2)
    t_all = time.perf_counter()
    desc, cores = '', 1
    for _ in range(args.steps):
        v, desc, cores = cpu_reference_sample(args.ref_tiles)
        vals.append(v)
    ms = (time.perf_counter() - t_all) * 1e3 / max(1, args.steps)
    v = float(np.mean(vals))
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'synthetic {SCENE}x{SCENE}x{BANDS} uint16 Sentinel-2 scene, variant-A U-Net, '
                               f'{KERNEL}px kernel + {BUFF}px buffer; bounded sample per step, extrapolated per chip'},
        'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': desc},
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'note': 'TensorFlow/Keras not installable in this image: oracle port (torch-CPU fp32) of the reference path',
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--scene', type=int, default=SCENE)
    ap.add_argument('--max-batch', type=int, default=126)
    ap.add_argument('--ref-tiles', type=int, default=24)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--profile-layers', action='store_true')
    ap.add_argument('--gather', action='store_true', help='also time the optional NCCL gather of the bands to rank 0')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if args.impl == 'reference':
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    from satellite_computervision_b200 import _lib, model_tools, processing
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (no CPU fallback)')
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    H = W = args.scene
    model = model_tools.binary_unet(nchannels=BANDS, device=local_rank, max_batch=args.max_batch, outputs='probs')
    model.set_weights(random_weights(model, seed=0))
    spec = processing.rescale_spec(BANDS, moments=[(0, 10000)] * BANDS)
    lib = model._lib
    eng = model._ensure_engine()
    if args.profile_layers:
        model.set_option('profile_layers', 1)

    # chip grid and this rank's tile rows (row-band sharding, no data-path collective)
    from satellite_computervision_b200 import sharding
    half, side = BUFF // 2, KERNEL + BUFF
    ys, xs = sharding.chip_grid(H, W, KERNEL, BUFF)
    band_info = sharding.rank_band(H, W, KERNEL, BUFF, rank, world)
    r0, r1 = band_info.tile_row_begin, band_info.tile_row_end
    n_chips_total = len(ys) * len(xs)
    src_row0, src_row1 = band_info.src_row0, band_info.src_row1
    dst_row0, dst_rows = band_info.dst_row0, band_info.dst_row1 - band_info.dst_row0

    # every rank generates the same scene and keeps its band (pinned host memory for the e2e leg)
    scene = make_scene(H, W, seed=1)
    band = _lib.pinned_empty((src_row1 - src_row0, W, BANDS), np.uint16)
    band[...] = scene[src_row0:src_row1]
    del scene
    h_prob = _lib.pinned_empty((H, W), np.float32)
    h_mask = _lib.pinned_empty((H, W), np.uint8)
    h_prob[...] = 0
    h_mask[...] = 0

    d_band = torch.from_numpy(band.view(np.int16)).cuda()
    d_prob = torch.zeros((dst_rows, W), dtype=torch.float32, device='cuda')
    d_mask = torch.zeros((dst_rows, W), dtype=torch.uint8, device='cuda')
    import ctypes as C
    tiling = _lib.Tiling(KERNEL, BUFF)
    cn = spec.to_c(BANDS)
    stream = torch.cuda.current_stream()

    def step_device():
        _lib.check(lib.scv_predict_mosaic_device(eng, C.c_void_p(d_band.data_ptr()), _lib.SCV_U16, H, W, BANDS, src_row0,
                                                 C.byref(tiling), C.byref(cn), r0, r1, 0,
                                                 C.c_void_p(d_prob.data_ptr()), C.c_void_p(d_mask.data_ptr()), dst_row0,
                                                 C.c_void_p(stream.cuda_stream)))

    def step_e2e():
        # public C-ABI call, host buffers: H2D of the band + D2H of the prob/mask rasters happen inside.
        # The call takes the mosaic base pointer and only touches rows [src_row0, src_row1): this rank
        # holds just its band, so the base is formed by pointer arithmetic and never dereferenced.
        _lib.check(lib.scv_predict_mosaic(eng, C.c_void_p(band.ctypes.data - src_row0 * W * BANDS * 2), _lib.SCV_U16, H, W,
                                          BANDS, C.byref(tiling), C.byref(cn), r0, r1, 0, _lib.ptr(h_prob), _lib.ptr(h_mask)))

    # ---------------- device-resident leg (value)
    for _ in range(args.warmup):
        step_device()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clk = ClockSampler(local_rank)
    clk.__enter__()  # sampled over both timed legs (device-resident and end-to-end): short multi-GPU steps still get samples
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    torch.cuda.synchronize()
    barrier()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device='cuda')
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    ms_step = ms_total / args.steps
    times = model.times()  # last step on this rank
    mp_scene = H * W / 1e6
    value = mp_scene / (ms_step / 1e3)

    # ---------------- end-to-end leg (host buffers through the C-ABI)
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    t_e2e = torch.tensor([(time.perf_counter() - t0) * 1e3], device='cuda')
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    clk.__exit__(None, None, None)
    e2e_ms = float(t_e2e.item()) / args.steps
    h2d = (src_row1 - src_row0) * W * BANDS * 2
    d2h = dst_rows * len(xs) * KERNEL * 5
    if world > 1:
        tot = torch.tensor([h2d, d2h], dtype=torch.float64, device='cuda')
        dist.all_reduce(tot)
        h2d, d2h = int(tot[0].item()), int(tot[1].item())

    gather_ms = None
    if args.gather and world > 1:
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        full = sharding.gather_mosaic(d_prob, band_info, H, W, dst=0)
        g1.record()
        torch.cuda.synchronize()
        gather_ms = sharding.max_over_ranks(g0.elapsed_time(g1), device='cuda')
        del full

    if rank == 0:
        pk = peaks()
        traffic = None  # DRAM bytes of the conv launches of one step: ncu-measured bytes per chip x chips of rank 0
        tp = os.path.join(ROOT, 'profiles', 'r01_k_conv_traffic.json')
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f)['dram_bytes_per_chip'] * (r1 - r0) * len(xs)
        n_my = (r1 - r0) * len(xs)
        flops_tile = sum(times['layer_flops'])
        net_s = times['network_ms'] / 1e3
        tc = flops_tile * n_my / net_s / 1e12 if net_s > 0 else 0.0
        ex_bytes = ((ys[r1 - 1] + KERNEL + half - (ys[r0] - half)) * (xs[-1] + KERNEL + half - (xs[0] - half)) * BANDS * 2
                    + n_my * side * side * 8 * 2)
        st_bytes = n_my * KERNEL * KERNEL * 9
        ex_gbs = ex_bytes / (times['extract_ms'] / 1e3) / 1e9 if times['extract_ms'] > 0 else 0.0
        st_gbs = st_bytes / (times['stitch_ms'] / 1e3) / 1e9 if times['stitch_ms'] > 0 else 0.0
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'bf16',
            'data': 'synthetic',
            'config': {'workload': f'synthetic {H}x{W}x{BANDS} uint16 Sentinel-2 scene (BASELINE configs[1]), variant-A U-Net '
                                   f'(31.1 M params, BN folded), {KERNEL}px kernel + {BUFF}px buffer, {n_chips_total} chips, '
                                   f'sharded by tile rows over {world} GPU(s)',
                       'tiles_per_batch': args.max_batch, 'stitched_megapixels': n_chips_total * KERNEL * KERNEL / 1e6,
                       'scene_megapixels': mp_scene,
                       'l2': 'inputs larger than L2 (scene band 723 MB/N, activations > 126 MB per batch); no explicit flush'},
            'e2e': {'value': mp_scene / (e2e_ms / 1e3), 'unit': UNIT, 'ms_per_step': e2e_ms, 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h},
            'gpu_launches': int(times['n_launches']) * args.steps,  # rank 0's kernels in the timed device-resident region
            'clocks': clk.summary(),
            'roofline': {'bound': 'tensor', 'kernel': 'conv_umma_kernel (all conv/convT layers of the U-Net)',
                         'achieved': tc, 'peak': pk['tc_sustained'], 'unit': 'TFLOP/s',
                         'frac': tc / pk['tc_sustained'] if pk['tc_sustained'] else None,
                         'frac_of_burst_peak': tc / pk['tc_burst'], 'peak_burst': pk['tc_burst'], 'traffic': traffic,
                         'traffic_source': 'profiles/r01_k_conv_traffic.json (ncu dram__bytes_read+write, per chip)',
                         'peak_source': pk['source'],
                         'how': 'algorithmic FLOPs (67.41 GFLOP per 384x384x6 chip) x chips of rank 0 / CUDA-event time of '
                                'the conv launches of the last timed step'},
            'roofline_extract': {'bound': 'hbm', 'achieved': ex_gbs, 'peak': pk['hbm'], 'unit': 'GB/s',
                                 'frac': ex_gbs / pk['hbm'], 'ms': times['extract_ms']},
            'roofline_stitch': {'bound': 'hbm', 'achieved': st_gbs, 'peak': pk['hbm'], 'unit': 'GB/s',
                                'frac': st_gbs / pk['hbm'], 'ms': times['stitch_ms']},
            'stage_ms_last_step_rank0': {k: times[k] for k in ('total_ms', 'extract_ms', 'network_ms', 'stitch_ms')},
        }
        if gather_ms is not None:
            line['optional_gather_ms'] = gather_ms
        if args.profile_layers:
            line['layers'] = [{'name': n, 'ms': m, 'tflops': (f * n_my / (m / 1e3) / 1e12 if m > 0 else None)}
                              for n, m, f in zip(layer_names(model), times['layer_ms'], times['layer_flops'])]
        if world == 1 and not args.no_cpu_baseline:
            v, desc, cores = cpu_reference_sample(args.ref_tiles)
            line['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': desc}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def layer_names(model):
    names = []
    L = len(model.filters)
    nconv = 2 if model.double_conv else 1
    for i in range(L):
        names += [f'encoder_{i}/conv{j}' for j in range(nconv)]
    names += [f'center/conv{j}' for j in range(nconv)]
    for i in range(L - 1, -1, -1):
        names += [f'decoder_{i}/up', f'decoder_{i}/conv0', f'decoder_{i}/conv1']
    return names


if __name__ == '__main__':
    main()
