#!/usr/bin/env python
"""bench.py -- megapixels/sec of tiled U-Net inference on a synthetic 10980x10980x6 Sentinel-2 scene.

    python bench.py --gpus N --steps K --warmup W            # this repo's engine (CUDA, sm_100a)
    python bench.py --impl reference --gpus N ...            # the reference algorithm on the host CPUs
    python bench.py --gpus N --scenes 64 [--stream-mode bands]   # BASELINE configs[4]: sustained multi-scene streaming

One "step" = one pass of the hot path over the whole scene: buffered chip gather + per-band
normalise (K1) -> U-Net (tcgen05 implicit-GEMM convs) -> sigmoid/threshold + crop + stitch (K4).
`value` is timed with the scene shard already resident in HBM (CUDA events on the launching stream,
max over ranks); `e2e` goes through the public C-ABI call with pinned HOST buffers (H2D of the scene and D2H
of the stitched probability + mask rasters inside the timed region).  Multi-GPU: one process per GPU, the
row-major chip list split evenly across ranks (220 / 221 chips each at N = 8), no collective on the data path
("weak" is not claimed: total work is fixed, scaling = "strong").  After the timed legs `--verify` (default on)
compares chips of every rank's output with the CPU oracle and fails the run on a mismatch.

The reference has no published throughput -> vs_baseline is null.  TensorFlow is not
installable in this image, so the reference arm times the oracle port (oracle/: torch-CPU fp32
restatement of model.predict + predict_chips) on all host cores, on a bounded sample of the same tiles;
if TensorFlow does import on the box, the line says so and pins the oracle against real Keras.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SCENE = 10980
BANDS = 6
KERNEL, BUFF = 256, 128
METRIC = 'megapixels/sec tiled U-Net inference'
UNIT = 'MP/s'


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def use_all_host_threads(share=1):
    """torch.distributed.run exports OMP_NUM_THREADS=1 to every rank: undo that for the CPU legs (before torch
    is imported) so the CPU baseline uses the cores it is reported with."""
    n = max(1, host_cores() // max(1, share))
    os.environ['OMP_NUM_THREADS'] = str(n)
    os.environ['MKL_NUM_THREADS'] = str(n)
    import torch
    torch.set_num_threads(n)
    return n


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
    import numpy as np
    rng = np.random.default_rng(seed)
    dn = rng.integers(0, 10000, (h, w, BANDS), dtype=np.uint16)
    yy = np.linspace(0, 6 * np.pi, h, dtype=np.float32)[:, None]
    xx = np.linspace(0, 4 * np.pi, w, dtype=np.float32)[None, :]
    field = (0.6 + 0.4 * np.sin(yy) * np.cos(xx)).astype(np.float32)
    return (dn * field[..., None]).astype(np.uint16)


def random_weights(model, seed=0):
    """Random-init weights of the BASELINE architecture: keras glorot-uniform kernels (the model's own
    init) with randomised BatchNorm statistics so the BN fold is exercised."""
    import numpy as np
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
        if self.gpu < 0:
            return self
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
        import numpy as np
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


def cpu_reference_sample(n_tiles, seed=1, variant='A', arch='unet'):
    """The reference algorithm (generate_chip_indices + per-tile batch-1 predict + crop/stitch) on the
    host cores via the oracle port; returns (MP/s in scene-equivalent pixels, description, cores)."""
    import numpy as np
    import torch

    from oracle import normalize as onorm
    from oracle import tiling as otile
    from oracle import unet as ounet
    if arch == 'siamese':
        from oracle import siamese as osi
        specs = osi.weight_specs(BANDS // 2)
    else:
        specs = ounet.weight_specs(variant, BANDS, 1)
    w = ounet.init_weights(specs, seed=0)
    cols = max(1, n_tiles)
    width = 64 + KERNEL * cols + 192  # exactly `cols` chips in one tile row
    arr = make_scene(KERNEL + BUFF + 1 + 64, width, seed)[:KERNEL + BUFF + 65]
    x = onorm.rescale_tensor(arr.astype(np.float32), moments=[(0, 10000)] * BANDS)
    idx = otile.generate_chip_indices(x.shape, BUFF, KERNEL)
    idx = idx[:n_tiles]
    fn = osi.make_predict_fn(w, BANDS // 2) if arch == 'siamese' else ounet.make_predict_fn(w, variant=variant)
    fn(x[None, :384, :384])  # warm-up (thread pool, oneDNN primitives)
    t0 = time.perf_counter()
    otile.predict_chips(x, idx, np.zeros(x.shape[:2]), fn, KERNEL, BUFF)
    dt = time.perf_counter() - t0
    # scene-equivalent pixels: the full scene has 1764 chips for 120.56 MP
    px_per_chip = SCENE * SCENE / 1764.0
    mps = len(idx) * px_per_chip / dt / 1e6
    return mps, (f'{len(idx)} of 1764 chips (384x384x6, batch-1 predict + crop/stitch), {dt:.2f} s, extrapolated per chip; '
                 f'{torch.get_num_threads()} torch threads'), torch.get_num_threads()


def run_reference(args):
    import numpy as np
    cores = use_all_host_threads()
    from oracle import keras_probe
    tfp = keras_probe.probe()
    vals = []
    for _ in range(args.warmup):
        cpu_reference_sample(2)
    t_all = time.perf_counter()
    desc = ''
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
                               f'{KERNEL}px kernel + {BUFF}px buffer; bounded sample of {args.ref_tiles} chips per step, '
                               'extrapolated per chip to the 1764-chip scene (same chips, same network, same algorithm)'},
        'cpu_baseline': {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': desc},
        'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'tensorflow_probe': tfp,
        'note': 'TensorFlow/Keras not installable in this image: oracle port (torch-CPU fp32) of the reference path, '
                'rank 0 only, all host cores',
    }
    print(json.dumps(line), flush=True)


def verify_outputs(model, shard, scene_rows, src_row0, W, h_prob, h_mask, d_prob_host, dst_row0, n_chips, threads_share):
    """Compare `n_chips` chips spread over this rank's shard (first and last included) with the CPU oracle:
    probabilities of the e2e leg, the mask, and bit-equality of the device-resident leg with the e2e leg."""
    import numpy as np

    from oracle import normalize as onorm
    from oracle import unet as ounet
    use_all_host_threads(threads_share)
    ncols = shard.n_tile_cols
    picks = sorted({int(t) for t in np.linspace(shard.tile_begin, shard.tile_end - 1, n_chips)})
    if getattr(model, 'is_siamese', False):
        from oracle import siamese as osi
        fn = osi.make_predict_fn(model.get_weights(), BANDS // 2, filters=tuple(model.filters))
    else:
        fn = ounet.make_predict_fn(model.get_weights(), variant='A')
    half = BUFF // 2
    max_abs, agree_n, agree_d, dev_equal = 0.0, 0, 0, True
    for t in picks:
        y, x = shard.ys[t // ncols], shard.xs[t % ncols]
        chip = scene_rows[y - half - src_row0:y - half - src_row0 + KERNEL + BUFF, x - half:x - half + KERNEL + BUFF]
        xin = onorm.rescale_tensor(chip.astype(np.float32), moments=[(0, 10000)] * BANDS)
        ref = fn(xin[None])[0, half:half + KERNEL, half:half + KERNEL, 0]
        got = h_prob[y:y + KERNEL, x:x + KERNEL]
        max_abs = max(max_abs, float(np.abs(got - ref).max()))
        agree_n += int(((got > 0.5) == (ref > 0.5)).sum())
        agree_d += ref.size
        if not np.array_equal(h_mask[y:y + KERNEL, x:x + KERNEL], (got > 0.5).astype(np.uint8)):
            dev_equal = False
        if not np.array_equal(d_prob_host[y - dst_row0:y - dst_row0 + KERNEL, x:x + KERNEL], got):
            dev_equal = False
    return len(picks), max_abs, agree_n, agree_d, dev_equal


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
    ap.add_argument('--no-verify', action='store_true')
    ap.add_argument('--verify-chips', type=int, default=24, help='chips compared with the oracle (whole job)')
    ap.add_argument('--profile-layers', action='store_true')
    ap.add_argument('--shard', default='chips', choices=['chips', 'rows'], help='chips: tile-balanced chip ranges; rows: whole tile rows')
    ap.add_argument('--gather', action='store_true', help='also time the optional NCCL gather of the shards to rank 0')
    ap.add_argument('--python-api', action='store_true', help='also time prediction_tools.predict_chips (reference signature, '
                                                               'pageable arrays, float64 template) on the full scene (N = 1)')
    ap.add_argument('--arch', default='unet', choices=['unet', 'siamese'],
                    help="siamese: make_siamese_unet(3, [32, 64, 128]) on the same scene read as two stacked 3-band dates "
                         "(SURVEY 8(f) N4; a secondary workload, not the BASELINE metric's configuration)")
    ap.add_argument('--scenes', type=int, default=0, help='BASELINE configs[4]: stream this many scenes through the GPUs')
    ap.add_argument('--stream-mode', default='scenes', choices=['scenes', 'bands'],
                    help='scenes: round-robin whole scenes over the ranks; bands: every scene sharded over all ranks')
    args = ap.parse_args()

    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if args.impl == 'reference':
        if rank == 0:  # the other ranks exit before importing anything
            run_reference(args)
        return

    import ctypes as C

    import numpy as np
    import torch
    import torch.distributed as dist

    from satellite_computervision_b200 import _lib, model_tools, processing, sharding
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

    def reduce_max(v):
        t = torch.tensor([float(v)], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    H = W = args.scene
    if args.arch == 'siamese':
        model = model_tools.make_siamese_unet(BANDS // 2, device=local_rank, max_batch=args.max_batch, outputs='probs', seed=0)
        model.is_siamese = True
    else:
        model = model_tools.binary_unet(nchannels=BANDS, device=local_rank, max_batch=args.max_batch, outputs='probs', seed=0)
    model.set_weights(random_weights(model, seed=0))
    spec = processing.rescale_spec(BANDS, moments=[(0, 10000)] * BANDS)
    lib = model._lib
    eng = model._ensure_engine()
    if args.profile_layers:
        model.set_option('profile_layers', 1)
    tiling = _lib.Tiling(KERNEL, BUFF)
    cn = spec.to_c(BANDS)
    half, side = BUFF // 2, KERNEL + BUFF
    ys, xs = sharding.chip_grid(H, W, KERNEL, BUFF)
    n_chips_total = len(ys) * len(xs)
    mp_scene = H * W / 1e6

    def shard_of(r, w):
        if args.shard == 'rows':
            b = sharding.rank_band(H, W, KERNEL, BUFF, r, w)
            sh = sharding.Shard(r, w, b.tile_row_begin * len(xs), b.tile_row_end * len(xs), len(xs), b.tile_row_begin,
                                b.tile_row_end, b.src_row0, b.src_row1, b.dst_row0, b.dst_row1, b.n_chips, KERNEL)
        else:
            sh = sharding.rank_shard(H, W, KERNEL, BUFF, r, w)
        return sh

    if args.scenes > 0:
        return run_streaming(args, rank, local_rank, world, model, lib, eng, tiling, cn, shard_of, barrier, reduce_max)

    sh = shard_of(rank, world)

    class _S:  # shard + grid, for verify_outputs
        pass
    shv = _S()
    shv.tile_begin, shv.tile_end, shv.n_tile_cols, shv.ys, shv.xs = sh.tile_begin, sh.tile_end, sh.n_tile_cols, ys, xs
    src_row0, src_row1 = sh.src_row0, sh.src_row1
    dst_row0, dst_rows = sh.dst_row0, sh.dst_row1 - sh.dst_row0
    opts = _lib.MosaicOpts()
    opts.tile_begin, opts.tile_end = sh.tile_begin, sh.tile_end

    # every rank generates the same scene and keeps the rows its shard reads (pinned host memory for the e2e leg)
    scene = make_scene(H, W, seed=1)
    band = _lib.pinned_empty((src_row1 - src_row0, W, BANDS), np.uint16)
    band[...] = scene[src_row0:src_row1]
    if not (args.python_api and world == 1):
        del scene
    h_prob = _lib.pinned_empty((H, W), np.float32)
    h_mask = _lib.pinned_empty((H, W), np.uint8)
    h_prob[...] = 0
    h_mask[...] = 0

    d_band = torch.from_numpy(band.view(np.int16)).cuda()
    d_prob = torch.zeros((dst_rows, W), dtype=torch.float32, device='cuda')
    d_mask = torch.zeros((dst_rows, W), dtype=torch.uint8, device='cuda')
    stream = torch.cuda.current_stream()

    def step_device():
        _lib.check(lib.scv_predict_mosaic_device_ex(eng, C.c_void_p(d_band.data_ptr()), _lib.SCV_U16, H, W, BANDS, src_row0,
                                                    C.byref(tiling), C.byref(cn), C.byref(opts),
                                                    C.c_void_p(d_prob.data_ptr()), C.c_void_p(d_mask.data_ptr()), dst_row0,
                                                    C.c_void_p(stream.cuda_stream)))

    def step_e2e():
        # public C-ABI call, host buffers: H2D of the shard's rows + D2H of its prob/mask cores happen inside.
        # The call takes the mosaic base pointer and only touches rows [src_row0, src_row1): this rank
        # holds just those rows, so the base is formed by pointer arithmetic and never dereferenced.
        _lib.check(lib.scv_predict_mosaic_ex(eng, C.c_void_p(band.ctypes.data - src_row0 * W * BANDS * 2), _lib.SCV_U16, H, W,
                                             BANDS, C.byref(tiling), C.byref(cn), C.byref(opts), _lib.ptr(h_prob),
                                             _lib.ptr(h_mask)))

    # ---------------- device-resident leg (value)
    for _ in range(args.warmup):
        step_device()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clk = ClockSampler(local_rank if rank == 0 else -1)  # rank 0's GPU only: eight nvidia-smi pollers contend for the driver
    clk.__enter__()  # sampled over both timed legs (device-resident and end-to-end): short multi-GPU steps still get samples
    ev0.record(stream)
    for _ in range(args.steps):
        step_device()
    ev1.record(stream)
    torch.cuda.synchronize()
    barrier()
    _lib.check(lib.scv_check(eng))  # a watchdog trip inside the timed region must fail the run, not inflate it
    ms_step = reduce_max(ev0.elapsed_time(ev1)) / args.steps
    times = model.times()  # last step on this rank
    value = mp_scene / (ms_step / 1e3)

    # ---------------- end-to-end leg (host buffers through the C-ABI)
    for _ in range(max(1, args.warmup)):  # the same W as the device-resident leg (PCIe / copy engines warm up too)
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_local = (time.perf_counter() - t0) * 1e3 / args.steps
    e2e_ms = reduce_max(e2e_local * args.steps) / args.steps
    clk.__exit__(None, None, None)
    times_e2e = model.times()
    print(f'[rank {rank}] chips {sh.n_chips} device-leg kernels {times["total_ms"]:.2f} ms | e2e {e2e_local:.2f} ms = lead '
          f'{times_e2e["h2d_lead_ms"]:.2f} + kernels {times_e2e["total_ms"]:.2f} + tail {times_e2e["d2h_tail_ms"]:.2f} + host '
          f'{e2e_local - times_e2e["h2d_lead_ms"] - times_e2e["total_ms"] - times_e2e["d2h_tail_ms"]:.2f}', file=sys.stderr, flush=True)
    rects = sharding.shard_rects(sh, H, W, BUFF)
    h2d = (src_row1 - src_row0) * W * BANDS * 2  # upper bound: partial first / last tile rows copy fewer columns
    d2h = sum((y1 - y0) * (x1 - x0) for y0, y1, x0, x1 in rects) * 5
    if world > 1:
        tot = torch.tensor([h2d, d2h], dtype=torch.float64, device='cuda')
        dist.all_reduce(tot)
        h2d, d2h = int(tot[0].item()), int(tot[1].item())

    gather_ms = None
    if args.gather and world > 1:
        for _ in range(2):
            full = sharding.gather_shards(d_prob, sh, H, W, BUFF, dst=0)
            del full
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        full = sharding.gather_shards(d_prob, sh, H, W, BUFF, dst=0)
        g1.record()
        torch.cuda.synchronize()
        gather_ms = reduce_max(g0.elapsed_time(g1))
        del full

    # ---------------- verification against the oracle (after the timed legs, on their outputs)
    verify = None
    if not args.no_verify:
        d_prob_host = d_prob.cpu().numpy()
        k = max(3, -(-args.verify_chips // world))
        n, mx, an, ad, eq = verify_outputs(model, shv, band, src_row0, W, h_prob, h_mask, d_prob_host, dst_row0, k, world)
        v = torch.tensor([n, an, ad, 1.0 if eq else 0.0], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(v)
        mx = reduce_max(mx)
        agree = float(v[1].item() / max(1.0, v[2].item()))
        verify = {'chips': int(v[0].item()), 'max_abs': mx, 'mask_agree': agree, 'device_leg_equals_e2e_leg': bool(v[3].item() == world),
                  'bars': {'max_abs': 1e-2, 'mask_agree': 0.999}, 'checker': 'oracle port (torch-CPU fp32), chips spread over every rank\'s shard',
                  'ok': bool(mx <= 1e-2 and agree >= 0.999 and v[3].item() == world)}

    py_api = None
    if args.python_api and world == 1:
        from satellite_computervision_b200 import prediction_tools as pt
        idx = pt.generate_chip_indices(scene, BUFF, KERNEL)
        for _ in range(2):
            t0 = time.perf_counter()
            template = np.zeros((H, W))
            out = pt.predict_chips(scene, idx, template, model, KERNEL, BUFF, norm=spec)
            dt = time.perf_counter() - t0
        same = bool(np.array_equal(out.astype(np.float32), h_prob))
        py_api = {'value': mp_scene / dt, 'unit': UNIT, 'ms_per_step': dt * 1e3, 'equals_e2e_output': same,
                  'what': 'prediction_tools.predict_chips(scene uint16 pageable ndarray, indices, float64 zeros template, model): '
                          'reference signature; the library page-locks the caller\'s arrays for the call, the stitch kernel '
                          'accumulates into the float64 template on the device (template H2D + D2H inside)'}

    if rank == 0:
        pk = peaks()
        n_my = sh.n_chips
        traffic = None  # DRAM bytes of the conv launches of one step: ncu-measured bytes per chip x chips of rank 0
        for name in ('r02_conv_traffic.json', 'r01_k_conv_traffic.json'):
            tp = os.path.join(ROOT, 'profiles', name)
            if os.path.exists(tp) and args.arch == 'unet':
                with open(tp) as f:
                    traffic = json.load(f)['dram_bytes_per_chip'] * n_my
                traffic_src = f'profiles/{name} (ncu dram__bytes_read+write, per chip)'
                break
        flops_tile = sum(times['layer_flops'])
        net_s = times['network_ms'] / 1e3
        tc = flops_tile * n_my / net_s / 1e12 if net_s > 0 else 0.0
        r_lo, r_hi = sh.tile_row_begin, sh.tile_row_end
        ex_bytes = ((ys[r_hi - 1] + KERNEL + half - (ys[r_lo] - half)) * (xs[-1] + KERNEL + half - (xs[0] - half)) * BANDS * 2
                    * (n_my / ((r_hi - r_lo) * len(xs))) + n_my * side * side * 8 * 2)
        st_bytes = n_my * KERNEL * KERNEL * 9
        ex_gbs = ex_bytes / (times['extract_ms'] / 1e3) / 1e9 if times['extract_ms'] > 0 else 0.0
        st_gbs = st_bytes / (times['stitch_ms'] / 1e3) / 1e9 if times['stitch_ms'] > 0 else 0.0
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': 'bf16',
            'data': 'synthetic',
            'config': {'workload': (f'synthetic {H}x{W}x{BANDS} uint16 scene read as two stacked 3-band dates, siamese U-Net + ASPP '
                                    f'(make_siamese_unet(3, [32, 64, 128]), 2.4 M params; SECONDARY workload, SURVEY 8(f) N4), '
                                    if args.arch == 'siamese' else
                                    f'synthetic {H}x{W}x{BANDS} uint16 Sentinel-2 scene (BASELINE configs[1]), variant-A U-Net '
                                    f'(31.1 M params, BN folded), ') +
                                   f'{KERNEL}px kernel + {BUFF}px buffer, {n_chips_total} chips, '
                                   f'{"chip list split evenly" if args.shard == "chips" else "tile rows sharded"} over {world} GPU(s)',
                       'tiles_per_batch': args.max_batch, 'stitched_megapixels': n_chips_total * KERNEL * KERNEL / 1e6,
                       'scene_megapixels': mp_scene, 'chips_rank0': n_my,
                       'l2': 'inputs larger than L2 (scene shard 1.43 GB/N, activations > 126 MB per batch); no explicit flush'},
            'e2e': {'value': mp_scene / (e2e_ms / 1e3), 'unit': UNIT, 'ms_per_step': e2e_ms, 'h2d_bytes_per_step': h2d,
                    'd2h_bytes_per_step': d2h,
                    'pipeline_ms_last_step_rank0': {'h2d_lead': times_e2e['h2d_lead_ms'], 'kernels': times_e2e['total_ms'],
                                                    'd2h_tail': times_e2e['d2h_tail_ms']}},
            'gpu_launches': int(times['n_launches']) * args.steps,  # rank 0's kernels in the timed device-resident region
            'clocks': clk.summary(),
            'roofline': {'bound': 'tensor', 'kernel': ('tcgen05 implicit-GEMM conv kernels (all 21 conv/convT launches of the siamese U-Net + ASPP)' if args.arch == 'siamese'
                                                       else 'tcgen05 implicit-GEMM conv kernels (all 27 conv/convT layers of the U-Net)'),
                         'achieved': tc, 'peak': pk['tc_sustained'], 'unit': 'TFLOP/s',
                         'frac': tc / pk['tc_sustained'] if pk['tc_sustained'] else None,
                         'frac_of_burst_peak': tc / pk['tc_burst'], 'peak_burst': pk['tc_burst'], 'traffic': traffic,
                         'traffic_source': traffic_src if traffic else None,
                         'peak_source': pk['source'],
                         'how': f'algorithmic FLOPs ({flops_tile / 1e9:.2f} GFLOP per 384x384x6 chip) x chips of rank 0 / CUDA-event time of '
                                'the conv launches of the last timed step; peak = sustained cuBLAS bf16 (kernels timed inside a long step)'},
            'roofline_extract': {'bound': 'hbm', 'achieved': ex_gbs, 'peak': pk['hbm'], 'unit': 'GB/s',
                                 'frac': ex_gbs / pk['hbm'], 'ms': times['extract_ms']},
            'roofline_stitch': {'bound': 'hbm', 'achieved': st_gbs, 'peak': pk['hbm'], 'unit': 'GB/s',
                                'frac': st_gbs / pk['hbm'], 'ms': times['stitch_ms']},
            'stage_ms_last_step_rank0': {k: times[k] for k in ('total_ms', 'extract_ms', 'network_ms', 'stitch_ms')},
        }
        if verify is not None:
            line['verify'] = verify
        if gather_ms is not None:
            line['optional_gather_ms'] = gather_ms
        if py_api is not None:
            line['e2e_reference_signature'] = py_api
        if args.profile_layers:
            line['layers'] = [{'name': n, 'ms': m, 'tflops': (f * n_my / (m / 1e3) / 1e12 if m > 0 else None)}
                              for n, m, f in zip(layer_names(model), times['layer_ms'], times['layer_flops'])]
        if world == 1 and not args.no_cpu_baseline:
            use_all_host_threads()
            v, desc, cores = cpu_reference_sample(args.ref_tiles, arch=args.arch)
            line['cpu_baseline'] = {'value': v, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': desc}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if verify is not None and not verify['ok']:
        raise SystemExit(f'verification against the oracle FAILED: {verify}')


def run_streaming(args, rank, local_rank, world, model, lib, eng, tiling, cn, shard_of, barrier, reduce_max):
    """BASELINE configs[4]: `--scenes S` synthetic scenes streamed through the GPUs, H2D and D2H of every scene
    inside the timed region, two scenes in flight per GPU (scv_stream_submit / scv_stream_wait).
    mode 'scenes': whole scenes round-robin over the ranks (no communication at all);
    mode 'bands' : every scene sharded over all ranks (lower latency per scene, same total work)."""
    import ctypes as C

    import numpy as np
    import torch
    import torch.distributed as dist

    from satellite_computervision_b200 import _lib
    H = W = args.scene
    S = args.scenes
    nbuf = 2  # distinct synthetic scenes resident in pinned host memory per rank (every logical scene is still copied H2D)
    if args.stream_mode == 'scenes':
        mine = list(range(rank, S, world))
        sh = shard_of(0, 1)
    else:
        mine = list(range(S))
        sh = shard_of(rank, world)
    opts = _lib.MosaicOpts()
    opts.tile_begin, opts.tile_end = sh.tile_begin, sh.tile_end
    src_row0, src_row1 = sh.src_row0, sh.src_row1
    ins = []
    for b in range(nbuf):
        full = make_scene(H, W, seed=1 + rank * nbuf + b if args.stream_mode == 'scenes' else 1 + b)
        buf = _lib.pinned_empty((src_row1 - src_row0, W, BANDS), np.uint16)
        buf[...] = full[src_row0:src_row1]
        del full
        ins.append(buf)
    outs = [(_lib.pinned_zeros((H, W), np.float32), _lib.pinned_zeros((H, W), np.uint8)) for _ in range(2)]

    def submit(i):
        src = ins[i % nbuf]
        op, om = outs[i % 2]
        tk = C.c_int(-1)
        _lib.check(lib.scv_stream_submit(eng, C.c_void_p(src.ctypes.data - src_row0 * W * BANDS * 2), _lib.SCV_U16, H, W, BANDS,
                                         C.byref(tiling), C.byref(cn), C.byref(opts), _lib.ptr(op), _lib.ptr(om), C.byref(tk)))
        return tk.value

    for i in range(2):  # warm-up: plans, buffers, clocks
        submit(i)
    _lib.check(lib.scv_stream_wait(eng, -1))
    barrier()
    clk = ClockSampler(local_rank)
    clk.__enter__()
    t0 = time.perf_counter()
    for i in range(len(mine)):
        submit(i)
    _lib.check(lib.scv_stream_wait(eng, -1))
    torch.cuda.synchronize()
    dt = reduce_max(time.perf_counter() - t0)
    clk.__exit__(None, None, None)
    _lib.check(lib.scv_check(eng))
    # last scene's output must look like a prediction (cores written, margins untouched)
    op, om = outs[(len(mine) - 1) % 2]
    ok = bool(op[sh.dst_row0:sh.dst_row1].max() > 0 and op[:64].max() == 0) if len(mine) else True
    if rank == 0:
        mp = H * W / 1e6
        h2d = (src_row1 - src_row0) * W * BANDS * 2
        line = {'metric': 'sustained megapixels/sec, multi-scene streaming (BASELINE configs[4])', 'value': S * mp / dt, 'unit': UNIT,
                'n_gpus': world, 'scenes': S, 'seconds': dt, 'ms_per_scene': dt / S * 1e3, 'higher_is_better': True,
                'mode': args.stream_mode, 'dtype': 'bf16', 'data': 'synthetic',
                'config': {'workload': f'{S} synthetic {H}x{W}x{BANDS} uint16 scenes, variant-A U-Net, {KERNEL}+{BUFF} tiling; '
                                       f'{"whole scenes round-robin over ranks" if args.stream_mode == "scenes" else "every scene sharded over all ranks"}; '
                                       f'{nbuf} distinct scenes per rank in pinned host memory, every logical scene copied H2D and its '
                                       'probability + mask rasters copied D2H inside the timed region; two scenes in flight per GPU'},
                'h2d_bytes_per_scene_per_rank': h2d, 'clocks': clk.summary(), 'output_sane': ok}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not ok:
        raise SystemExit('streamed output does not look like a prediction')


def layer_names(model):
    names = []
    L = len(model.filters)
    if getattr(model, 'is_siamese', False):
        for i in range(L):
            names += [f'encoder_{i}/conv0[a]', f'encoder_{i}/conv0[b]']
        names += ['ASPP/cba', 'ASPP/cba3_3', 'ASPP/cba3_6', 'ASPP/cba3_12', 'ASPP/cba3[a]', 'ASPP/cba3[b]']
        for i in range(L - 1, -1, -1):
            names += [f'decoder_{i}/up', f'decoder_{i}/conv0', f'decoder_{i}/conv1']
        return names
    nconv = 2 if model.double_conv else 1
    for i in range(L):
        names += [f'encoder_{i}/conv{j}' for j in range(nconv)]
    names += [f'center/conv{j}' for j in range(nconv)]
    for i in range(L - 1, -1, -1):
        names += [f'decoder_{i}/up', f'decoder_{i}/conv0', f'decoder_{i}/conv1']
    return names


if __name__ == '__main__':
    main()
