#!/bin/bash
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$TR bench.py --gpus $N --gather > gpurun_out/r02_y_bench_${N}gpu.json 2> gpurun_out/r02_y_bench_${N}gpu.err
$TR bench.py --gpus $N --scenes $((8 * N)) > gpurun_out/r02_y_stream_${N}gpu_scenes.json 2> gpurun_out/r02_y_stream_${N}gpu_scenes.err
grep "^\[rank" gpurun_out/r02_y_bench_${N}gpu.err | sort
for f in gpurun_out/r02_y_bench_${N}gpu.json gpurun_out/r02_y_stream_${N}gpu_scenes.json; do
python - "$f" <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    keep = {k: d.get(k) for k in ('value', 'ms_per_step', 'e2e', 'optional_gather_ms', 'verify', 'ms_per_scene', 'mode', 'scenes') if d.get(k) is not None}
    keep['clocks'] = (d.get('clocks') or {}).get('sm_mhz')
    print(sys.argv[1].split('/')[-1], json.dumps(keep)[:900])
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex, open(sys.argv[1].replace('.json', '.err')).read()[-1200:])
P
done
