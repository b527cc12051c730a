#!/bin/bash
# N-GPU sanity of the final build: bench (with verify and the optional gather), no cpu baseline
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$TR bench.py --gpus $N --gather --no-cpu-baseline > gpurun_out/r02_s_bench_${N}gpu.json 2> gpurun_out/r02_s_bench_${N}gpu.err
grep "^\[rank" gpurun_out/r02_s_bench_${N}gpu.err | sort
python - gpurun_out/r02_s_bench_${N}gpu.json <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    keep = {k: d.get(k) for k in ('value', 'ms_per_step', 'e2e', 'optional_gather_ms', 'verify') if d.get(k) is not None}
    keep['clocks'] = (d.get('clocks') or {}).get('sm_mhz')
    print(sys.argv[1].split('/')[-1], json.dumps(keep)[:1200])
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex, open(sys.argv[1].replace('.json', '.err')).read()[-1200:])
P
