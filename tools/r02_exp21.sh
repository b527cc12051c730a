#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fused.py -m gpu -q 2>&1 | tail -2
B="python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-verify"
for v in 3 1 3 1 0; do
SCV_FUSE=$v $B > gpurun_out/r02_x_fuse$v.json 2> gpurun_out/r02_x_fuse$v.err
python - gpurun_out/r02_x_fuse$v.json $v <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('SCV_FUSE', sys.argv[2], 'value', round(d['value'], 1), round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), d['clocks']['sm_mhz'], 'net', round(d['stage_ms_last_step_rank0']['network_ms'], 2))
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex, open(sys.argv[1].replace('.json','.err')).read()[-600:])
P
done
