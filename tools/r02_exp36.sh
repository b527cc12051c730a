#!/bin/bash
# device batch size: 126 (14 batches per scene) vs 196 (9) vs 252 (7)
mkdir -p gpurun_out
for mb in 126 252 196; do
timeout 100 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-verify --max-batch $mb > gpurun_out/r02_o_mb$mb.json 2> gpurun_out/r02_o_mb$mb.err
python - gpurun_out/r02_o_mb$mb.json $mb <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('max_batch', sys.argv[2], 'value', round(d['value'], 1), round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), d['clocks']['sm_mhz'])
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex, open(sys.argv[1].replace('.json','.err')).read()[-400:])
P
done
