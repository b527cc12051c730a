#!/bin/bash
# CTA-pair slab kernel: correctness first, then the level-1 layers
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "cta_pair" 2>&1 | tail -15
B="python bench.py --steps 4 --warmup 2 --no-cpu-baseline --profile-layers"
for v in 1 0 1; do
SCV_SLAB2=$v timeout 200 $B > gpurun_out/r02_y_slab2_$v.json 2> gpurun_out/r02_y_slab2_$v.err
python - gpurun_out/r02_y_slab2_$v.json "$v" <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    L = {l['name']: round(l['ms'], 2) for l in d['layers']}
    print('SCV_SLAB2', sys.argv[2], 'value', round(d['value'], 1), round(d['ms_per_step'], 2), d['clocks']['sm_mhz'], 'verify', d.get('verify', {}).get('ok'), {k: v for k, v in L.items() if k.startswith(('encoder_1', 'decoder_1'))})
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex, open(sys.argv[1].replace('.json','.err')).read()[-1200:])
P
done
SCV_PLAN_DEBUG=1 timeout 100 python bench.py --scene 2048 --steps 1 --warmup 0 --no-cpu-baseline --no-verify 2>&1 | grep "slab2" | head -6
