#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 4 --warmup 2 --no-cpu-baseline --profile-layers"
for i in 1 2 3 4; do
timeout 200 $B > gpurun_out/r02_y_slab2b_$i.json 2> gpurun_out/r02_y_slab2b_$i.err
python - gpurun_out/r02_y_slab2b_$i.json "$i" <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    L = {l['name']: round(l['ms'], 2) for l in d['layers']}
    print('run', sys.argv[2], 'value', round(d['value'], 1), round(d['ms_per_step'], 2), d['clocks']['sm_mhz'], 'verify', d.get('verify', {}).get('ok'), {k: v for k, v in L.items() if k.startswith(('encoder_1', 'decoder_1'))})
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex, open(sys.argv[1].replace('.json','.err')).read()[-1200:])
P
done
