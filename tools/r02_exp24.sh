#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --profile-layers"
for v in 1 0; do
SCV_ROWS_PARTIAL=$v $B > gpurun_out/r02_ab_part$v.json 2> gpurun_out/r02_ab_part$v.err
python - gpurun_out/r02_ab_part$v.json $v <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    L = {l['name']: round(l['ms'], 2) for l in d['layers']}
    print('partial', sys.argv[2], 'value', round(d['value'], 1), round(d['ms_per_step'], 2), d['clocks']['sm_mhz'], d.get('verify', {}).get('ok'), {k.replace('encoder_', 'e').replace('decoder_', 'd').replace('conv', 'c'): v for k, v in L.items() if k.startswith('decoder_1') or k.startswith('encoder_1')})
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex, open(sys.argv[1].replace('.json','.err')).read()[-600:])
P
done
