#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline"
for v in 1 0 1 0; do
SCV_PDL=$v $B > gpurun_out/r02_q_pdl$v.json 2> gpurun_out/r02_q_pdl$v.err
python - gpurun_out/r02_q_pdl$v.json <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'value', round(d['value'], 1), round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), 'tc', round(d['roofline']['achieved'], 1), d['clocks']['sm_mhz'], d.get('verify', {}).get('ok'), d.get('verify', {}).get('max_abs'))
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex, open(sys.argv[1].replace('.json','.err')).read()[-800:])
P
done
