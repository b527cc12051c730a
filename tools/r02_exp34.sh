#!/bin/bash
# mbarrier try_wait with a suspend-time hint (fewer spin iterations in the issue-bound kernels)
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 200 python bench.py --steps 4 --warmup 2 --no-cpu-baseline --profile-layers > gpurun_out/r02_p_hint.json 2> gpurun_out/r02_p_hint.err
python - gpurun_out/r02_p_hint.json <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    L = {l['name']: round(l['ms'], 2) for l in d['layers']}
    print('value', round(d['value'], 1), round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), d['clocks']['sm_mhz'], 'verify', d.get('verify', {}).get('ok'), 'frac', round(d['roofline']['frac'], 4))
    print(L)
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex, open(sys.argv[1].replace('.json','.err')).read()[-1200:])
P
