#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused.py -m gpu -q -x 2>&1 | tail -5
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --profile-layers"
SCV_FUSE=3 $B > gpurun_out/r02_v_fuse3.json 2> gpurun_out/r02_v_fuse3.err
python - gpurun_out/r02_v_fuse3.json <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'value', round(d['value'], 1), round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), 'tc', round(d['roofline']['achieved'], 1), d['clocks']['sm_mhz'], d.get('verify', {}).get('ok'), d.get('verify', {}).get('max_abs'))
    L = {l['name']: round(l['ms'], 2) for l in d['layers']}
    print('   ', {k.replace('encoder_', 'e').replace('decoder_', 'd').replace('conv', 'c'): v for k, v in L.items() if k.startswith('decoder_0') or k.startswith('encoder_0')})
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex, open(sys.argv[1].replace('.json','.err')).read()[-800:])
P
