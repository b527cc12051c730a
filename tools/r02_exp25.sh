#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_fused.py tests/test_gpu_baseline_configs.py -m gpu -q 2>&1 | tail -3
B="python bench.py --steps 4 --warmup 2 --no-cpu-baseline"
for i in 1 2; do
$B --profile-layers > gpurun_out/r02_ac.json 2> gpurun_out/r02_ac.err
python - gpurun_out/r02_ac.json <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    L = {l['name']: round(l['ms'], 2) for l in d['layers']}
    print('value', round(d['value'], 1), round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), d['clocks']['sm_mhz'], d.get('verify', {}).get('ok'), 'e0', L['encoder_0/conv0'], 'd0', L['decoder_0/conv0'])
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex, open(sys.argv[1].replace('.json','.err')).read()[-600:])
P
done
