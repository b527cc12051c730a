#!/bin/bash
# early accumulator release in epilogue_slab (slab, slabw, slab2 kernels)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -3
B="python bench.py --steps 4 --warmup 2 --no-cpu-baseline --profile-layers"
for v in 1 3 1; do
SCV_SLAB2=$v timeout 200 $B > gpurun_out/r02_y_early_$v.json 2> gpurun_out/r02_y_early_$v.err
python - gpurun_out/r02_y_early_$v.json "$v" <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    L = {l['name']: round(l['ms'], 2) for l in d['layers']}
    print('SCV_SLAB2', sys.argv[2], 'value', round(d['value'], 1), round(d['ms_per_step'], 2), d['clocks']['sm_mhz'], 'verify', d.get('verify', {}).get('ok'))
    print('   ', {k: v for k, v in L.items() if k.startswith(('encoder_1', 'decoder_1', 'encoder_2', 'decoder_2')) or k.endswith('/up')})
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex, open(sys.argv[1].replace('.json','.err')).read()[-1200:])
P
done
