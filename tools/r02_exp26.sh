#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 4 --warmup 2 --no-cpu-baseline --profile-layers --no-verify"
for lib in "" tools/microbench/build/libscv_deep.so tools/microbench/build/libscv_ni3.so; do
SCV_LIB_PATH=$lib $B > gpurun_out/r02_ad.json 2> gpurun_out/r02_ad.err
python - gpurun_out/r02_ad.json "$lib" <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    L = {l['name']: round(l['ms'], 2) for l in d['layers']}
    print('lib', sys.argv[2] or 'default', 'value', round(d['value'], 1), round(d['ms_per_step'], 2), d['clocks']['sm_mhz'], 'e0', L['encoder_0/conv0'], 'd0', L['decoder_0/conv0'])
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex, open(sys.argv[1].replace('.json','.err')).read()[-600:])
P
done
for lib in tools/microbench/build/libscv_deep.so tools/microbench/build/libscv_ni3.so; do SCV_LIB_PATH=$lib python -m pytest tests/test_gpu_fused.py -m gpu -q 2>&1 | tail -1; done
