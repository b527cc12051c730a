#!/bin/bash
# time the level-0 layers with each kernel-variant library
for v in "$@"; do
  SCV_LIB_PATH=$PWD/tools/microbench/build/libscv_$v.so python bench.py --steps 2 --warmup 2 --profile-layers --no-cpu-baseline > gpurun_out/variant_$v.json 2> gpurun_out/variant_$v.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/variant_$v.json').read().strip().splitlines()[-1])
L={l['name']:l['ms'] for l in d['layers']}
print('$v', round(d['ms_per_step'],2), {k:round(L[k],2) for k in ('encoder_0/conv1','decoder_0/conv0','decoder_0/conv1')})
PY
done
