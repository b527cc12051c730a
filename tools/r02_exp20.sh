#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 3 --warmup 2 --no-cpu-baseline --profile-layers --no-verify"
run() { # name env...
  name=$1; shift
  env "$@" $B > gpurun_out/r02_w_$name.json 2> gpurun_out/r02_w_$name.err
  python - gpurun_out/r02_w_$name.json <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    L = {l['name']: round(l['ms'], 2) for l in d['layers']}
    print(sys.argv[1].split('/')[-1], 'value', round(d['value'], 1), round(d['ms_per_step'], 2), d['clocks']['sm_mhz'], {k.replace('encoder_', 'e').replace('decoder_', 'd').replace('conv', 'c'): v for k, v in L.items() if k.startswith('decoder_0') or k.startswith('encoder_0')})
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex, open(sys.argv[1].replace('.json','.err')).read()[-600:])
P
}
run base_fuse1 SCV_FUSE=1
run g31_fuse1 SCV_FUSE=1 SCV_LIB_PATH=tools/microbench/build/libscv_g31.so
run g31_fuse3 SCV_FUSE=3 SCV_LIB_PATH=tools/microbench/build/libscv_g31.so
SCV_LIB_PATH=tools/microbench/build/libscv_g31.so python -m pytest tests/test_gpu_fused.py -m gpu -q 2>&1 | tail -2
