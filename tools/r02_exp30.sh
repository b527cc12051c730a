#!/bin/bash
# fused encoder pair: epilogue-B cycle accounting, double-buffered staging, ring-depth / group-count combinations
mkdir -p gpurun_out
for v in ; do
echo "== $v"
SCV_LIB_PATH=tools/microbench/build/libscv_$v.so SCV_ROWS_DBG=32 timeout 120 python bench.py --scene 2048 --steps 1 --warmup 0 --no-cpu-baseline --no-verify 2>&1 | grep "fused prof" | head -12
done
B="python bench.py --steps 4 --warmup 2 --no-cpu-baseline --profile-layers --no-verify"
for v in "" g26n4 g36n5r5 n4r6 g26r5 deep g25n5r5 ""; do
lib=""; [ -n "$v" ] && lib=tools/microbench/build/libscv_$v.so
SCV_LIB_PATH=$lib timeout 120 $B > gpurun_out/r02_y_sweep3_$v.json 2> gpurun_out/r02_y_sweep3_$v.err
python - gpurun_out/r02_y_sweep3_$v.json "$v" <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    L = {l['name']: round(l['ms'], 2) for l in d['layers']}
    print('lib', sys.argv[2] or 'default', 'value', round(d['value'], 1), round(d['ms_per_step'], 2), d['clocks']['sm_mhz'], 'e0', L['encoder_0/conv0'], L['encoder_0/conv1'], 'd0', L['decoder_0/conv0'])
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex, open(sys.argv[1].replace('.json','.err')).read()[-600:])
P
done
for v in g26r5 g36n5r5; do SCV_LIB_PATH=tools/microbench/build/libscv_$v.so timeout 200 python -m pytest tests/test_gpu_fused.py -m gpu -q 2>&1 | tail -1; done
