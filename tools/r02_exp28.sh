#!/bin/bash
# fused-kernel knob sweep after the async-proxy halo exchange + per-role cycle accounting
mkdir -p gpurun_out
SCV_LIB_PATH=tools/microbench/build/libscv_prof.so SCV_ROWS_DBG=32 timeout 120 python bench.py --scene 2048 --steps 1 --warmup 0 --no-cpu-baseline --no-verify 2>&1 | grep "fused prof" | head -30
B="python bench.py --steps 4 --warmup 2 --no-cpu-baseline --profile-layers --no-verify"
for v in "" g36 g26 g46 g13 deep ni3 h22 ""; do
lib=""; [ -n "$v" ] && lib=tools/microbench/build/libscv_$v.so
SCV_LIB_PATH=$lib timeout 120 $B > gpurun_out/r02_y_sweep_$v.json 2> gpurun_out/r02_y_sweep_$v.err
python - gpurun_out/r02_y_sweep_$v.json "$v" <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    L = {l['name']: round(l['ms'], 2) for l in d['layers']}
    print('lib', sys.argv[2] or 'default', 'value', round(d['value'], 1), round(d['ms_per_step'], 2), d['clocks']['sm_mhz'], 'e0', L['encoder_0/conv0'], 'd0', L['decoder_0/conv0'])
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex, open(sys.argv[1].replace('.json','.err')).read()[-600:])
P
done
