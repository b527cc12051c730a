#!/bin/bash
# A/B on one box: suspend-time hint (default build) vs plain try_wait spin (libscv_nohint.so)
mkdir -p gpurun_out
B="python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-verify"
for v in nohint "" nohint ""; do
lib=""; [ -n "$v" ] && lib=tools/microbench/build/libscv_$v.so
SCV_LIB_PATH=$lib timeout 120 $B > gpurun_out/r02_p_ab_$v.json 2> gpurun_out/r02_p_ab_$v.err
python - gpurun_out/r02_p_ab_$v.json "$v" <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print('lib', sys.argv[2] or 'hint(default)', 'value', round(d['value'], 1), round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), d['clocks']['sm_mhz'])
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex, open(sys.argv[1].replace('.json','.err')).read()[-600:])
P
done
