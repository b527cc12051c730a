#!/bin/bash
# round-2 experiment batch 1: full GPU test suite, per-layer bench, K1/K4 variants
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r02_c_pytest.txt
B="python bench.py --steps 2 --warmup 2 --no-verify --no-cpu-baseline"
$B --profile-layers > gpurun_out/r02_c_layers.json 2> gpurun_out/r02_c_layers.err
SCV_ROWS_FIRST=0 $B --profile-layers > gpurun_out/r02_c_layers_nofirst.json 2>&1
SCV_K1_UNORDERED=1 $B > gpurun_out/r02_c_k1_unordered.json 2>&1
SCV_K4_PX=4 $B > gpurun_out/r02_c_k4_px4.json 2>&1
for f in gpurun_out/r02_c_layers.json gpurun_out/r02_c_layers_nofirst.json gpurun_out/r02_c_k1_unordered.json gpurun_out/r02_c_k4_px4.json; do
  python - "$f" <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'K1', round(d['roofline_extract']['frac'], 3), d['roofline_extract']['ms'],
          'K4', round(d['roofline_stitch']['frac'], 3), d['roofline_stitch']['ms'], 'tc', round(d['roofline']['achieved'], 1), d['clocks']['sm_mhz'])
    if 'layers' in d:
        print('   ', [(l['name'].replace('encoder_', 'e').replace('decoder_', 'd').replace('conv', 'c'), round(l['ms'], 2)) for l in d['layers'][:2] + d['layers'][-3:]])
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex)
P
done
cat gpurun_out/r02_c_pytest.txt | tail -25
