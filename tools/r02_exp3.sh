#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_e_pytest.txt
tail -8 gpurun_out/r02_e_pytest.txt
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline"
$B --profile-layers --python-api > gpurun_out/r02_e_layers.json 2> gpurun_out/r02_e_layers.err
SCV_LINEAR_STORE=0 $B --profile-layers --no-verify > gpurun_out/r02_e_layers_nolinear.json 2>&1
for f in gpurun_out/r02_e_layers.json gpurun_out/r02_e_layers_nolinear.json; do python - "$f" <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'K1', round(d['roofline_extract']['frac'], 3), d['roofline_extract']['ms'],
          'K4', round(d['roofline_stitch']['frac'], 3), d['roofline_stitch']['ms'], 'tc', round(d['roofline']['achieved'], 1), d['clocks']['sm_mhz'], d.get('verify'), d.get('e2e_reference_signature', {}).get('ms_per_step'))
    if 'layers' in d:
        print('   ', [(l['name'].replace('encoder_', 'e').replace('decoder_', 'd').replace('conv', 'c'), round(l['ms'], 2)) for l in d['layers'][:2] + d['layers'][-3:]])
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex, open(sys.argv[1]).read()[-800:])
P
done
