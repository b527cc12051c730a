#!/bin/bash
# what the driver runs at round end, in its order
mkdir -p gpurun_out
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
( time python bench.py --impl reference --gpus 1 --steps 3 --warmup 3 ) > gpurun_out/r02_drv_ref.json 2> gpurun_out/r02_drv_ref.err; tail -3 gpurun_out/r02_drv_ref.err | tr '\n' ' '; echo
( time python bench.py --gpus 1 --steps 3 --warmup 3 ) > gpurun_out/r02_drv_ours.json 2> gpurun_out/r02_drv_ours.err; grep real gpurun_out/r02_drv_ours.err
python - <<'P'
import json
r = json.loads(open('gpurun_out/r02_drv_ref.json').read().strip().splitlines()[-1])
d = json.loads(open('gpurun_out/r02_drv_ours.json').read().strip().splitlines()[-1])
print('reference', round(r['value'], 3), r['cpu_baseline']['cores'], '| ours value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'frac', round(d['roofline']['frac'], 3), 'K1', round(d['roofline_extract']['frac'], 3), 'K4', round(d['roofline_stitch']['frac'], 3), 'verify', d['verify']['ok'], d['verify']['mask_agree'], 'launches', d['gpu_launches'], d['clocks'])
print('ratio e2e', round(d['e2e']['value'] / r['value']), 'ratio value', round(d['value'] / r['value']))
P
