#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -m gpu -q -k "rows or slabw or persistent" > gpurun_out/r02_l_synccheck.txt 2>&1
echo "synccheck exit $?" >> gpurun_out/r02_l_synccheck.txt
grep "ERROR SUMMARY\|passed\|failed\|exit" gpurun_out/r02_l_synccheck.txt | tail -4
grep "Barrier error\|Device Frame: void" gpurun_out/r02_l_synccheck.txt | sort | uniq -c | head
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline --profile-layers"
for g in 0 1; do
SCV_ROWS_GROUP=$g $B > gpurun_out/r02_m_grp$g.json 2> gpurun_out/r02_m_grp$g.err
python - gpurun_out/r02_m_grp$g.json <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'tc', round(d['roofline']['achieved'], 1), d['clocks']['sm_mhz'], d.get('verify', {}).get('ok'), d.get('verify', {}).get('mask_agree'))
    print('   ', [(l['name'].replace('encoder_', 'e').replace('decoder_', 'd').replace('conv', 'c'), round(l['ms'], 2)) for l in d['layers'][:2] + d['layers'][-3:]])
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex, open(sys.argv[1].replace('.json','.err')).read()[-800:])
P
done
