#!/bin/bash
# which part of the row kernel bounds the 384x384 layers: switch parts of the pipeline off (results are wrong!)
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-verify --profile-layers"
for dbg in 0 4 8 16 1 20 12 28; do
  SCV_ROWS_DBG=$dbg $B > gpurun_out/r02_f_dbg$dbg.json 2> gpurun_out/r02_f_dbg$dbg.err
  python - gpurun_out/r02_f_dbg$dbg.json $dbg <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    L = {l['name']: l['ms'] for l in d['layers']}
    print('dbg', sys.argv[2], 'e0c0 %.2f e0c1 %.2f d0c0 %.2f d0c1 %.2f  (d0up %.2f)' % (L['encoder_0/conv0'], L['encoder_0/conv1'], L['decoder_0/conv0'], L['decoder_0/conv1'], L['decoder_0/up']), d['clocks']['sm_mhz'])
except Exception as ex:
    print('dbg', sys.argv[2], 'FAILED', ex, open(sys.argv[1]).read()[-300:])
P
done
