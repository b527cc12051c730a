#!/bin/bash
# timing experiments for the row kernel (SCV_ROWS_DBG flags: 1 = no epilogue math/stores, 2 = no TMEM zeroing, 4 = no A loads, 8 = no MMAs, 16 = no TMA store)
for d in ${SWEEP:-0 1 2 3 16}; do
  SCV_ROWS_DBG=$d python bench.py --steps 1 --warmup 1 --profile-layers --no-cpu-baseline > gpurun_out/rows_dbg_$d.json 2> gpurun_out/rows_dbg_$d.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/rows_dbg_$d.json').read().strip().splitlines()[-1])
L={l['name']:l['ms'] for l in d['layers']}
print('dbg=$d', {k:round(L[k],2) for k in ('encoder_0/conv1','decoder_0/conv0','decoder_0/conv1','encoder_1/conv1')})
PY
done
