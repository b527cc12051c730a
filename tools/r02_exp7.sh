#!/bin/bash
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
for v in 1 0; do
SCV_HOST_FIRST_ROW=$v $TR bench.py --gpus $N --no-verify > gpurun_out/r02_j_first$v.json 2> gpurun_out/r02_j_first$v.err
python - gpurun_out/r02_j_first$v.json <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], 'value %.1f (%.2f ms) e2e %.1f (%.2f ms)' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step']), d['clocks']['sm_mhz'], d['e2e']['pipeline_ms_last_step_rank0'], d['stage_ms_last_step_rank0'])
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex, open(sys.argv[1].replace('.json', '.err')).read()[-1200:])
P
done
