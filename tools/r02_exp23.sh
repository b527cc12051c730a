#!/bin/bash
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$TR bench.py --gpus $N --gather > gpurun_out/r02_tr.json 2> gpurun_out/r02_tr.err

grep "^\[rank" gpurun_out/r02_tr.err | sed 's/\(host [-0-9.]*\)\[rank/\1\n[rank/g' | sort | head -8
python - <<'P'
import json
d = json.loads(open('gpurun_out/r02_tr.json').read().strip().splitlines()[-1])
print('value', round(d['value'], 1), d['ms_per_step'], 'e2e', d['e2e'], 'gather', d.get('optional_gather_ms'), 'verify', d.get('verify', {}).get('ok'), d['clocks'])
P
