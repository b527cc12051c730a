#!/bin/bash
# halo exchange of the fused kernel through async-proxy bulk copies (no MEMBAR.ALL.GPU on the conv-1 -> conv-2 path)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused.py -m gpu -q 2>&1 | tail -3
B="python bench.py --steps 4 --warmup 2 --no-cpu-baseline --no-verify"
timeout 200 $B --profile-layers > gpurun_out/r02_y_halo_layers.json 2> gpurun_out/r02_y_halo_layers.err
timeout 200 $B > gpurun_out/r02_y_halo.json 2> gpurun_out/r02_y_halo.err
python - <<'P'
import json
for f in ('gpurun_out/r02_y_halo_layers.json', 'gpurun_out/r02_y_halo.json'):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value', round(d['value'], 1), round(d['ms_per_step'], 2), 'e2e', round(d['e2e']['value'], 1), d['clocks']['sm_mhz'], 'net', round(d['stage_ms_last_step_rank0']['network_ms'], 2))
        for l in d.get('layers', []):
            if l['name'].startswith(('encoder_0', 'decoder_0')): print('  ', l['name'], round(l['ms'], 2))
    except Exception as ex:
        print(f, 'FAILED', ex, open(f.replace('.json', '.err')).read()[-800:])
P
