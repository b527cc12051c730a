#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r02_d_pytest.txt
tail -6 gpurun_out/r02_d_pytest.txt
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline"
$B --python-api > gpurun_out/r02_d_pyapi_reg1.json 2> gpurun_out/r02_d_pyapi_reg1.err
SCV_HOST_REGISTER=0 $B --python-api > gpurun_out/r02_d_pyapi_reg0.json 2>&1
python bench.py --scenes 8 > gpurun_out/r02_d_stream8.json 2> gpurun_out/r02_d_stream8.err
for f in gpurun_out/r02_d_pyapi_reg1.json gpurun_out/r02_d_pyapi_reg0.json gpurun_out/r02_d_stream8.json; do python - "$f" <<'P'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'value', round(d['value'], 1), d.get('e2e'), d.get('e2e_reference_signature', {}).get('ms_per_step'), d.get('verify'), d.get('ms_per_scene'))
except Exception as ex:
    print(sys.argv[1], 'FAILED', ex, open(sys.argv[1]).read()[-600:])
P
done
timeout 300 ncu --set full --import-source on --clock-control none -k regex:'extract_u16x6|stitch_kernel' -c 2 -o gpurun_out/r02_d_k1k4 -f python bench.py --steps 1 --warmup 0 --no-verify --no-cpu-baseline > gpurun_out/r02_d_ncu1.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:'conv_slab_kernel<.int.64, .int.128, .int.2|conv_rows_kernel<.int.16|conv_slab_kernel<.int.64, .int.64, .int.0, .int.9, .int.2' -o gpurun_out/r02_d_convs -f python tools/ncu_one_batch.py 63 > gpurun_out/r02_d_ncu2.log 2>&1
tail -3 gpurun_out/r02_d_ncu1.log gpurun_out/r02_d_ncu2.log; ls -la gpurun_out/*.ncu-rep
