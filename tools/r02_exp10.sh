#!/bin/bash
mkdir -p gpurun_out
T="tests/test_gpu_kernels.py -m gpu -q -x -k persistent"
for ni in 1 2; do
SCV_PTILE_ISSUERS=$ni timeout 600 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest $T > gpurun_out/r02_n_sync_ni$ni.txt 2>&1
echo "ni=$ni exit $?"; grep "ERROR SUMMARY\|passed\|failed" gpurun_out/r02_n_sync_ni$ni.txt | tail -2
grep -A4 "Barrier error" gpurun_out/r02_n_sync_ni$ni.txt | head -12
done
