#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_fused.py -m gpu -q -x -k "1-8" > gpurun_out/r02_l2_memcheck.txt 2>&1; echo "memcheck exit $?" >> gpurun_out/r02_l2_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 7 python -m pytest tests/test_gpu_fused.py -m gpu -q -x -k "1-8" > gpurun_out/r02_l2_racecheck.txt 2>&1; echo "racecheck exit $?" >> gpurun_out/r02_l2_racecheck.txt
for f in memcheck racecheck; do echo "== $f"; grep "ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed\|exit" gpurun_out/r02_l2_$f.txt | tail -3; grep -m3 "Race reported\|Invalid\|error" gpurun_out/r02_l2_$f.txt; done
