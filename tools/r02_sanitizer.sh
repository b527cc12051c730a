#!/bin/bash
# compute-sanitizer over the kernels that hand mbarrier tokens between issuer warps
mkdir -p gpurun_out
SEL='rows or slabw or persistent or small_models or float64 or streamed'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests -m gpu -q -x -k "$SEL" > gpurun_out/r02_l_memcheck.txt 2>&1
echo "memcheck exit $?" >> gpurun_out/r02_l_memcheck.txt
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "rows or slabw or persistent" > gpurun_out/r02_l_racecheck.txt 2>&1
echo "racecheck exit $?" >> gpurun_out/r02_l_racecheck.txt
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "rows or slabw or persistent" > gpurun_out/r02_l_synccheck.txt 2>&1
echo "synccheck exit $?" >> gpurun_out/r02_l_synccheck.txt
for f in memcheck racecheck synccheck; do echo "== $f"; grep -c "=========" gpurun_out/r02_l_$f.txt; grep "ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed\|exit" gpurun_out/r02_l_$f.txt | tail -4; done
