#!/bin/bash
# full GPU suite, then compute-sanitizer over the kernels that changed late in round 2: the CTA-pair slab kernel and the
# fused kernel with the async-proxy halo exchange
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02_w_pytest.txt; cat gpurun_out/r02_w_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
SEL='tests/test_gpu_fused.py tests/test_gpu_kernels.py -k "1-8 or cta_pair"'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_fused.py tests/test_gpu_kernels.py -m gpu -q -x -k "1-8 or cta_pair" > gpurun_out/r02_w_memcheck.txt 2>&1; echo "memcheck exit $?" >> gpurun_out/r02_w_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 7 python -m pytest tests/test_gpu_fused.py tests/test_gpu_kernels.py -m gpu -q -x -k "1-8 or cta_pair" > gpurun_out/r02_w_racecheck.txt 2>&1; echo "racecheck exit $?" >> gpurun_out/r02_w_racecheck.txt
for f in memcheck racecheck; do echo "== $f"; grep "ERROR SUMMARY\|RACECHECK SUMMARY\|passed\|failed\|exit" gpurun_out/r02_w_$f.txt | tail -3; grep -m3 "Race reported\|Invalid\|error" gpurun_out/r02_w_$f.txt; done
