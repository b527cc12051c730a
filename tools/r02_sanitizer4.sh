#!/bin/bash
# full GPU suite of the final build + memcheck over the siamese network (dilated / 1x1 tile-kernel paths, half-buffer launches)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r02_t_pytest.txt; cat gpurun_out/r02_t_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_siamese.py -m gpu -q -x -k "96 or 64 or 128 or mosaic" > gpurun_out/r02_t_memcheck.txt 2>&1; echo "memcheck exit $?" >> gpurun_out/r02_t_memcheck.txt
grep "ERROR SUMMARY\|passed\|failed\|exit" gpurun_out/r02_t_memcheck.txt | tail -3; grep -m3 "Invalid\|error" gpurun_out/r02_t_memcheck.txt
