#!/bin/bash
# round-2 evidence run on one B200: tests, bench (both arms), ncu launch list, per-layer ncu counters
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02_v_pytest.txt; cat gpurun_out/r02_v_pytest.txt
python bench.py --impl reference > gpurun_out/r02_v_bench_reference_cpu.json 2> gpurun_out/r02_v_ref.err
python bench.py --profile-layers --python-api > gpurun_out/r02_v_bench_1gpu.json 2> gpurun_out/r02_v_bench.err
tail -c 300 gpurun_out/r02_v_bench_1gpu.json; tail -2 gpurun_out/r02_v_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 420 --csv --log-file gpurun_out/r02_v_launches.csv python bench.py --steps 1 --warmup 0 --no-verify --no-cpu-baseline > gpurun_out/r02_v_ncu_launches.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_v_layers_metrics.csv python tools/ncu_one_batch.py 63 > gpurun_out/r02_v_ncu_layers.log 2>&1
tail -2 gpurun_out/r02_v_ncu_layers.log; wc -l gpurun_out/r02_v_launches.csv gpurun_out/r02_v_layers_metrics.csv
