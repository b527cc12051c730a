#!/bin/bash
export SCV_LIB_PATH=tools/microbench/build/libscv_prof.so
SCV_ROWS_DBG=32 python bench.py --scene 2048 --steps 1 --warmup 0 --no-cpu-baseline --no-verify 2>&1 | grep "fused prof" | head -24
