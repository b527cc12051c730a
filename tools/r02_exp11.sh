#!/bin/bash
# per-role cycle accounting of the row kernel (CTA 0) on a 2048^2 raster (49 chips): full pipeline and skeleton only
export SCV_LIB_PATH=tools/microbench/build/libscv_prof.so
for d in 32 60 33; do
  echo "== SCV_ROWS_DBG=$d"
  SCV_ROWS_DBG=$d python bench.py --scene 2048 --steps 1 --warmup 0 --no-cpu-baseline --no-verify 2>&1 | grep "rows prof" | head -36
done
