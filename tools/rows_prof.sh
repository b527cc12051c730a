#!/bin/bash
# per-role cycle accounting of the row kernel (one batch of 49 chips), for each SCV_ROWS_DBG setting in $SWEEP
for d in ${SWEEP:-32}; do
  echo "== SCV_ROWS_DBG=$d"
  SCV_ROWS_DBG=$d python bench.py --scene 2048 --steps 1 --warmup 0 --no-cpu-baseline 2>&1 | grep "rows prof" | head -${LINES_MAX:-18}
done
