#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fused.py -m gpu -q -x 2>&1 | tail -15
