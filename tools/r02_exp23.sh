#!/bin/bash
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
SCV_HOST_TRACE=1 $TR bench.py --gpus $N --no-verify --steps 4 > gpurun_out/r02_tr.json 2> gpurun_out/r02_tr.err
grep "host trace" gpurun_out/r02_tr.err | sed 's/\]\[scv/]\n[scv/g' | grep "dev [03]\]" | tail -16
grep "^\[rank" gpurun_out/r02_tr.err | sed 's/\(host [-0-9.]*\)\[rank/\1\n[rank/g' | sort | head -8
