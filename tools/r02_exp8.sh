#!/bin/bash
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
nvidia-smi topo -m > gpurun_out/r02_k_topo.txt 2>&1
$TR bench.py --gpus $N --no-verify --steps 5 > gpurun_out/r02_k_a.json 2> gpurun_out/r02_k_a.err
grep "^\[rank" gpurun_out/r02_k_a.err | sort
SCV_HOST_FIRST_ROW=0 $TR bench.py --gpus $N --no-verify --steps 5 > gpurun_out/r02_k_b.json 2> gpurun_out/r02_k_b.err
grep "^\[rank" gpurun_out/r02_k_b.err | sort
head -14 gpurun_out/r02_k_topo.txt | cut -c1-150
