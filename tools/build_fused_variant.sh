#!/bin/bash
# Fast kernel-variant build for the fused kernel's compile-time knobs: only conv_fused.cu and engine.cu see
# conv_fused.cuh, the other objects come from the default in-tree build.  tools/build_fused_variant.sh NAME "-D..."
set -e
NAME=$1; DEFS=$2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/satellite_computervision_b200/csrc
OUT=$ROOT/tools/microbench/build; mkdir -p $OUT/$NAME
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden $DEFS"
for f in conv_fused engine; do nvcc $FLAGS -c $SRC/$f.cu -o $OUT/$NAME/$f.o & done; wait
nvcc -shared -o $OUT/libscv_$NAME.so $OUT/$NAME/conv_fused.o $OUT/$NAME/engine.o $SRC/build/conv_umma.o $SRC/build/conv_rows.o \
  $SRC/build/conv_slabw.o $SRC/build/conv_slab2.o $SRC/build/tile_kernels.o -cudart static
echo $OUT/libscv_$NAME.so
