#!/bin/bash
# Build a kernel-variant copy of libscv.so for A/B timing:  tools/build_variant.sh NAME "-DSCV_ROWS_ISSUERS=2 ..."
# -> tools/microbench/build/libscv_NAME.so (select with SCV_LIB_PATH).  Experiments only.
set -e
NAME=$1; DEFS=$2
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/satellite_computervision_b200/csrc
OUT=$ROOT/tools/microbench/build; mkdir -p $OUT/$NAME
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden $DEFS"
for f in conv_umma conv_rows conv_slabw conv_slab2 conv_fused tile_kernels engine; do nvcc $FLAGS -c $SRC/$f.cu -o $OUT/$NAME/$f.o & done; wait
nvcc -shared -o $OUT/libscv_$NAME.so $OUT/$NAME/*.o -cudart static
echo $OUT/libscv_$NAME.so
