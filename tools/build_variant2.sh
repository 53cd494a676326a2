#!/bin/bash
# build_variant2.sh NAME "-DFLAG=..." FILE.cu [FILE.cu ...] : like build_variant.sh for several sources
set -e
cd "$(dirname "$0")/../gaussian-splatting-toolkit_b200/csrc"
NAME=$1; FLAGS=$2; shift 2
mkdir -p ../../gpurun_variants build_variants
OBJS=""; EXCL=""
for SRC in "$@"; do
  B=$(basename $SRC .cu); OBJ=build_variants/${NAME}_$B.o
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --use_fast_math -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v --expt-relaxed-constexpr $FLAGS -c $SRC -o $OBJ 2> build_variants/${NAME}_$B.ptxas.log
  OBJS="$OBJS $OBJ"; EXCL="$EXCL -e build/$B.o"
done
OTHERS=$(ls build/*.o | grep -v $EXCL)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../gpurun_variants/libgsr_${NAME}.so $OBJS $OTHERS -lcudart
