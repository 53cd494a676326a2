#!/bin/bash
# build_variant.sh NAME FILE.cu "-DFLAG=..." : libgsr variant with one object recompiled with extra flags -> gpurun_variants/libgsr_NAME.so
set -e
cd "$(dirname "$0")/../gaussian-splatting-toolkit_b200/csrc"
NAME=$1; SRC=$2; FLAGS=$3
mkdir -p ../../gpurun_variants build_variants
OBJ=build_variants/${NAME}_$(basename $SRC .cu).o
PRECISE=""
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --use_fast_math -Xcompiler -fPIC,-fvisibility=hidden -Xptxas -v --expt-relaxed-constexpr $FLAGS -c $SRC -o $OBJ 2> build_variants/${NAME}.ptxas.log
OTHERS=$(ls build/*.o | grep -v "build/$(basename $SRC .cu).o")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../gpurun_variants/libgsr_${NAME}.so $OBJ $OTHERS -lcudart
grep -E "registers|spill" build_variants/${NAME}.ptxas.log | sort | uniq -c | sort -rn | head -4
