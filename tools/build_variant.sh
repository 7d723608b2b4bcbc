#!/bin/bash
# Development aid: build a variant of the library with extra -D flags for conv_tc.cu into ab/lib_<name>.so
# usage: tools/build_variant.sh name "-DFOO=1 -DBAR=0"
set -e
cd "$(dirname "$0")/../libsrcnn_b200/csrc"
mkdir -p ../../ab build
NVF="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -ccbin /usr/bin/g++ -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off"
/usr/local/cuda/bin/nvcc $NVF $2 -c conv_tc.cu -o build/conv_tc_$1.o
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o ../../ab/lib_$1.so build/color.o build/resize.o build/prep.o build/conv_exact.o build/conv_tc_$1.o build/srcnn_ctx.o build/contrib.o build/libsrcnn_api.o build/srcnn_multi.o -cudart static -lpthread -ldl -lrt
echo built ab/lib_$1.so
