#!/bin/bash
# Development aid: build a variant of the library with extra -D switches for ONE translation unit.
# usage: tools/ab_build.sh NAME UNIT "-DFLAG=1 ..."      ->  ab/lib_NAME.so   (UNIT = conv_tc | prep | color | ...)
set -e
cd "$(dirname "$0")/../libsrcnn_b200/csrc"
NAME=$1; UNIT=$2; FLAGS=$3
make -s >/dev/null
mkdir -p ../../ab/obj
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -ccbin /usr/bin/g++ \
  -Xcompiler -fPIC,-fvisibility=hidden,-ffp-contract=off -Xptxas -v $FLAGS -c $UNIT.cu -o ../../ab/obj/${UNIT}_$NAME.o 2> ../../ab/obj/${UNIT}_$NAME.log
OBJS=$(ls build/*.o | grep -v "build/$UNIT.o" | grep -v -e imgio -e srcnn_cli -e jpegdec)
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o ../../ab/lib_$NAME.so $OBJS ../../ab/obj/${UNIT}_$NAME.o -cudart static -lpthread -ldl -lrt
echo "ab/lib_$NAME.so"
