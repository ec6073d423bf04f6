#!/bin/bash
# Builds an experimental variant of libcxlspeckv.so: the tuned codec translation unit is compiled
# from SRC (default: the tree's kv_codec_fast.cu) with extra nvcc flags, the other objects are
# reused.  usage: profiles/build_variant.sh NAME [SRC] [-DFLAG ...]   -> build/variants/libNAME.so
# Select at run time with SPECKV_LIB=cxl_speckv_b200/build/variants/libNAME.so
set -e
cd "$(dirname "$0")/.."
NAME=$1; shift
SRC=cxl_speckv_b200/csrc/kv_codec_fast.cu
if [ -n "$1" ] && [ "${1:0:1}" != "-" ]; then SRC=$1; shift; fi
B=cxl_speckv_b200/build
mkdir -p $B/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --ftz=false --prec-div=true --prec-sqrt=true \
  --fmad=false -Xcompiler -fPIC,-O2,-fvisibility=hidden -Icxl_speckv_b200/csrc -Iinclude "$@" -x cu -c $SRC \
  -o $B/variants/$NAME.fast.o
OBJS=$(ls $B/*.o | grep -v kv_codec_fast)
nvcc -shared -o $B/variants/lib$NAME.so $OBJS $B/variants/$NAME.fast.o -gencode arch=compute_100a,code=sm_100a \
  -lcudart_static -lpthread -ldl -lrt
echo $B/variants/lib$NAME.so
