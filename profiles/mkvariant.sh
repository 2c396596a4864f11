#!/bin/bash
# usage: profiles/mkvariant.sh NAME file.cu "-DX=1 ..."   -> build/variants/libNAME.so (other objects reused from csrc/)
set -e
ROOT=$(git rev-parse --show-toplevel)
mkdir -p "$ROOT/build/variants"
cd "$ROOT/roskfpos_b200/csrc"
NAME=$1; SRC=$2; DEFS=$3
FMAD=--fmad=true; [ "$SRC" = kfpos_exact.cu ] && FMAD=--fmad=false
/usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC $FMAD $DEFS -c $SRC -o /tmp/var_$NAME.o
OBJS=$(ls *.o | grep -v "^${SRC%.cu}.o$")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/variants/lib$NAME.so $OBJS /tmp/var_$NAME.o -cudart static
echo built $NAME
