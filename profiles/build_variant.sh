#!/bin/bash
# Kernel A/B experiments: rebuild ONE translation unit with extra nvcc flags and link it with the other objects of the in-tree
# build into scratch/variants/lib_<NAME>.so (load it with OUZELUM_B200_LIB=<path>).
# usage: profiles/build_variant.sh NAME TU.cu "<extra nvcc flags>"
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
cd $ROOT/ouzelum_b200
NAME=$1; TU=$2; EXTRA=$3
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -fmad=false"
mkdir -p ../scratch/variants/obj_$NAME
nvcc $F $EXTRA -c -o ../scratch/variants/obj_$NAME/${TU%.cu}.o csrc/$TU
OBJS=$(ls build/*.o | grep -v "/${TU%.cu}.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../scratch/variants/lib_$NAME.so ../scratch/variants/obj_$NAME/${TU%.cu}.o $OBJS
echo built scratch/variants/lib_$NAME.so
