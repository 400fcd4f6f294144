#!/bin/bash
# usage: mkvar.sh TAG "-DSBP_BSR_CH=16 ..."
set -e
cd /root/repo/summationbypartsoperatorsextra.jl_b200/csrc
TAG=$1; shift
mkdir -p build/var_$TAG
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr $@"
for f in api bsr bsr3; do nvcc $FLAGS -c $f.cu -o build/var_$TAG/$f.o 2> build/var_$TAG/$f.ptxas.log & done; wait
OBJS=""
for f in dense_small dense_small_sym dense_small_v4 dense_gemm sparse table_generic reduce solve_small lowstorage objective; do OBJS="$OBJS build/$f.o"; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libsbp_b200_$TAG.so build/var_$TAG/api.o build/var_$TAG/bsr.o build/var_$TAG/bsr3.o $OBJS -lcudart -ldl
echo built $TAG
