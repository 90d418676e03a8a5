#!/bin/bash
# tools/build_variant.sh NAME "-DFOO=1 ..." : tuning build of librevealer_b200 into variants/NAME.so (use with RVL_LIB=...)
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
N=$1; shift
C=revealer_b200/csrc
F="-ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
nvcc $F -fmad=false -c $C/bandwidth.cu -o variants/$N.bandwidth.o
nvcc $F $@ -Xptxas -v -c $C/score.cu -o variants/$N.score.o 2> variants/$N.ptxas.txt
nvcc $F -c $C/capi.cu -o variants/$N.capi.o
nvcc $F -c $C/gct.cu -o variants/$N.gct.o
nvcc $F -Xcompiler -O3 -c $C/hostpack.cpp -o variants/$N.hostpack.o
nvcc -ccbin /usr/bin/g++ -gencode arch=compute_100a,code=sm_100a -shared -o variants/$N.so variants/$N.bandwidth.o variants/$N.score.o variants/$N.capi.o variants/$N.gct.o variants/$N.hostpack.o -lcudart_static -lpthread -ldl -lrt
grep -A2 "k_score" variants/$N.ptxas.txt | grep -E "Used|spill"
