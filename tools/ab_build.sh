#!/bin/bash
# A/B build of ONE translation unit with extra -D flags: ab/lib_<name>.so = the current objects with csrc/<stem>.cu recompiled.  Select with EIGB200_LIB.
# usage: tools/ab_build.sh <stem> <name> "<-D flags>"     e.g. tools/ab_build.sh k4_gemm_fused nopf "-DFG_NO_L2_PREFETCH"
set -e
PKG=task-level-insights-from-eigenvalues-across-sequence-models_b200
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -diag-suppress 177 -I include"
mkdir -p ab /tmp/abl
nvcc $FLAGS $3 -c $PKG/csrc/$1.cu -o /tmp/abl/$1_$2.o
OBJS=$(ls $PKG/build/*.o | grep -v -e "/$1.o")
nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ab/lib_$2.so $OBJS /tmp/abl/$1_$2.o
echo ab/lib_$2.so
