#!/bin/bash
# Ablation builds of the fused front kernel (WRONG results by construction -- timing only, never shipped): ab/lib_front_<name>.so with one stage of the
# pipeline removed, selected with EIGB200_LIB.  usage: tools/ablate_front.sh   (from the repository root, after `python -m eigb200.build`)
set -e
PKG=task-level-insights-from-eigenvalues-across-sequence-models_b200
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -diag-suppress 177 -I include"
mkdir -p ab /tmp/abl
OBJS=$(ls $PKG/build/*.o | grep -v -e k7_front_fused.o)
VARS=("trace:-DFF_TRACE" "scan:-DFF_ABL_SCAN" "bpre:-DFF_BPREFETCH")
for v in "${VARS[@]}"; do
  name=${v%%:*}; defs=${v#*:}
  nvcc $FLAGS $defs -c $PKG/csrc/k7_front_fused.cu -o /tmp/abl/k7_$name.o &
done
wait
for v in "${VARS[@]}"; do
  name=${v%%:*}
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ab/lib_front_$name.so $OBJS /tmp/abl/k7_$name.o
done
ls -la ab/
