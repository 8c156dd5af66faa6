#!/bin/bash
# Ablation builds of the tcgen05 GEMM (WRONG results by construction -- timing only, never shipped): ab/lib_abl_<name>.so with parts of the pipeline
# removed, selected with EIGB200_LIB.  usage: tools/ablate_gemm.sh   (from the repository root, after `python -m eigb200.build`)
set -e
PKG=task-level-insights-from-eigenvalues-across-sequence-models_b200
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -diag-suppress 177 -I include"
mkdir -p ab /tmp/abl
OBJS=$(ls $PKG/build/*.o | grep -v -e k4_gemm_tc.o -e stubs.o)
for v in "epi:-DEIGB_ABL_EPI" "conv:-DEIGB_ABL_CONV" "both:-DEIGB_ABL_EPI -DEIGB_ABL_CONV"; do
  name=${v%%:*}; defs=${v#*:}
  nvcc $FLAGS $defs -c $PKG/csrc/k4_gemm_tc.cu -o /tmp/abl/k4_$name.o &
done
wait
for name in epi conv both; do
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ab/lib_abl_$name.so $OBJS /tmp/abl/k4_$name.o
done
ls -la ab/
