#!/bin/bash
# Build reg.cu variants into image_stitcher_b200/_lib/var_<name>.so for A/B sweeps on the GPU box.
cd "$(dirname "$0")/.."
L=image_stitcher_b200/_lib
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2 --expt-relaxed-constexpr"
build() { # name, defines...
  name=$1; shift
  nvcc $FLAGS "$@" -Xptxas -v -c image_stitcher_b200/csrc/reg.cu -o /tmp/reg_$name.o 2>/tmp/err_$name.txt || { echo FAILED $name; grep error /tmp/err_$name.txt | head -3; return; }
  nvcc -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o $L/var_$name.so $L/api.o $L/fuse.o /tmp/reg_$name.o $L/u8.o
  echo built $name: $(grep -A2 "rows_fwd_kernelIfLi4" /tmp/err_$name.txt | grep -E "spill|Used" | tr '\n' ' ' | cut -c1-200)
}
rm -f $L/var_*.so
build c4l2 -DSB_REG_CTAS=4 -DSB_GEMM_LT=2 &
build c2l4 -DSB_REG_CTAS=2 -DSB_GEMM_LT=4 &
wait
