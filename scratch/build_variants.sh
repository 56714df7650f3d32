#!/bin/bash
# Build fuse.cu variants (paste-kernel tunables) into image_stitcher_b200/_lib/var_<name>.so for A/B sweeps on the GPU box.
cd "$(dirname "$0")/.."
L=image_stitcher_b200/_lib
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2 --expt-relaxed-constexpr"
build() { # name, defines...
  name=$1; shift
  nvcc $FLAGS "$@" -c image_stitcher_b200/csrc/fuse.cu -o /tmp/fuse_$name.o 2>/tmp/err_$name.txt || { echo FAILED $name; grep error /tmp/err_$name.txt | head -3; return; }
  nvcc -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o $L/var_$name.so $L/api.o /tmp/fuse_$name.o $L/reg.o
  echo built $name
}
rm -f $L/var_*.so
build f_ph4w16s4 -DSB_P1_PH=4 -DSB_P1_WARPS=16 -DSB_P1_SLOTS=4 &
build f_ph4w16s3 -DSB_P1_PH=4 -DSB_P1_WARPS=16 -DSB_P1_SLOTS=3 &
build f_ph4w24s2 -DSB_P1_PH=4 -DSB_P1_WARPS=24 -DSB_P1_SLOTS=2 &
build n_ph4w16s4 -DSB_P0_PH=4 -DSB_P0_WARPS=16 -DSB_P0_SLOTS=4 &
build n_ph4w8s6 -DSB_P0_PH=4 -DSB_P0_WARPS=8 -DSB_P0_SLOTS=6 &
build n_ph2w16s4 -DSB_P0_PH=2 -DSB_P0_WARPS=16 -DSB_P0_SLOTS=4 &
wait
