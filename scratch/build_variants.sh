#!/bin/bash
# Build fuse.cu variants into image_stitcher_b200/_lib/var_<name>.so for A/B sweeps on the GPU box.
cd "$(dirname "$0")/.."
L=image_stitcher_b200/_lib
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2 --expt-relaxed-constexpr"
build() { # name, defines...
  name=$1; shift
  nvcc $FLAGS "$@" -c image_stitcher_b200/csrc/fuse.cu -o /tmp/fuse_$name.o 2>/tmp/err_$name.txt || { echo FAILED $name; grep error /tmp/err_$name.txt | head -3; return; }
  nvcc -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o $L/var_$name.so $L/api.o /tmp/fuse_$name.o $L/reg.o $L/u8.o
  echo built $name
}
rm -f $L/var_*.so
build r2p1 -DSB_RECT_ROWS=2 -DSB_RECT_PREFETCH=1 &
build r2p4 -DSB_RECT_ROWS=2 -DSB_RECT_PREFETCH=4 &
build r3p2 -DSB_RECT_ROWS=3 -DSB_RECT_PREFETCH=2 &
build r4p2 -DSB_RECT_ROWS=4 -DSB_RECT_PREFETCH=2 &
build r2p2w4 -DSB_RECT_ROWS=2 -DSB_RECT_PREFETCH=2 -DSB_RECT_WARPS=4 &
build r2p2w16 -DSB_RECT_ROWS=2 -DSB_RECT_PREFETCH=2 -DSB_RECT_WARPS=16 &
wait
