#!/bin/bash
# Build fuse.cu variants into image_stitcher_b200/_lib/var_<name>.so for A/B sweeps on the GPU box.
cd "$(dirname "$0")/.."
L=image_stitcher_b200/_lib
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-O2 --expt-relaxed-constexpr"
build() { # name, defines...
  name=$1; shift
  nvcc $FLAGS "$@" -c image_stitcher_b200/csrc/fuse.cu -o /tmp/fuse_$name.o 2>/tmp/err_$name.txt || { echo FAILED $name; grep error /tmp/err_$name.txt | head -3; return; }
  nvcc -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o $L/var_$name.so $L/api.o /tmp/fuse_$name.o $L/reg.o $L/u8.o $L/pyramid.o $L/flatfield.o
  echo built $name
}
rm -f $L/var_*.so
build keep1 -DSB_RECT_FLAT_KEEP=1 &
build keep2 -DSB_RECT_FLAT_KEEP=2 &
build minb5 -DSB_RECT_MINB=5 &
build minb6 -DSB_RECT_MINB=6 &
wait
build keep1minb5 -DSB_RECT_FLAT_KEEP=1 -DSB_RECT_MINB=5 &
build keep2minb5 -DSB_RECT_FLAT_KEEP=2 -DSB_RECT_MINB=5 &
wait
