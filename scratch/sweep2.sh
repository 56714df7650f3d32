#!/bin/bash
for v in libstitchb200 $(cd image_stitcher_b200/_lib; ls var_*.so | sed 's/.so$//'); do
  export SB_LIB_PATH=/root/repo/image_stitcher_b200/_lib/$v.so
  echo "== $v"; timeout 120 python scratch/perf_fuse2.py paste 2>&1 | tail -2
done
