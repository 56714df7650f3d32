#!/bin/bash
for v in libstitchb200 var_A var_B var_C; do
  export SB_LIB_PATH=/root/repo/image_stitcher_b200/_lib/$v.so
  a=$(timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fuse_ -s 4 -c 1 --csv python scratch/perf_fuse2.py paste 2>&1 | grep "fuse_" | rev | cut -d, -f1 | rev)
  b=$(timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:fuse_ -s 28 -c 1 --csv python scratch/perf_fuse2.py paste 2>&1 | grep "fuse_" | rev | cut -d, -f1 | rev)
  echo "$v noflat=$a flat=$b"
done
