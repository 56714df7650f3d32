#!/bin/bash
# A/B of the paste kernel's block-row order: time (CUDA events) and DRAM bytes (ncu) with and without SB_FUSE_ROWPERM
python -m pytest tests -m gpu -x -q > gpurun_out/t2.log 2>&1; tail -5 gpurun_out/t2.log
for rp in 1 0; do
  echo "ROWPERM=$rp"; SB_FUSE_ROWPERM=$rp python scratch/perf_fuse2.py paste 2>&1 | tail -2
done
for rp in 1 0; do
  SB_FUSE_ROWPERM=$rp timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:fuse_paste -s 28 -c 1 --csv python scratch/perf_fuse2.py paste 2>&1 | grep -E "fuse_paste" | awk -F'","' '{print "rp='$rp'", $(NF-2), $(NF-1), $NF}'
done
