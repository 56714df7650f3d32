#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/t2.log 2>&1; tail -3 gpurun_out/t2.log
export SB_FUSE_ROWPERM=0
for d in 0 16 32 64; do
  echo "DEBUG=$d"; SB_FUSE_DEBUG=$d python scratch/perf_fuse2.py paste 2>&1 | tail -1
  SB_FUSE_DEBUG=$d timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:fuse_paste -s 28 -c 1 --csv python scratch/perf_fuse2.py paste 2>&1 | grep -E "fuse_paste" | awk -F'","' '{printf "%s=%s ", $(NF-2), $NF}'; echo
done
export SB_FUSE_ROWPERM=1
for d in 16 32; do
  echo "ROWPERM DEBUG=$d"; SB_FUSE_DEBUG=$d python scratch/perf_fuse2.py paste 2>&1 | tail -1
  SB_FUSE_DEBUG=$d timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:fuse_paste -s 28 -c 1 --csv python scratch/perf_fuse2.py paste 2>&1 | grep -E "fuse_paste" | awk -F'","' '{printf "%s=%s ", $(NF-2), $NF}'; echo
done
