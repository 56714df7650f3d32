#!/bin/bash
# Round-1 profile collection (run under gpurun, one GPU).  Every ncu pass follows a plain run of the same command.
set -x
CMD="python bench.py --wells 12 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_final.csv $CMD > gpurun_out/ncu_launches.log 2>&1
CMD3="python bench.py --wells 3 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --fuse-lanes 1"
$CMD3 > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:paste_rect -s 3 -c 1 -f -o gpurun_out/prof_r1_fuse $CMD3 > gpurun_out/ncu_fuse.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"rows_fwd|cols_xpower|rows_inv|updft_rows|tile_minmax" -s 5 -c 5 -f -o gpurun_out/prof_r1_reg $CMD3 > gpurun_out/ncu_reg.log 2>&1
ls -la gpurun_out/prof_r1_*.ncu-rep
