#!/bin/bash
# Round 2, GPU call 12: ncu --set full with source counters for the three tensor-core kernels (shipped build).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
CMD="python bench.py --wells 24 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-f64"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"xdft_tc" -s 6 -c 3 -o $O/c12_tc $CMD > $O/c12_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $O/c12_ncu.log
ls -la $O/c12_tc.ncu-rep
ncu -i $O/c12_tc.ncu-rep --page raw --csv > $O/c12_tc_raw.csv 2>/dev/null
