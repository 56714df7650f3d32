#!/bin/bash
# Round 2, GPU call 17: ncu source counters of the TC kernels after the cp.async ring.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
CMD="python bench.py --wells 24 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-f64"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"xdft_tc" -s 6 -c 3 -o $O/c17_tc $CMD > $O/c17_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $O/c17_ncu.log
ls -la $O/c17_tc.ncu-rep
ncu -i $O/c17_tc.ncu-rep --page raw --csv > $O/c17_tc_raw.csv 2>/dev/null
