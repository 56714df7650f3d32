#!/bin/bash
# Round 2, GPU call 25: ncu --set full of the ring form of paste_rect_kernel (why 18.5 ms?)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
CMD="python bench.py --wells 24 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-f64"
timeout 300 $CMD > $O/c25_plain.log 2>&1; echo "plain rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"paste_rect" -s 2 -c 1 -o $O/c25_ring $CMD > $O/c25_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i $O/c25_ring.ncu-rep --page raw --csv > $O/c25_ring_raw.csv 2>/dev/null
ncu -i $O/c25_ring.ncu-rep --page details > $O/c25_ring_details.txt 2>/dev/null
ncu -i $O/c25_ring.ncu-rep --page source --csv > $O/c25_ring_src.csv 2>/dev/null
ls -la $O/c25_ring.ncu-rep
