#!/bin/bash
# Round 2, GPU call 40: ncu --set full of blend_cells_kernel and paste_rect_kernel<0,1> on a 48-well configs[3] batch.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
CMD="python bench.py --config 3 --wells 48 --steps 1 --warmup 1 --no-cpu-baseline --no-f64"
timeout 300 $CMD > $O/c40_plain.json 2> $O/c40_plain.err; echo "plain rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"blend_cells|paste_rect" -s 2 -c 2 -o $O/c40_blend $CMD > $O/c40_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i $O/c40_blend.ncu-rep --page raw --csv > $O/c40_blend_raw.csv 2>/dev/null
ncu -i $O/c40_blend.ncu-rep --page details > $O/c40_blend_details.txt 2>/dev/null
ncu -i $O/c40_blend.ncu-rep --page source --csv > $O/c40_blend_src.csv 2>/dev/null
ls -la $O/c40_blend.ncu-rep
