#!/bin/bash
# Round 2, GPU call 33: launch list of the 3000^2-tile registration after the radix-3/5 butterflies (8 x 8 mosaic, 112 pairs).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
CMD="python bench.py --config 4 --grid 8 --num-z 1 --steps 1 --warmup 1 --no-cpu-baseline --no-f64"
timeout 300 $CMD > $O/c33_plain.json 2> $O/c33_plain.err; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c33_launches.csv $CMD > $O/c33_ncu1.log 2>&1
python scripts/summarize_launches.py $O/c33_launches.csv $O/c33_launches_cfg4_grid8 > $O/c33_sum.log 2>&1; rm -f $O/c33_launches.csv
cat $O/c33_launches_cfg4_grid8.md | head -20
python -c "
import json; d=json.load(open('gpurun_out/c33_plain.json')); print('grid8 step', round(d['ms_per_step'],2), 'reg_ms', round(d['registration_ms_per_step'],2), 'pairs', d['config']['pairs_per_step'])"
