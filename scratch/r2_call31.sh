#!/bin/bash
# Round 2, GPU call 31: where do the 55.8 ms of configs[4] registration (760 pairs of 1500 x 314, radix engine) go?
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
CMD="python bench.py --config 4 --steps 1 --warmup 1 --no-cpu-baseline --no-f64"
timeout 300 $CMD > $O/c31_plain.json 2> $O/c31_plain.err; echo "plain rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c31_launches.csv $CMD > $O/c31_ncu1.log 2>&1
python scripts/summarize_launches.py $O/c31_launches.csv $O/c31_launches_cfg4 > $O/c31_sum.log 2>&1; rm -f $O/c31_launches.csv
cat $O/c31_launches_cfg4.md | head -30
for cfg in "256 4" "768 4" "2048 2" "4096 2" "8192 1"; do
  set -- $cfg
  SB_REG_L2_MB=$1 SB_REG_WAYS=$2 timeout 300 python bench.py --config 4 --steps 3 --warmup 2 --no-cpu-baseline --no-f64 > $O/c31_b.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/c31_b.json')); print('L2_MB $1 ways $2 reg_ms', round(d['registration_ms_per_step'],2), 'fuse', round(d['fusion_ms_per_step'],2), d['registration_truth_wells_ok'])"
done
