#!/bin/bash
# Round 2, GPU call 21: blend modes batched over the regions of a plate; bench lines of the other BASELINE configs.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_fuse_gpu.py -m gpu -q -x > $O/c21_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/c21_pytest.log
for c in 3; do
  timeout 900 python bench.py --config $c --no-cpu-baseline --steps 3 > $O/c21_bench_cfg$c.json 2> $O/c21_bench_cfg$c.err; echo "cfg$c rc=$?"; tail -2 $O/c21_bench_cfg$c.err
done
python - <<'PY'
import json
for c in [3]:
    try:
        d=json.load(open(f"gpurun_out/c21_bench_cfg{c}.json"))
        print("cfg",c,"value",round(d["value"]),"ms",round(d["ms_per_step"],3),"reg",round(d["registration_ms_per_step"],3),"fuse",round(d["fusion_ms_per_step"],3),"frac",round(d["roofline"]["frac"],3),"launches",d["gpu_launches"], d["registration_truth_wells_ok"])
    except Exception as e: print(c,"failed",e)
PY
CMD="python bench.py --config 3 --wells 48 --steps 2 --warmup 1 --no-cpu-baseline --no-f64"
timeout 300 $CMD > $O/c21_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c21_launches.csv $CMD > $O/c21_ncu1.log 2>&1
python scripts/summarize_launches.py $O/c21_launches.csv $O/c21_launches_cfg3 > $O/c21_sum.log 2>&1; rm -f $O/c21_launches.csv
cat $O/c21_launches_cfg3.md
