#!/bin/bash
# Round 2, GPU call 53: final check of the committed tree (full GPU suite, smoke, default bench) + ncu launch list of the shipped build.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > $O/c53_pytest_all.log 2>&1; echo "all rc=$?"; tail -3 $O/c53_pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/c53_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/c53_smoke.log
timeout 900 python bench.py > $O/c53_bench.json 2> $O/c53_bench.err; echo "bench rc=$?"; tail -2 $O/c53_bench.err
CMD="python bench.py --wells 24 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-f64"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c53_launches.csv $CMD > $O/c53_ncu1.log 2>&1
python scripts/summarize_launches.py $O/c53_launches.csv $O/c53_launches_own > $O/c53_sum.log 2>&1; rm -f $O/c53_launches.csv
python - <<'PY'
import json
d=json.load(open("gpurun_out/c53_bench.json"))
print("value",round(d["value"]),"ms",round(d["ms_per_step"],3),"reg",round(d["registration_ms_per_step"],3),"fuse",round(d["fusion_ms_per_step"],3),"frac",round(d["roofline"]["frac"],3),"coord",round(d["fusion_coordinate_only"]["frac"],3),"launches",d["gpu_launches"], d.get("registration_truth_wells_ok"), "e2e", round(d["e2e"]["value"]), d["e2e"]["matches_device_result"], "f64", round(d["registration_f64"]["ms_per_step"],2), "cpu", round(d["cpu_baseline"]["value"]), d["clocks"]["reasons"])
PY
head -8 $O/c53_launches_own.md
