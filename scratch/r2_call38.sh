#!/bin/bash
# Round 2, GPU call 38: the round's evidence on the shipped build -- full suite, smoke, bench (all legs), bench lines of
# the other BASELINE configs, ncu launch list and ncu --set full of the hot kernels.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > $O/c38_pytest_all.log 2>&1; echo "all rc=$?"; tail -3 $O/c38_pytest_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/c38_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/c38_smoke.log
timeout 900 python bench.py > $O/c38_bench.json 2> $O/c38_bench.err; echo "bench rc=$?"; tail -2 $O/c38_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/c38_bench_reference.json 2> $O/c38_bench_reference.err; echo "reference rc=$?"
for c in 0 1 3 4; do
  timeout 900 python bench.py --config $c --no-cpu-baseline --steps 3 > $O/c38_bench_cfg$c.json 2> $O/c38_bench_cfg$c.err; echo "cfg$c rc=$?"; tail -2 $O/c38_bench_cfg$c.err
done
CMD="python bench.py --wells 24 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-f64"
timeout 300 $CMD > $O/c38_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c38_launches.csv $CMD > $O/c38_ncu1.log 2>&1
python scripts/summarize_launches.py $O/c38_launches.csv $O/c38_launches_own > $O/c38_sum.log 2>&1; rm -f $O/c38_launches.csv
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"paste_rect|cols_warp|tile_minmax|xdft_tc|updft_cols|updft_tables" -s 28 -c 14 -o $O/c38_hot $CMD > $O/c38_ncu2.log 2>&1; echo "ncu full rc=$?"
ncu -i $O/c38_hot.ncu-rep --page raw --csv > $O/c38_hot_raw.csv 2>/dev/null
ls -la $O/c38_hot.ncu-rep
python - <<'PY'
import json
for n in ["bench","bench_cfg0","bench_cfg1","bench_cfg3","bench_cfg4"]:
    try:
        d=json.load(open(f"gpurun_out/c38_{n}.json"))
        print(n,"value",round(d["value"]),"ms",round(d["ms_per_step"],3),"reg",round(d["registration_ms_per_step"],3),"fuse",round(d["fusion_ms_per_step"],3),"frac",round(d["roofline"]["frac"],3),"launches",d["gpu_launches"], d.get("registration_truth_wells_ok"), "e2e", d.get("e2e") and round(d["e2e"]["value"]), "f64", d.get("registration_f64") and round(d["registration_f64"]["ms_per_step"],2), "cpu", d.get("cpu_baseline") and round(d["cpu_baseline"]["value"]))
    except Exception as e: print(n,"failed",e)
PY
