#!/bin/bash
# Round 2, GPU call 9: magic-number stretch in the forward converters; full suite; launch list of the shipped build.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/c9_pytest_all.log 2>&1; echo "all rc=$?"; tail -5 $O/c9_pytest_all.log
B="python bench.py --no-e2e --no-cpu-baseline --no-f64"
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[2], "reg_ms", round(d["registration_ms_per_step"],3), "fuse_ms", round(d["fusion_ms_per_step"],3), "step", round(d["ms_per_step"],3), d["registration_truth_wells_ok"], "redo", d["registration_f64_redo_pairs"])
except Exception as e: print("failed", sys.argv[2], e)
PY
}
timeout 300 $B > $O/c9_bench.json 2> $O/c9_bench.err; echo "bench rc=$?"; tail -3 $O/c9_bench.err; show $O/c9_bench.json default
TCP_WELLS=48 SB_REG_L2_MB=16384 SB_REG_WAYS=1 SB_LIB_PATH=image_stitcher_b200/_lib/variants/libstitchb200_tcprof.so timeout 300 python scratch/tc_profile.py > $O/c9_prof_tcprof.log 2>&1
grep -A12 "^rep 1" $O/c9_prof_tcprof.log
CMD="python bench.py --wells 24 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-f64"
timeout 300 $CMD > $O/c9_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c9_launches.csv $CMD > $O/c9_ncu1.log 2>&1
python scripts/summarize_launches.py $O/c9_launches.csv $O/c9_launches_own > $O/c9_sum.log 2>&1; rm -f $O/c9_launches.csv
cat $O/c9_launches_own.md
