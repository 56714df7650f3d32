#!/bin/bash
# Round 2, GPU call 36: flat-field of a vector by ONE 256-bit load (LDG.E.256) instead of two half-line 128-bit loads.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_fuse_gpu.py tests/test_configs_gpu.py tests/test_stitcher_process_gpu.py -m gpu -q -x > $O/c36_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/c36_pytest.log
B="python bench.py --no-e2e --no-cpu-baseline --no-f64"
for v in a b; do
  timeout 300 $B > $O/c36_bench_$v.json 2> $O/c36_bench_$v.err; rc=$?
  python - $v $rc <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/c36_bench_{v}.json")); print(v, "rc", sys.argv[2], "reg_ms", round(d["registration_ms_per_step"],3), "fuse_ms", round(d["fusion_ms_per_step"],3), "frac", round(d["roofline"]["frac"],4), "coord-only", round(d["fusion_coordinate_only"]["ms_per_step"],3))
except Exception as e: print(v, "rc", sys.argv[2], "ERR", e)
PY
done
