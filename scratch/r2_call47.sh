#!/bin/bash
# Round 2, GPU call 47: what the flat-field costs -- its loads without the divide (diag1) vs the divide without its loads (diag2).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline --no-f64"
run() {
  timeout 300 $B > $O/c47_bench_$1.json 2> $O/c47_bench_$1.err; rc=$?
  python - $1 $rc <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/c47_bench_{v}.json")); print(v, "rc", sys.argv[2], "fuse_ms", round(d["fusion_ms_per_step"],3), "frac", round(d["roofline"]["frac"],4), "coord-only", round(d["fusion_coordinate_only"]["ms_per_step"],3))
except Exception as e: print(v, "rc", sys.argv[2], "ERR", e)
PY
}
run shipped
SB_LIB_PATH=image_stitcher_b200/_lib/variants/libstitchb200_diag1.so run loads_only
SB_LIB_PATH=image_stitcher_b200/_lib/variants/libstitchb200_diag2.so run divide_only
