#!/bin/bash
# Round 2, GPU call 48: paste kernel with fewer groups per chunk and more resident warps (G = max groups, M = min blocks / SM).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline --no-f64"
run() {
  timeout 300 $B > $O/c48_bench_$1.json 2> $O/c48_bench_$1.err; rc=$?
  python - $1 $rc <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/c48_bench_{v}.json")); print(v, "rc", sys.argv[2], "fuse_ms", round(d["fusion_ms_per_step"],3), "frac", round(d["roofline"]["frac"],4), "coord-only", round(d["fusion_coordinate_only"]["ms_per_step"],3), d["registration_truth_wells_ok"])
except Exception as e: print(v, "rc", sys.argv[2], "ERR", e)
PY
}
run shipped
for v in g2m6 g1m8 g2m7 g2m5; do SB_LIB_PATH=image_stitcher_b200/_lib/variants/libstitchb200_$v.so run $v; done
