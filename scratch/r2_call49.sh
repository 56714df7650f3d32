#!/bin/bash
# Round 2, GPU call 49: paste kernel at 48 warps per SM (2 groups per chunk with the field, 40 registers): rows per warp 2 / 4 / 1, 128-thread blocks.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline --no-f64"
run() {
  timeout 300 $B > $O/c49_bench_$1.json 2> $O/c49_bench_$1.err; rc=$?
  python - $1 $rc <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/c49_bench_{v}.json")); print(v, "rc", sys.argv[2], "fuse_ms", round(d["fusion_ms_per_step"],3), "frac", round(d["roofline"]["frac"],4), "coord-only", round(d["fusion_coordinate_only"]["ms_per_step"],3), d["registration_truth_wells_ok"])
except Exception as e: print(v, "rc", sys.argv[2], "ERR", e)
PY
}
run shipped
for v in ha hb hc hd ha; do SB_LIB_PATH=image_stitcher_b200/_lib/variants/libstitchb200_$v.so run $v; done
