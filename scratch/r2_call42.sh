#!/bin/bash
# Round 2, GPU call 42: distance of the per-row L2 bulk prefetch of paste_rect_kernel (SB_RECT_PREFETCH blocks of 16 rows ahead).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --no-e2e --no-cpu-baseline --no-f64"
run() {
  timeout 300 $B > $O/c42_bench_$1.json 2> $O/c42_bench_$1.err; rc=$?
  python - $1 $rc <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/c42_bench_{v}.json")); print(v, "rc", sys.argv[2], "fuse_ms", round(d["fusion_ms_per_step"],3), "frac", round(d["roofline"]["frac"],4), "coord-only", round(d["fusion_coordinate_only"]["ms_per_step"],3))
except Exception as e: print(v, "rc", sys.argv[2], "ERR", e)
PY
}
run d2_shipped
for d in 1 3 4 8 16; do SB_LIB_PATH=image_stitcher_b200/_lib/variants/libstitchb200_pfd$d.so run d$d; done
run d2_shipped_again
