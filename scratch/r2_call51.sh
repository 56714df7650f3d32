#!/bin/bash
# Round 2, GPU call 51: blend_cells_kernel at higher occupancy (NX = groups of 4 pixels per lane in flight, M = min blocks / SM).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --config 3 --wells 192 --steps 3 --warmup 2 --no-cpu-baseline --no-f64"
run() {
  timeout 300 $B > $O/c51_bench_$1.json 2> $O/c51_bench_$1.err; rc=$?
  python - $1 $rc <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/c51_bench_{v}.json")); print(v, "rc", sys.argv[2], "fuse_ms", round(d["fusion_ms_per_step"],3), "frac", round(d["roofline"]["frac"],4), d["registration_truth_wells_ok"])
except Exception as e: print(v, "rc", sys.argv[2], "ERR", e)
PY
}
run shipped
for v in bx1m6 bx2m5 bx1m8 bx1m5; do SB_LIB_PATH=image_stitcher_b200/_lib/variants/libstitchb200_$v.so run $v; done
