#!/bin/bash
# Round 2, GPU call 37: next-wave L2 prefetch in paste_rect_kernel (distance sweep), with / without the per-row prefetch.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_fuse_gpu.py tests/test_configs_gpu.py -m gpu -q -x > $O/c37_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/c37_pytest.log
B="python bench.py --no-e2e --no-cpu-baseline --no-f64"
run() {
  timeout 300 $B > $O/c37_bench_$1.json 2> $O/c37_bench_$1.err; rc=$?
  python - $1 $rc <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/c37_bench_{v}.json")); print(v, "rc", sys.argv[2], "fuse_ms", round(d["fusion_ms_per_step"],3), "frac", round(d["roofline"]["frac"],4), "coord-only", round(d["fusion_coordinate_only"]["ms_per_step"],3))
except Exception as e: print(v, "rc", sys.argv[2], "ERR", e)
PY
}
run default
for pf in 0 370 1480 2960 5920; do SB_RECT_PF=$pf run pf$pf; done
export SB_LIB_PATH=image_stitcher_b200/_lib/variants/libstitchb200_nopf.so
run norow_default
SB_RECT_PF=1480 run norow_pf1480
SB_RECT_PF=0 run norow_pf0
