#!/bin/bash
# Round 2, GPU call 29: warps of a paste block = the same rows in 8 regions (flat-field shared through L1).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_fuse_gpu.py tests/test_configs_gpu.py tests/test_stitcher_process_gpu.py -m gpu -q -x > $O/c29_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/c29_pytest.log
B="python bench.py --no-e2e --no-cpu-baseline --no-f64"
run() {
  timeout 300 $B > $O/c29_bench_$1.json 2> $O/c29_bench_$1.err; rc=$?
  python - $1 $rc <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/c29_bench_{v}.json")); print(v, "rc", sys.argv[2], "reg_ms", round(d["registration_ms_per_step"],3), "fuse_ms", round(d["fusion_ms_per_step"],3), "frac", round(d["roofline"]["frac"],4), "coord-only", round(d["fusion_coordinate_only"]["ms_per_step"],3))
except Exception as e: print(v, "rc", sys.argv[2], "ERR", e)
PY
}
run r8
SB_RECT_REGIONS=1 run r1
SB_RECT_REGIONS=2 run r2
SB_RECT_REGIONS=4 run r4
export SB_LIB_PATH=image_stitcher_b200/_lib/variants/libstitchb200_cap5.so
run cap5_r8
SB_RECT_REGIONS=1 run cap5_r1
SB_RECT_REGIONS=4 run cap5_r4
