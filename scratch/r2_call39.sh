#!/bin/bash
# Round 2, GPU call 39: blend_cells_kernel with 1 / 2 / 4 / 8 rows per warp; configs[3] bench (fewer wells: geometry is per well).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_fuse_gpu.py -m gpu -q -x > $O/c39_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/c39_pytest.log
B="python bench.py --config 3 --wells 192 --steps 3 --warmup 2 --no-cpu-baseline --no-f64"
run() {
  timeout 300 $B > $O/c39_bench_$1.json 2> $O/c39_bench_$1.err; rc=$?
  python - $1 $rc <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/c39_bench_{v}.json")); print(v, "rc", sys.argv[2], "fuse_ms", round(d["fusion_ms_per_step"],3), "frac", round(d["roofline"]["frac"],4))
except Exception as e: print(v, "rc", sys.argv[2], "ERR", e)
PY
}
run rows4
for r in 1 2 8; do SB_LIB_PATH=image_stitcher_b200/_lib/variants/libstitchb200_br$r.so run rows$r; done
