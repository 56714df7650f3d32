#!/bin/bash
# Round 2, GPU call 26: paste_rect_kernel flat-field path as a per-lane cp.async ring (depth 2/3/4, rows 2/4) vs the register-load form.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_fuse_gpu.py tests/test_configs_gpu.py tests/test_stitcher_process_gpu.py -m gpu -q -x > $O/c26_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/c26_pytest.log
B="python bench.py --no-e2e --no-cpu-baseline --no-f64"
for v in main ring0 ring3m0 ring2 ring4 ring3r4 main; do
  if [ $v = main ]; then unset SB_LIB_PATH; else export SB_LIB_PATH=image_stitcher_b200/_lib/variants/libstitchb200_$v.so; fi
  timeout 300 $B > $O/c26_bench_$v.json 2> $O/c26_bench_$v.err; rc=$?
  python - $v $rc <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/c26_bench_{v}.json")); print(v, "rc", sys.argv[2], "reg_ms", round(d["registration_ms_per_step"],3), "fuse_ms", round(d["fusion_ms_per_step"],3), "frac", round(d["roofline"]["frac"],4))
except Exception as e: print(v, "rc", sys.argv[2], "ERR", e)
PY
done
