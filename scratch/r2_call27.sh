#!/bin/bash
# Round 2, GPU call 27: flat-field path with the next group held in registers (rect_band_pipe) vs the r2 shipped form.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_fuse_gpu.py tests/test_configs_gpu.py tests/test_stitcher_process_gpu.py -m gpu -q -x > $O/c27_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/c27_pytest.log
B="python bench.py --no-e2e --no-cpu-baseline --no-f64"
for v in main old pipe5 pipe4r4 pipe4r1 pipe3 main; do
  if [ $v = main ]; then unset SB_LIB_PATH; else export SB_LIB_PATH=image_stitcher_b200/_lib/variants/libstitchb200_$v.so; fi
  timeout 300 $B > $O/c27_bench_$v.json 2> $O/c27_bench_$v.err; rc=$?
  python - $v $rc <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/c27_bench_{v}.json")); print(v, "rc", sys.argv[2], "reg_ms", round(d["registration_ms_per_step"],3), "fuse_ms", round(d["fusion_ms_per_step"],3), "frac", round(d["roofline"]["frac"],4))
except Exception as e: print(v, "rc", sys.argv[2], "ERR", e)
PY
done
