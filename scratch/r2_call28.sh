#!/bin/bash
# Round 2, GPU call 28: paste path with the denormal-scale pack (FMUL2.RZ by 2^-149) and the short divide (no Newton step),
# proved over all 25 binades by sb_selftest; A/B against the r2 shipped arithmetic; the coordinate-only bench leg.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_exhaustive_gpu.py -m gpu -q --durations=5 > $O/c28_pytest_exh.log 2>&1; echo "exhaustive rc=$?"; tail -12 $O/c28_pytest_exh.log
timeout 900 python -m pytest tests/test_fuse_gpu.py tests/test_configs_gpu.py tests/test_stitcher_process_gpu.py -m gpu -q -x > $O/c28_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/c28_pytest.log
B="python bench.py --no-e2e --no-cpu-baseline --no-f64"
for v in main r2ship packonly main r2ship; do
  if [ $v = main ]; then unset SB_LIB_PATH; else export SB_LIB_PATH=image_stitcher_b200/_lib/variants/libstitchb200_$v.so; fi
  timeout 300 $B > $O/c28_bench_$v.json 2> $O/c28_bench_$v.err; rc=$?
  python - $v $rc <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/c28_bench_{v}.json")); print(v, "rc", sys.argv[2], "reg_ms", round(d["registration_ms_per_step"],3), "fuse_ms", round(d["fusion_ms_per_step"],3), "frac", round(d["roofline"]["frac"],4), "coord-only", d.get("fusion_coordinate_only"))
except Exception as e: print(v, "rc", sys.argv[2], "ERR", e)
PY
done
