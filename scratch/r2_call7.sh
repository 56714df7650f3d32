#!/bin/bash
# Round 2, GPU call 7: loader over 32 lanes, B stages on their own warp, branch-free inverse epilogue, half R / Y arrays.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_reg_tc_gpu.py tests/test_register_gpu.py tests/test_subpixel_gpu.py tests/test_configs_gpu.py -m gpu -q -x > $O/c7_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $O/c7_pytest.log
B="python bench.py --no-e2e --no-cpu-baseline --no-f64"
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[2], "reg_ms", round(d["registration_ms_per_step"],3), "fuse_ms", round(d["fusion_ms_per_step"],3), "step", round(d["ms_per_step"],3), d["registration_truth_wells_ok"], "redo", d["registration_f64_redo_pairs"])
except Exception as e: print("failed", sys.argv[2], e)
PY
}
timeout 300 $B > $O/c7_bench.json 2> $O/c7_bench.err; echo "bench rc=$?"; tail -3 $O/c7_bench.err; show $O/c7_bench.json default
for cfg in "768 4" "2048 2" "4096 1" "8192 2"; do
  set -- $cfg
  SB_REG_L2_MB=$1 SB_REG_WAYS=$2 timeout 300 $B > $O/c7_bench_l2_$1_w$2.json 2> $O/c7_bench_l2_$1_w$2.err
  show $O/c7_bench_l2_$1_w$2.json "L2_MB=$1 WAYS=$2"
done
for v in tcprof x_mma1 x_noepi x_nocv; do
  echo "== variant $v (one sub-batch per group)"
  TCP_WELLS=48 SB_REG_L2_MB=16384 SB_REG_WAYS=1 SB_LIB_PATH=image_stitcher_b200/_lib/variants/libstitchb200_$v.so timeout 300 python scratch/tc_profile.py > $O/c7_prof_$v.log 2>&1
  grep -A12 "^rep 1" $O/c7_prof_$v.log | grep -E "rep|cv_compute|ep_wait|mma_wait|cv_wait"
done
timeout 900 python -m pytest tests -m gpu -q -x > $O/c7_pytest_all.log 2>&1; echo "all rc=$?"; tail -5 $O/c7_pytest_all.log
