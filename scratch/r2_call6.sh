#!/bin/bash
# Round 2, GPU call 6: 16 converter warps + RegisterPlan; sub-batch size sweep; what bounds the TC kernels (variants).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_reg_tc_gpu.py tests/test_stitcher_process_gpu.py tests/test_register_gpu.py -m gpu -q -x > $O/c6_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/c6_pytest.log
B="python bench.py --no-e2e --no-cpu-baseline --no-f64"
timeout 300 $B > $O/c6_bench.json 2> $O/c6_bench.err; echo "bench rc=$?"; tail -3 $O/c6_bench.err
for cfg in "1536 4" "3072 4" "3072 2" "4608 2" "4608 1" "9216 1" "1536 2"; do
  set -- $cfg
  SB_REG_L2_MB=$1 SB_REG_WAYS=$2 timeout 300 $B > $O/c6_bench_l2_$1_w$2.json 2> $O/c6_bench_l2_$1_w$2.err
  python - "$O/c6_bench_l2_$1_w$2.json" "$1 $2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print("L2_MB WAYS", sys.argv[2], "reg_ms", round(d["registration_ms_per_step"],3), "fuse_ms", round(d["fusion_ms_per_step"],3), d["registration_truth_wells_ok"])
except Exception as e: print("failed", sys.argv[2], e)
PY
done
for v in tcprof x_mma1 x_noepi x_nocv; do
  echo "== variant $v (one sub-batch per group)"
  TCP_WELLS=48 SB_REG_L2_MB=16384 SB_REG_WAYS=1 SB_LIB_PATH=image_stitcher_b200/_lib/variants/libstitchb200_$v.so timeout 300 python scratch/tc_profile.py > $O/c6_prof_$v.log 2>&1
  grep -A12 "^rep 1" $O/c6_prof_$v.log | grep -E "rep|cv_compute|ep_wait|mma_wait_ab_full"
done
