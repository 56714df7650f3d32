#!/bin/bash
# Round 2, GPU call 14: cp.async prefetch ring in the converters of modes 1 and 2.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_reg_tc_gpu.py tests/test_register_gpu.py tests/test_subpixel_gpu.py -m gpu -q -x > $O/c14_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/c14_pytest.log
B="python bench.py --no-e2e --no-cpu-baseline --no-f64"
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[2], "reg_ms", round(d["registration_ms_per_step"],3), "fuse_ms", round(d["fusion_ms_per_step"],3), "step", round(d["ms_per_step"],3), d["registration_truth_wells_ok"], "redo", d["registration_f64_redo_pairs"])
except Exception as e: print("failed", sys.argv[2], e)
PY
}
timeout 300 $B > $O/c14_bench.json 2> $O/c14_bench.err; echo "bench rc=$?"; tail -3 $O/c14_bench.err; show $O/c14_bench.json default
TCP_WELLS=48 SB_REG_L2_MB=16384 SB_REG_WAYS=1 SB_LIB_PATH=image_stitcher_b200/_lib/variants/libstitchb200_tcprof.so timeout 300 python scratch/tc_profile.py > $O/c14_prof_tcprof.log 2>&1
grep -A12 "^rep 1" $O/c14_prof_tcprof.log
