#!/bin/bash
# Round 2, GPU call 23: multiply-high stretch in the forward converters (magic delivered with the tile record).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_reg_tc_gpu.py tests/test_register_gpu.py tests/test_subpixel_gpu.py tests/test_exhaustive_gpu.py tests/test_configs_gpu.py -m gpu -q -x > $O/c23_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/c23_pytest.log
B="python bench.py --no-e2e --no-cpu-baseline"
timeout 300 $B > $O/c23_bench.json 2> $O/c23_bench.err; echo "bench rc=$?"; tail -3 $O/c23_bench.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/c23_bench.json")); print("reg_ms", round(d["registration_ms_per_step"],3), "fuse_ms", round(d["fusion_ms_per_step"],3), "step", round(d["ms_per_step"],3), d["registration_truth_wells_ok"], "redo", d["registration_f64_redo_pairs"], "f64", d["registration_f64"])
PY
TCP_WELLS=48 SB_REG_L2_MB=16384 SB_REG_WAYS=1 SB_LIB_PATH=image_stitcher_b200/_lib/variants/libstitchb200_tcprof.so timeout 300 python scratch/tc_profile.py > $O/c23_prof_tcprof.log 2>&1
grep -A12 "^rep 1 mode 0" $O/c23_prof_tcprof.log
