#!/bin/bash
# Round 2, GPU call 18: paste_rect_jobs_kernel (flat-field chunk parked in shared memory, regions stream past it).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_fuse_gpu.py tests/test_configs_gpu.py -m gpu -q -x > $O/c18_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/c18_pytest.log
B="python bench.py --no-e2e --no-cpu-baseline --no-f64"
show() { python - "$1" "$2" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(sys.argv[2], "reg_ms", round(d["registration_ms_per_step"],3), "fuse_ms", round(d["fusion_ms_per_step"],3), "frac", round(d["roofline"]["frac"],3), "step", round(d["ms_per_step"],3), d["registration_truth_wells_ok"])
except Exception as e: print("failed", sys.argv[2], e)
PY
}
for j in 8 1 2 4 16 32 96; do
  SB_RECT_JOBS=$j timeout 300 $B > $O/c18_bench_j$j.json 2> $O/c18_bench_j$j.err; show $O/c18_bench_j$j.json "jobs=$j"
done
for v in jobs_b5 jobs_b6; do for j in 4 8 16; do
  SB_LIB_PATH=image_stitcher_b200/_lib/variants/libstitchb200_$v.so SB_RECT_JOBS=$j timeout 300 $B > $O/c18_bench_${v}_j$j.json 2> $O/c18_bench_${v}_j$j.err; show $O/c18_bench_${v}_j$j.json "$v jobs=$j"
done; done
