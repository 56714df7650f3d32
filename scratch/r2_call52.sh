#!/bin/bash
# Round 2, GPU call 52: shipped blend_cells_kernel = one group per lane, 32 registers, 8 blocks per SM: tests + configs[3] / [0] lines.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_fuse_gpu.py tests/test_configs_gpu.py tests/test_stitcher_process_gpu.py -m gpu -q -x > $O/c52_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/c52_pytest.log
timeout 600 python bench.py --config 3 --steps 3 --no-cpu-baseline --no-f64 > $O/c52_bench_cfg3.json 2> $O/c52_bench_cfg3.err; echo "cfg3 rc=$?"
timeout 600 python bench.py --config 0 --steps 3 --no-cpu-baseline --no-f64 > $O/c52_bench_cfg0.json 2> $O/c52_bench_cfg0.err; echo "cfg0 rc=$?"
python - <<'PY'
import json
for n in ["bench_cfg3","bench_cfg0"]:
    try:
        d=json.load(open(f"gpurun_out/c52_{n}.json"))
        print(n,"value",round(d["value"]),"ms",round(d["ms_per_step"],3),"reg",round(d["registration_ms_per_step"],3),"fuse",round(d["fusion_ms_per_step"],3),"frac",round(d["roofline"]["frac"],3), d.get("registration_truth_wells_ok"))
    except Exception as e: print(n,"failed",e)
PY
