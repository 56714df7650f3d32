#!/bin/bash
# Round 2, GPU call 41: blend_cells_kernel with the tile list in shared memory and the tiles loaded two at a time.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_fuse_gpu.py tests/test_configs_gpu.py -m gpu -q -x > $O/c41_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/c41_pytest.log
for w in 192 384; do
timeout 400 python bench.py --config 3 --wells $w --steps 3 --warmup 2 --no-cpu-baseline --no-f64 > $O/c41_bench_$w.json 2> $O/c41_bench_$w.err; rc=$?
python - $w $rc <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/c41_bench_{v}.json")); print(v, "rc", sys.argv[2], "step", round(d["ms_per_step"],2), "fuse_ms", round(d["fusion_ms_per_step"],3), "frac", round(d["roofline"]["frac"],4), d["registration_truth_wells_ok"])
except Exception as e: print(v, "rc", sys.argv[2], "ERR", e)
PY
done
