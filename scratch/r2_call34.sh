#!/bin/bash
# Round 2, GPU call 34: blend_cells_kernel on packed pixel pairs; configs[3] bench.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_fuse_gpu.py tests/test_configs_gpu.py -m gpu -q -x > $O/c34_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/c34_pytest.log
timeout 600 python bench.py --config 3 --steps 3 --warmup 2 --no-cpu-baseline --no-f64 > $O/c34_bench_cfg3.json 2> $O/c34_bench_cfg3.err; echo "cfg3 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/c34_bench_cfg3.json')); print('cfg3 step', round(d['ms_per_step'],2), 'reg_ms', round(d['registration_ms_per_step'],2), 'fuse', round(d['fusion_ms_per_step'],2), 'frac', round(d['roofline']['frac'],3), d['registration_truth_wells_ok'])"
