#!/bin/bash
# Round 2, GPU call 32: radix-3 / radix-5 butterflies in the shared-memory FFT engine (1500-long axis of configs[4]).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_configs_gpu.py tests/test_subpixel_gpu.py tests/test_register_gpu.py -m gpu -q -x > $O/c32_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/c32_pytest.log
timeout 400 python bench.py --config 4 --steps 3 --warmup 2 --no-cpu-baseline > $O/c32_bench_cfg4.json 2> $O/c32_bench_cfg4.err; echo "cfg4 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/c32_bench_cfg4.json')); print('cfg4 step', round(d['ms_per_step'],2), 'reg_ms', round(d['registration_ms_per_step'],2), 'fuse', round(d['fusion_ms_per_step'],2), d['registration_truth_wells_ok'], 'redo', d['registration_f64_redo_pairs'], 'f64', d.get('registration_f64'))"
