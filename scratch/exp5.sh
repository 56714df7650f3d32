#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/t5.log 2>&1; tail -3 gpurun_out/t5.log
python bench.py --no-e2e --no-cpu-baseline > gpurun_out/b5.json 2> gpurun_out/b5.err; tail -3 gpurun_out/b5.err; python -c "
import json; d=json.load(open('gpurun_out/b5.json')); print({k:d[k] for k in ['value','ms_per_step','registration_ms_per_step','fusion_ms_per_step','tile_pairs_per_s','registration_truth_wells_ok','registration_f64_redo_pairs']})"
SB_REG_NO_GEMM=1 python bench.py --no-e2e --no-cpu-baseline --steps 3 > gpurun_out/b5_nogemm.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/b5_nogemm.json')); print('nogemm', {k:d[k] for k in ['registration_ms_per_step','tile_pairs_per_s']})"
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1b.csv python bench.py --wells 12 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r1b.log 2>&1
