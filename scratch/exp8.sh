#!/bin/bash
for w in 4 3 2; do
SB_REG_WAYS=$w python bench.py --no-e2e --no-cpu-baseline > gpurun_out/b8_$w.json 2> gpurun_out/b8.err; tail -3 gpurun_out/b8.err; python -c "
import json; d=json.load(open('gpurun_out/b8_$w.json')); print('ways=$w', {k:d[k] for k in ['value','ms_per_step','registration_ms_per_step','fusion_ms_per_step','tile_pairs_per_s','registration_truth_wells_ok','gpu_launches']})"
done
