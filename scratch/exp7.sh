#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/t7.log 2>&1; tail -3 gpurun_out/t7.log
python bench.py > gpurun_out/b7.json 2> gpurun_out/b7.err; tail -3 gpurun_out/b7.err; python -c "
import json; d=json.load(open('gpurun_out/b7.json')); print({k:d[k] for k in ['value','ms_per_step','registration_ms_per_step','fusion_ms_per_step','tile_pairs_per_s','registration_truth_wells_ok','registration_f64_redo_pairs','gpu_launches']}); print(d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['cpu_baseline']['value'])"
