#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/t6.log 2>&1; tail -3 gpurun_out/t6.log
python bench.py --no-cpu-baseline > gpurun_out/b6.json 2> gpurun_out/b6.err; tail -3 gpurun_out/b6.err; python -c "
import json; d=json.load(open('gpurun_out/b6.json')); print({k:d[k] for k in ['value','ms_per_step','registration_ms_per_step','fusion_ms_per_step','registration_truth_wells_ok']}); print(d['e2e'])"
