#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/t3.log 2>&1; tail -3 gpurun_out/t3.log
python bench.py --wells 12 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/b3_small.json 2>gpurun_out/b3_small.err; tail -2 gpurun_out/b3_small.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --wells 12 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_r1.log 2>&1
tail -2 gpurun_out/ncu_r1.log
