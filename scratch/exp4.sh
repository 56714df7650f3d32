#!/bin/bash
python -m pytest tests -m gpu -x -q > gpurun_out/t4.log 2>&1; tail -3 gpurun_out/t4.log
python bench.py > gpurun_out/b4.json 2> gpurun_out/b4.err; tail -3 gpurun_out/b4.err; cat gpurun_out/b4.json
python bench.py --fuse-lanes 1 --no-e2e --no-cpu-baseline > gpurun_out/b4_l1.json 2> gpurun_out/b4_l1.err; cat gpurun_out/b4_l1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/b4_ref.json 2> gpurun_out/b4_ref.err; tail -3 gpurun_out/b4_ref.err; cat gpurun_out/b4_ref.json
nproc; free -g | head -2
