#!/bin/bash
for il in 8 16 32; do for lanes in 3 2; do
SB_FUSE_INTERLEAVE=$il python bench.py --no-e2e --no-cpu-baseline --steps 6 --fuse-lanes $lanes > gpurun_out/b10.json 2> gpurun_out/b10.err; tail -3 gpurun_out/b10.err; python -c "
import json; d=json.load(open('gpurun_out/b10.json')); print('il=$il lanes=$lanes', {k:round(d[k],2) if isinstance(d[k],float) else d[k] for k in ['ms_per_step','registration_ms_per_step','fusion_ms_per_step']}, round(d['roofline']['frac'],3))"
done; done
