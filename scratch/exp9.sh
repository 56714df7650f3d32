#!/bin/bash
for cfg in "3 384" "2 512" "4 256" "4 512" "1 256"; do
set -- $cfg
SB_REG_WAYS=$1 SB_REG_L2_MB=$2 python bench.py --no-e2e --no-cpu-baseline --steps 4 > gpurun_out/b9.json 2> gpurun_out/b9.err; tail -3 gpurun_out/b9.err; python -c "
import json; d=json.load(open('gpurun_out/b9.json')); print('ways=$1 l2=$2', {k:round(d[k],2) if isinstance(d[k],float) else d[k] for k in ['ms_per_step','registration_ms_per_step','fusion_ms_per_step','gpu_launches']})"
done
