#!/bin/bash
# re-tune the registration sub-batching (streams in flight, spectra budget) for the 3-blocks-per-SM build
for cfg in "4 640" "4 896" "4 1024" "3 768" "4 1536" "3 1152"; do
  set -- $cfg
  SB_REG_WAYS=$1 SB_REG_L2_MB=$2 python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/b10.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/b10.json')); print('ways $1 budget_mb $2:', round(d['ms_per_step'],2), 'reg', round(d['registration_ms_per_step'],2), 'fuse', round(d['fusion_ms_per_step'],2))"
done
