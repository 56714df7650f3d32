#!/bin/bash
# NOTE: record of an experiment -- the SB_REG_SMEM_KB / SB_FUSE_PERSIST knobs it used were removed again (no gain).
# concurrent registration + fusion: registration blocks per SM capped (SB_REG_SMEM_KB), fusion from a bounded persistent grid
for cfg in "100 1" "100 2" "120 2" "120 3" "0 5" "0 2"; do
  set -- $cfg
  SB_REG_SMEM_KB=$1 SB_FUSE_PERSIST=$2 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/b9.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/b9.json')); print('smem_kb $1 persist $2:', round(d['ms_per_step'],2), round(d['registration_ms_per_step'],2), round(d['fusion_ms_per_step'],2), 'overlapped', round(d['concurrent_phases']['ms_per_step'],2), d['concurrent_phases']['same_shifts_as_sequential'])"
done
