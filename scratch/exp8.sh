#!/bin/bash
# NOTE: record of an experiment -- the SB_REG_SMEM_KB / SB_FUSE_PERSIST knobs it used were removed again (no gain).
# concurrent registration + fusion with the registration blocks per SM capped through the shared-memory request
for kb in 0 100 120; do
  SB_REG_SMEM_KB=$kb python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/b8_$kb.json 2>/dev/null
  python -c "
import json; d=json.load(open('gpurun_out/b8_$kb.json')); print($kb, round(d['ms_per_step'],2), round(d['registration_ms_per_step'],2), round(d['fusion_ms_per_step'],2), 'overlapped', round(d['concurrent_phases']['ms_per_step'],2), d['concurrent_phases']['same_shifts_as_sequential'])"
done
