#!/bin/bash
for v in libstitchb200 $(cd image_stitcher_b200/_lib; ls var_*.so | sed 's/.so$//'); do
  export SB_LIB_PATH=/root/repo/image_stitcher_b200/_lib/$v.so
  python bench.py --no-e2e --no-cpu-baseline --steps 4 > gpurun_out/bs.json 2> gpurun_out/bs.err; tail -2 gpurun_out/bs.err; python -c "
import json; d=json.load(open('gpurun_out/bs.json')); print('$v', {k:round(d[k],2) if isinstance(d[k],float) else d[k] for k in ['ms_per_step','registration_ms_per_step','fusion_ms_per_step','registration_truth_wells_ok']})"
done
