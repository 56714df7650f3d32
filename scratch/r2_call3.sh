#!/bin/bash
# Round 2, GPU call 3: persistent warp-specialised tensor-core kernels (forward + inverse), paste kernel variants.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_reg_tc_gpu.py -m gpu -q -s > $O/c3_pytest_tc.log 2>&1; echo "tc stages rc=$?"; tail -8 $O/c3_pytest_tc.log
SB_REG_NO_TC_INV=1 timeout 600 python -m pytest tests/test_reg_tc_gpu.py tests/test_subpixel_gpu.py -m gpu -q > $O/c3_pytest_tc_fwdonly.log 2>&1; echo "tc fwd-only rc=$?"; tail -3 $O/c3_pytest_tc_fwdonly.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/c3_pytest_all.log 2>&1; echo "all rc=$?"; tail -5 $O/c3_pytest_all.log
timeout 600 python bench.py > $O/c3_bench.json 2> $O/c3_bench.err; echo "bench rc=$?"
SB_REG_NO_TC_INV=1 timeout 300 python bench.py --no-e2e --no-cpu-baseline > $O/c3_bench_fwdonly.json 2> $O/c3_bench_fwdonly.err
for v in minb6 warps4 pf0 pf4; do
  SB_LIB_PATH=image_stitcher_b200/_lib/variants/libstitchb200_$v.so timeout 300 python bench.py --no-e2e --no-cpu-baseline > $O/c3_bench_$v.json 2> $O/c3_bench_$v.err
done
CMD="python bench.py --wells 24 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > $O/c3_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c3_launches.csv $CMD > $O/c3_ncu1.log 2>&1
python scripts/summarize_launches.py $O/c3_launches.csv $O/c3_launches_own > $O/c3_sum.log 2>&1; rm -f $O/c3_launches.csv
timeout 900 ncu --set full --clock-control none -k regex:"xdft_tc|cols_warp|updft_rows|tile_minmax" -s 10 -c 10 -o /tmp/c3_prof $CMD > $O/c3_ncu2.log 2>&1
ncu -i /tmp/c3_prof.ncu-rep --page raw --csv > $O/c3_prof_raw.csv 2>/dev/null
du -sh $O
