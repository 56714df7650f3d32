#!/bin/bash
# Round 2, GPU call 2: (1) everything on the radix registration path, (2) paste kernel with the region loop at several
# regions-per-block settings, (3) tcgen05 building blocks, (4) tensor-core registration stages, (5) full suite + bench on them.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
SB_REG_NO_TC=1 timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_reg_tc_gpu.py > $O/c2_pytest_radix.log 2>&1; echo "radix pytest rc=$?"; tail -3 $O/c2_pytest_radix.log
for j in 1 4 16 48; do
  SB_REG_NO_TC=1 SB_RECT_JOBS=$j timeout 300 python bench.py --no-e2e --no-cpu-baseline > $O/c2_bench_jobs$j.json 2> $O/c2_bench_jobs$j.err
done
SB_REG_NO_TC=1 timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-flatfield > $O/c2_bench_noflat.json 2> $O/c2_bench_noflat.err
timeout 300 python -m pytest tests/test_exhaustive_gpu.py -m gpu -x -q -s -k "tensor_core" > $O/c2_pytest_umma.log 2>&1; echo "umma rc=$?"; tail -5 $O/c2_pytest_umma.log
timeout 600 python -m pytest tests/test_reg_tc_gpu.py -m gpu -q -s > $O/c2_pytest_tc.log 2>&1; echo "tc stages rc=$?"; tail -15 $O/c2_pytest_tc.log
timeout 900 python -m pytest tests -m gpu -x -q > $O/c2_pytest_all.log 2>&1; echo "all rc=$?"; tail -5 $O/c2_pytest_all.log
timeout 600 python bench.py > $O/c2_bench.json 2> $O/c2_bench.err; echo "bench rc=$?"
CMD="python bench.py --wells 24 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > $O/c2_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c2_launches.csv $CMD > $O/c2_ncu1.log 2>&1
python scripts/summarize_launches.py $O/c2_launches.csv $O/c2_launches_own > $O/c2_sum.log 2>&1; rm -f $O/c2_launches.csv
timeout 900 ncu --set full --clock-control none -k regex:"paste_rect|fwd_x_tc|cols_warp" -s 12 -c 8 -o /tmp/c2_prof $CMD > $O/c2_ncu2.log 2>&1
ncu -i /tmp/c2_prof.ncu-rep --page raw --csv > $O/c2_prof_raw.csv 2>/dev/null
du -sh $O
