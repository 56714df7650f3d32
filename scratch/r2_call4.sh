#!/bin/bash
# Round 2, GPU call 4: branch-free converters + deeper prefetch; bench lines of the other BASELINE configs.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > $O/c4_pytest_all.log 2>&1; echo "all rc=$?"; tail -5 $O/c4_pytest_all.log
timeout 600 python bench.py > $O/c4_bench.json 2> $O/c4_bench.err; echo "bench rc=$?"; tail -3 $O/c4_bench.err
timeout 600 python bench.py --config 1 --no-cpu-baseline > $O/c4_bench_cfg1.json 2> $O/c4_bench_cfg1.err; echo "cfg1 rc=$?"; tail -3 $O/c4_bench_cfg1.err
timeout 600 python bench.py --config 0 --no-cpu-baseline > $O/c4_bench_cfg0.json 2> $O/c4_bench_cfg0.err; echo "cfg0 rc=$?"; tail -3 $O/c4_bench_cfg0.err
timeout 900 python bench.py --config 3 --no-cpu-baseline --steps 3 > $O/c4_bench_cfg3.json 2> $O/c4_bench_cfg3.err; echo "cfg3 rc=$?"; tail -3 $O/c4_bench_cfg3.err
timeout 900 python bench.py --config 4 --no-cpu-baseline --steps 3 > $O/c4_bench_cfg4.json 2> $O/c4_bench_cfg4.err; echo "cfg4 rc=$?"; tail -3 $O/c4_bench_cfg4.err
CMD="python bench.py --wells 24 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-f64"
timeout 300 $CMD > $O/c4_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c4_launches.csv $CMD > $O/c4_ncu1.log 2>&1
python scripts/summarize_launches.py $O/c4_launches.csv $O/c4_launches_own > $O/c4_sum.log 2>&1; rm -f $O/c4_launches.csv
timeout 900 ncu --set full --clock-control none -k regex:"xdft_tc" -s 8 -c 4 -o /tmp/c4_prof $CMD > $O/c4_ncu2.log 2>&1
ncu -i /tmp/c4_prof.ncu-rep --page raw --csv > $O/c4_prof_raw.csv 2>/dev/null
du -sh $O
