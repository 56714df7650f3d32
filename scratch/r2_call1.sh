#!/bin/bash
# Round 2, GPU call 1: full GPU test suite, bench (batched fusion), fusion variants, per-well comparison,
# ncu launch list + --set full captures of the shipped registration configuration and the batched paste kernel.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > $O/c1_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $O/c1_pytest.log 2>&1; echo "pytest rc=$?" >> $O/c1_pytest.log
tail -5 $O/c1_pytest.log
timeout 600 python bench.py > $O/c1_bench.json 2> $O/c1_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --per-well-fusion --no-e2e --no-cpu-baseline > $O/c1_bench_perwell.json 2> $O/c1_bench_perwell.err
for v in keep1 keep2 rows4; do
  SB_LIB_PATH=image_stitcher_b200/_lib/variants/libstitchb200_$v.so timeout 300 python bench.py --no-e2e --no-cpu-baseline > $O/c1_bench_$v.json 2> $O/c1_bench_$v.err
done
timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-flatfield > $O/c1_bench_noflat.json 2> $O/c1_bench_noflat.err
# launch list (24 wells: four 54-pair sub-batches per registration, like the full plate)
CMD="python bench.py --wells 24 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 300 $CMD > $O/c1_plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/c1_launches.csv $CMD > $O/c1_ncu1.log 2>&1
timeout 1200 ncu --set full --clock-control none -k regex:"rows_fwd|cols_xpower|rows_inv|updft_rows|paste_rect" -s 50 -c 26 -o /tmp/c1_prof $CMD > $O/c1_ncu2.log 2>&1
ncu -i /tmp/c1_prof.ncu-rep --page raw --csv > $O/c1_prof_raw.csv 2>/dev/null
python scripts/summarize_launches.py $O/c1_launches.csv $O/c1_launches_own > $O/c1_sum.log 2>&1; rm -f $O/c1_launches.csv
rm -f $O/c1_launches.csv.bak
du -sh $O
ls -la $O | tail -20
