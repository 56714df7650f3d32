#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
SB_LIB_PATH=image_stitcher_b200/_lib/variants/libstitchb200_tcprof.so timeout 300 python scratch/tc_profile.py > $O/c5_tcprof.log 2>&1; echo "prof rc=$?"; cat $O/c5_tcprof.log | tail -60
timeout 600 python -m pytest tests/test_reg_tc_gpu.py tests/test_stitcher_process_gpu.py -m gpu -q > $O/c5_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/c5_pytest.log
timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-f64 > $O/c5_bench.json 2> $O/c5_bench.err; echo "bench rc=$?"; tail -3 $O/c5_bench.err
SB_REG_NO_TC_UPDFT=1 timeout 300 python bench.py --no-e2e --no-cpu-baseline --no-f64 > $O/c5_bench_noupdft.json 2> $O/c5_bench_noupdft.err
