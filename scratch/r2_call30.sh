#!/bin/bash
# Round 2, GPU call 30: one region split over several workers inside StitcherProcess.run (band mode), field_c0 in the ABI.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_stitcher_process_gpu.py -m gpu -q -x -k "split or splits" > $O/c30_pytest_split.log 2>&1; echo "split rc=$?"; tail -30 $O/c30_pytest_split.log
timeout 1200 python -m pytest tests -m gpu -q > $O/c30_pytest_all.log 2>&1; echo "all rc=$?"; tail -5 $O/c30_pytest_all.log
