#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python scratch/umma_rate.py > gpurun_out/c8_umma_rate.log 2>&1; echo "rc=$?"; cat gpurun_out/c8_umma_rate.log
