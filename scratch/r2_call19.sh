#!/bin/bash
# Round 2, GPU call 19: sharded-mosaic bench mode on one GPU (small, then BASELINE configs[4] at full size).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python bench.py --config 4 --scaling strong --grid 4 --tile 2048 --num-z 2 --steps 2 --warmup 1 > $O/c19_mosaic_small.json 2> $O/c19_mosaic_small.err; echo "small rc=$?"; tail -3 $O/c19_mosaic_small.err | head -2; cut -c1-1500 $O/c19_mosaic_small.json
SECONDS=0; timeout 900 python bench.py --config 4 --scaling strong --steps 3 > $O/c19_mosaic_n1.json 2> $O/c19_mosaic_n1.err; echo "full rc=$? in ${SECONDS}s"; grep -E "Error|error" $O/c19_mosaic_n1.err | head; cut -c1-2500 $O/c19_mosaic_n1.json
