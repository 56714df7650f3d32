#!/bin/bash
for d in 0 1 2 3 4 6 7; do echo "debug=$d"; SB_FUSE_DEBUG=$d timeout 120 python scratch/perf_fuse2.py paste 2>&1 | tail -2; done
