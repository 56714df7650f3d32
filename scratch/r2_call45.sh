#!/bin/bash
# Round 2, GPU call 35 (4 GPUs): weak-scaling bench line with the e2e leg and the node's concurrent PCIe ceiling.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi topo -m > $O/c45_topo_8gpu.txt 2>&1
nproc > $O/c45_nproc.txt; free -g | head -2 >> $O/c45_nproc.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 3 --warmup 3 --no-cpu-baseline --no-f64 > $O/c45_bench_n4.json 2> $O/c45_bench_n4.err; echo "n4 rc=$?"
tail -3 $O/c45_bench_n4.err
python -c "
import json; d=json.load(open('gpurun_out/c45_bench_n4.json')); print('n4 value', round(d['value']), 'step', round(d['ms_per_step'],2), 'e2e', d['e2e']['value'], d['e2e'].get('pcie_ceiling'))"
