#!/bin/bash
# Round 2, GPU call 20 (2 GPUs): weak + strong scaling of the plate, the sharded mosaic, the 2-GPU context test.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi topo -m > $O/c20_topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
SECONDS=0
timeout 900 $TR bench.py --gpus 2 --no-cpu-baseline --no-f64 > $O/c20_weak_n2.json 2> $O/c20_weak_n2.err; echo "weak rc=$? ${SECONDS}s"; tail -2 $O/c20_weak_n2.err
timeout 600 $TR bench.py --gpus 2 --scaling strong --no-e2e --no-cpu-baseline --no-f64 > $O/c20_strong_n2.json 2> $O/c20_strong_n2.err; echo "strong rc=$? ${SECONDS}s"; tail -2 $O/c20_strong_n2.err
timeout 600 $TR bench.py --gpus 2 --config 4 --scaling strong --steps 3 > $O/c20_mosaic_n2.json 2> $O/c20_mosaic_n2.err; echo "mosaic rc=$? ${SECONDS}s"; tail -2 $O/c20_mosaic_n2.err
timeout 600 python -m pytest tests -m gpu -q -k "two or multi or device" > $O/c20_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/c20_pytest.log
python - <<'PY'
import json
for n in ["weak_n2","strong_n2","mosaic_n2"]:
    try:
        d=json.load(open(f"gpurun_out/c20_{n}.json"))
        print(n, "value", round(d["value"]), "ms", round(d["ms_per_step"],2), "reg", round(d["registration_ms_per_step"],2), "fuse", round(d["fusion_ms_per_step"],2),
              "e2e", d.get("e2e") and (round(d["e2e"]["value"]), round(d["e2e"]["pcie_ceiling"]["gb_per_s_per_direction_per_gpu"],1), round(d["e2e"]["pcie_ceiling"]["e2e_frac_of_ceiling"],2)),
              d.get("strong_scaling_check"), d.get("plane0_equals_single_call_canvas"), d.get("registration_truth_all_pairs_ok"), d.get("cpu_affinity"))
    except Exception as e: print(n, "failed", e)
PY
