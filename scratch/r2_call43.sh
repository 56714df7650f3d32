#!/bin/bash
# Round 2, GPU call 43: e2e with uploads restricted to the pixels that can reach the canvas (paste mode) vs whole tiles.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_register_gpu.py tests/test_fuse_gpu.py -m gpu -q -x > $O/c43_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/c43_pytest.log
B="python bench.py --no-cpu-baseline --no-f64 --e2e-steps 3"
run() {
  timeout 400 $B > $O/c43_bench_$1.json 2> $O/c43_bench_$1.err; rc=$?
  python - $1 $rc <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/c43_bench_{v}.json")); e=d["e2e"]; print(v, "rc", sys.argv[2], "e2e", round(e["value"]), "ms", round(e["ms_per_step"],1), "h2d GB", round(e["h2d_bytes_per_step"]/1e9,2), "d2h GB", round(e["d2h_bytes_per_step"]/1e9,2), "ceil", round(e["pcie_ceiling"]["gb_per_s_per_direction_per_gpu"],1), "frac", round(e["pcie_ceiling"]["e2e_frac_of_ceiling"],3), e["matches_device_result"], e["registration_truth_wells_ok"])
except Exception as ex: print(v, "rc", sys.argv[2], "ERR", ex)
PY
}
run partial
SB_E2E_FULL_UPLOAD=1 run full
run partial2
