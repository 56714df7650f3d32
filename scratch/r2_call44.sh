#!/bin/bash
# Round 2, GPU call 44: e2e with uploads restricted to the ROWS that can reach the canvas (contiguous copies).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
B="python bench.py --no-cpu-baseline --no-f64 --e2e-steps 3 --no-coordinate-only"
run() {
  timeout 400 $B $2 > $O/c44_bench_$1.json 2> $O/c44_bench_$1.err; rc=$?
  python - $1 $rc <<'PY'
import json,sys
v=sys.argv[1]
try:
    d=json.load(open(f"gpurun_out/c44_bench_{v}.json")); e=d["e2e"]; print(v, "rc", sys.argv[2], "e2e", round(e["value"]), "ms", round(e["ms_per_step"],1), "h2d GB", round(e["h2d_bytes_per_step"]/1e9,2), "d2h GB", round(e["d2h_bytes_per_step"]/1e9,2), "ceil", round(e["pcie_ceiling"]["gb_per_s_per_direction_per_gpu"],1), "frac", round(e["pcie_ceiling"]["e2e_frac_of_ceiling"],3), e["matches_device_result"], e["registration_truth_wells_ok"])
except Exception as ex: print(v, "rc", sys.argv[2], "ERR", ex)
PY
}
run rows "--e2e-partial-upload rows"
run off ""
run rows2 "--e2e-partial-upload rows"
