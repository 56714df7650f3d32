#!/bin/bash
# Round 2, GPU call 50: shipped paste kernel = 2 groups per chunk, 40 registers, 6 blocks per SM: tests, bench lines, ncu.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_fuse_gpu.py tests/test_configs_gpu.py tests/test_stitcher_process_gpu.py tests/test_pyramid_gpu.py -m gpu -q -x > $O/c50_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/c50_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/c50_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py > $O/c50_bench.json 2> $O/c50_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --config 3 --steps 3 --no-cpu-baseline --no-f64 > $O/c50_bench_cfg3.json 2> $O/c50_bench_cfg3.err; echo "cfg3 rc=$?"
timeout 600 python bench.py --config 1 --steps 3 --no-cpu-baseline --no-f64 > $O/c50_bench_cfg1.json 2> $O/c50_bench_cfg1.err; echo "cfg1 rc=$?"
CMD="python bench.py --wells 24 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-f64"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"paste_rect" -s 4 -c 1 -o $O/c50_paste $CMD > $O/c50_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i $O/c50_paste.ncu-rep --page raw --csv > $O/c50_paste_raw.csv 2>/dev/null
python - <<'PY'
import json
for n in ["bench","bench_cfg3","bench_cfg1"]:
    try:
        d=json.load(open(f"gpurun_out/c50_{n}.json"))
        print(n,"value",round(d["value"]),"ms",round(d["ms_per_step"],3),"reg",round(d["registration_ms_per_step"],3),"fuse",round(d["fusion_ms_per_step"],3),"frac",round(d["roofline"]["frac"],3),"coord", d.get("fusion_coordinate_only") and round(d["fusion_coordinate_only"]["frac"],3), d.get("registration_truth_wells_ok"), "e2e", d.get("e2e") and (round(d["e2e"]["value"]), d["e2e"]["matches_device_result"]))
    except Exception as e: print(n,"failed",e)
PY
