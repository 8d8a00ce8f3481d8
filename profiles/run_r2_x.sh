#!/bin/bash
# round 2, GPU call X: drop-in host side (position gather, A/B check beside the scan, staged uploads of pageable arrays); trace
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "not large and not cohort and not tensor and not forms and not gemm" > gpurun_out/r2x_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r2x_pytest.log
FPT_TRACE=1 timeout 600 python bench.py --steps 3 --warmup 2 --skip-large --skip-cpu > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; echo "bench rc=$?"
grep "css drop-in" gpurun_out/r2x_bench.err | head -n 30 | tail -n 4; grep "fet drop-in" gpurun_out/r2x_bench.err | tail -n 2
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2x_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["e2e_pageable"]["value"])
print("fet", d["fet"]["value"], d["fet"]["e2e"]["value"], d["fet"]["e2e_int8"]["value"], d["fet"]["e2e_pageable"]["value"])
PY
