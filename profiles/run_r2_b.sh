#!/bin/bash
# round 2, GPU call B: new bench.py at N=1 (strong-scaling code path, per-kernel rooflines), peaks with the fixed smem microbenchmark
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
./profiles/microbench/peaks > gpurun_out/r2_peaks.json 2> gpurun_out/r2_peaks.err; cat gpurun_out/r2_peaks.json
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2b_bench_1gpu.json 2> gpurun_out/r2b_bench_1gpu.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2b_bench_1gpu.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2b_bench_1gpu.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus")}, d["e2e"]["value"], d.get("cpu_baseline", {}).get("value"))
    print({k: (round(v["ms_per_launch"], 3), v.get("frac")) for k, v in d["kernels"].items()})
    print("fet", d["fet"]["value"], d["fet"]["e2e"]["value"], d["fet"]["e2e_int8"]["value"], d["fet"]["e2e_pageable"])
    print("tables", d["fet_tables"]["value"], d["fet_tables"]["roofline"], d["fet_tables"]["work_model"])
    print("large", d["large_cohort"]["value"], d["large_cohort"]["e2e"]["value"], d["large_cohort"]["kernel_ms_per_chromosome"], d["large_cohort"].get("cpu_baseline"))
    print("variants", d["mds_variants"])
except Exception as e:
    print("parse failed", e)
PY
