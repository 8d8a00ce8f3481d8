#!/bin/bash
# round 2: both bench arms with the final build and the final bench.py (no captures)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/bench_r2_1gpu.json 2> gpurun_out/bench_r2_1gpu.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference > gpurun_out/bench_r2_reference_arm.json 2> gpurun_out/bench_r2_reference_arm.err; echo "reference arm rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2_1gpu.json"))
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, "e2e", d["e2e"]["value"], "pageable", d["e2e_pageable"]["value"], "cpu", d.get("cpu_baseline", {}).get("value"))
print("fet", d["fet"]["value"], d["fet"]["e2e"]["value"], d["fet"]["e2e_int8"]["value"], d["fet"]["e2e_pageable"]["value"], "tables", d["fet_tables"]["value"], d["fet_tables"]["cpu_baseline"]["parity_vs_gpu"])
print("large", d["large_cohort"]["value"], d["large_cohort"]["kernels"]["css_mds_large"].get("fp64"))
print({k: (v.get("e2e", {}).get("value"), v.get("cpu_baseline", {}).get("value"), v.get("cpu_baseline", {}).get("parity_vs_gpu")) for k, v in d["mds_variants"].items() if isinstance(v, dict)})
r = json.load(open("gpurun_out/bench_r2_reference_arm.json")); print("reference arm", r.get("value"))
PY
