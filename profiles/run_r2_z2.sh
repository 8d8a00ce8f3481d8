#!/bin/bash
# round 2, after the pass rounding of the large-cohort route: the bench's large-cohort leg (8 chromosomes x 2600 windows) with the
# per-chromosome work models, beside a two-chromosome headline (profiling shape, not the headline figure); no CPU legs
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 50 python bench.py --skip-fet --skip-cpu --chromosomes 2 --steps 2 --warmup 3 > gpurun_out/r2z2_bench_large.json 2> gpurun_out/r2z2_bench_large.err
echo "rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2z2_bench_large.json"))
lc = d["large_cohort"]
print("large", lc["value"], lc["ms_per_step"], "e2e", lc["e2e"]["value"], lc["gpu_launches"], lc["kernel_ms_per_chromosome"])
m = lc["kernels"]["css_mds_large"]
print("mds", m["scopes_per_chromosome"], m.get("fp64", {}).get("frac"), m.get("traffic_over_algorithmic"), "perm", lc["kernels"]["css_perm"].get("frac"))
print("headline (2 chromosomes)", d["value"], d["ms_per_step"], d["kernels"]["css_perm"].get("issue", {}).get("frac"))
PY
