#!/bin/bash
# round 2, GPU call V: tile-sorting FET score kernel, block-wise Gram-Schmidt in the Lanczos kernel, host-side trace of the drop-in
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "fet or large or cohort or forms" > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/r2v_pytest.log
timeout 300 python profiles/probe_large_cohort.py 296 2>&1 | tail -n 6 > gpurun_out/r2v_probe_large.log; timeout 300 python profiles/probe_large_cohort.py 2600 2>&1 | tail -n 6 >> gpurun_out/r2v_probe_large.log
grep -E "windows|lanczos phase" gpurun_out/r2v_probe_large.log | cut -c1-420
FPT_TRACE=1 timeout 900 python bench.py --chromosomes 2 --steps 2 --warmup 1 --skip-cpu --skip-large > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?"
grep "css drop-in" gpurun_out/r2v_bench.err | tail -n 4
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2v_bench.json"))
ft = d["fet_tables"]; print("fet_tables", ft["value"], ft["ms_per_step"], ft["roofline"]["fp64"]["frac"])
print("css", d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e_pageable"])
print({k: v.get("e2e") for k, v in d["mds_variants"].items() if isinstance(v, dict)})
PY
