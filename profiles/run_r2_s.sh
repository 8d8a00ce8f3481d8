#!/bin/bash
# round 2, GPU call S: FET log-mode walk with Newton reciprocals; code-route Lanczos split into arithmetic / general kernels (256 threads)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "fet or large or cohort or forms" > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/r2s_pytest.log
timeout 900 python bench.py --chromosomes 1 --steps 3 --warmup 2 --skip-cpu > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2s_bench.json"))
ft = d["fet_tables"]; print("fet_tables", ft["value"], ft["ms_per_step"], ft["e2e"]["value"], ft["roofline"]["fp64"]["frac"])
print("fet", d["fet"]["value"], d["fet"]["e2e"]["value"])
lc = d["large_cohort"]; print("large", lc["value"], lc["ms_per_step"], lc["e2e"]["value"], {k: round(v["ms_per_launch"], 2) for k, v in lc["kernels"].items()})
PY
