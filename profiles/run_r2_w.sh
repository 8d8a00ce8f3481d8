#!/bin/bash
# round 2, GPU call W: tridiagonalisation hand-over through shared memory, upload in three growing pieces
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "css and not large and not cohort and not gemm and not tensor and not forms or rationals or cohort_beyond or larger_cohorts" > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 8 gpurun_out/r2w_pytest.log
timeout 600 python bench.py --steps 3 --warmup 2 --skip-fet --skip-large --skip-cpu > gpurun_out/r2w_bench.json 2> gpurun_out/r2w_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2w_bench.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["perm_rechecks"]["exact_rescorings_rank0"])
print({k: round(v["ms_per_launch"], 3) for k, v in d["kernels"].items()})
PY
B="python bench.py --chromosomes 1 --steps 1 --warmup 1 --skip-cpu --skip-fet --skip-large"
