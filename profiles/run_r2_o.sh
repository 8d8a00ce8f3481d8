#!/bin/bash
# round 2, GPU call O: Lanczos product, arithmetic squares + blank list (form 3) against the table form (2), 512 / 384 threads
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for f in 3 2; do for t in 512 384; do echo "== lanczos form $f threads $t"; FPT_LANCZOS_FORM=$f FPT_LANCZOS_THREADS=$t timeout 300 python profiles/probe_large_cohort.py 296 2>&1 | tail -n 6; done; done > gpurun_out/r2o_probe_large.log 2>&1
grep -E "==|css_mds_large|lanczos phase" gpurun_out/r2o_probe_large.log | cut -c1-420
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "large or cohort or forms" > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/r2o_pytest.log
