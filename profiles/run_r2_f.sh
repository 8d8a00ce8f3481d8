#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python profiles/probe_large_cohort.py 296 2>&1 | tail -n 5
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "large or cohort or gemm or tensor or forms" > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/r2f_pytest.log
