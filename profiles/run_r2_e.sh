#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python profiles/debug_k4.py 2>&1 | tail -n 5
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "gemm" > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 25 gpurun_out/r2e_pytest.log
