#!/bin/bash
# round 2, GPU call D: the genotype GEMM route (tcgen05) vs popcounts vs the round-1 route: parity tests and per-kernel times
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "large or cohort or gemm or tensor or forms" > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 15 gpurun_out/r2d_pytest.log
for mode in 2 1 0; do
  echo "== k4 mode $mode"; FPT_K4_MODE=$mode timeout 300 python profiles/probe_large_cohort.py 296 2>&1 | tail -n 5
done | tee gpurun_out/r2d_probe_large.log
