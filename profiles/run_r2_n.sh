#!/bin/bash
# round 2, GPU call N: Lanczos product with software-pipelined loads (512 / 384 threads), then the whole GPU suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for t in 512 384; do echo "== lanczos threads $t"; FPT_LANCZOS_THREADS=$t timeout 300 python profiles/probe_large_cohort.py 296 2>&1 | tail -n 6; done > gpurun_out/r2n_probe_large.log 2>&1
grep -E "==|css_mds_large|lanczos phase" gpurun_out/r2n_probe_large.log | cut -c1-400
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 8 gpurun_out/r2n_pytest.log
