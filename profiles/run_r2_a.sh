#!/bin/bash
# round 2, GPU call A: full GPU test-suite, peak microbenchmarks, large-cohort phase shares, sanitizers on the risky kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r2a_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
./profiles/microbench/peaks > gpurun_out/r2_peaks.json 2> gpurun_out/r2_peaks.err; cat gpurun_out/r2_peaks.json
timeout 300 python profiles/probe_large_cohort.py 296 > gpurun_out/r2a_probe_large.log 2>&1; tail -6 gpurun_out/r2a_probe_large.log
# compute-sanitizer over every kernel family (profiles/sanitizer_cases.py): hand-rolled mbarrier / tensor-memory protocols,
# shared-memory regions aliased between phases, chunked uploads on two streams
for tool in memcheck racecheck synccheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 30 python profiles/sanitizer_cases.py > gpurun_out/r2a_sanitizer_$tool.log 2>&1
  echo "rc=$?" >> gpurun_out/r2a_sanitizer_$tool.log
  tail -4 gpurun_out/r2a_sanitizer_$tool.log
done
