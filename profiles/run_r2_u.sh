#!/bin/bash
# round 2, GPU call U: tcgen05 permutation kernel with full shuffle rounds running ahead of the batches and the four-sub-stream shuffle
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -k "large or cohort or tensor" > gpurun_out/r2u_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 5 gpurun_out/r2u_pytest.log
timeout 300 python profiles/probe_large_cohort.py 296 2>&1 | tail -n 6 > gpurun_out/r2u_probe_large.log; timeout 300 python profiles/probe_large_cohort.py 2600 2>&1 | tail -n 6 >> gpurun_out/r2u_probe_large.log
grep -E "windows|umma phase" gpurun_out/r2u_probe_large.log | cut -c1-420
