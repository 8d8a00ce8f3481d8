#!/bin/bash
# round 2, last call (under two GPU-minutes left): the pass rounding of the large-cohort code route checked bit for bit against
# single-pass window ranges on a whole 2600-window chromosome, its kernel times, then bench.py --small in the same process
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 95 python profiles/check_passes.py 2600 > gpurun_out/r2z_check_passes.log 2> gpurun_out/r2z_check_passes.err
echo "rc=$?"; tail -12 gpurun_out/r2z_check_passes.log
