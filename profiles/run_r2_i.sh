#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --chromosomes 1 --steps 1 --warmup 1 --skip-cpu --skip-fet --skip-large"
$B > /dev/null 2>&1 || { echo "bench failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:perm3 -c 1 -f -o gpurun_out/r2i_perm3 $B > gpurun_out/r2i_ncu_perm3.log 2>&1
ls -la gpurun_out/r2i_perm3.ncu-rep
