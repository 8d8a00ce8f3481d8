#!/bin/bash
# round 2, GPU call G: ncu source-level captures of the three dominant kernels (perm2 at m = 40; Lanczos-on-codes and the tcgen05 permutation kernel at m = 1000)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
B="python bench.py --chromosomes 1 --steps 1 --warmup 1 --skip-cpu --skip-fet --skip-large"
$B > /dev/null 2>&1 || { echo "bench failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:perm2 -c 1 -f -o gpurun_out/r2g_perm2 $B > gpurun_out/r2g_ncu_perm2.log 2>&1
python profiles/probe_large_cohort.py 296 > /dev/null 2>&1 || { echo "probe failed"; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:mds_codes -c 1 -f -o gpurun_out/r2g_mds_codes python profiles/probe_large_cohort.py 296 > gpurun_out/r2g_ncu_mds.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:perm_umma -c 1 -f -o gpurun_out/r2g_perm_umma python profiles/probe_large_cohort.py 296 > gpurun_out/r2g_ncu_umma.log 2>&1
ls -la gpurun_out/*.ncu-rep
