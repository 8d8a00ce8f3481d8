#!/bin/bash
# final round-1 captures: run each command without ncu first, then under ncu
set -x
B="python bench.py --chromosomes 2 --steps 1 --warmup 1 --skip-cpu --skip-fet --skip-large"
$B > /dev/null 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --chromosomes 2 --steps 2 --warmup 1 --skip-cpu --skip-large > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:perm2 -c 1 -f -o gpurun_out/prof_css_perm2 $B > gpurun_out/ncu_perm2.log 2>&1
python profiles/probe_large_cohort.py 296 > /dev/null 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:mds_large -c 1 -f -o gpurun_out/prof_css_mds_large python profiles/probe_large_cohort.py 296 > gpurun_out/ncu_mdsl.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:perm_umma -c 1 -f -o gpurun_out/prof_css_perm_umma python profiles/probe_large_cohort.py 296 > gpurun_out/ncu_pl.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:observed -c 1 -f -o gpurun_out/prof_css_observed python profiles/probe_large_cohort.py 296 > gpurun_out/ncu_obs.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r1_large.csv python profiles/probe_large_cohort.py 296 > gpurun_out/ncu_launches_large.log 2>&1
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --impl reference > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err
ls -la gpurun_out | tail -12
