#!/bin/bash
# round 2, final captures: GPU suite, smoke, both bench arms, launch lists, one ncu --set full capture per kernel (each after its
# command has run once without ncu)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2_final_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/r2_final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1500 python bench.py > gpurun_out/bench_r2_1gpu.json 2> gpurun_out/bench_r2_1gpu.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference > gpurun_out/bench_r2_reference_arm.json 2> gpurun_out/bench_r2_reference_arm.err; echo "reference arm rc=$?"
B="python bench.py --chromosomes 1 --steps 1 --warmup 1 --skip-cpu --skip-fet --skip-large"
L="python profiles/probe_large_cohort.py 296"
F="python profiles/probe_fet_tables.py 4000000"
W="python bench.py --chromosomes 1 --steps 1 --warmup 1 --skip-cpu --skip-large"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --chromosomes 2 --steps 2 --warmup 1 --skip-cpu --skip-large > gpurun_out/ncu_launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_r2_large.csv $L > gpurun_out/ncu_launches_large.log 2>&1
cap() { ncu --set full --clock-control none --import-source on -k regex:$1 -c 1 -f -o gpurun_out/prof_$2 ${@:3} > gpurun_out/ncu_$2.log 2>&1; }
cap perm3 css_perm3 $B
cap tridiag_reg css_tridiag_reg $B
cap eigvec css_eigvec $B
cap observed css_observed $B
cap fet_window fet_window $W
cap fet_count fet_count $W
cap fet_score fet_score_tables $F
cap mds_codes css_mds_codes $L
cap perm_umma css_perm_umma $L
cap k4_umma css_k4_umma $L
ls -la gpurun_out/prof_*.ncu-rep | awk '{print $5, $9}'
# gpurun brings back at most 64 MiB: reduce the reports to their summaries here (raw metric tables, SASS opcode histograms) and keep two
mkdir -p gpurun_out/summ
FPT_SUMM_OUT=gpurun_out/summ python profiles/summarise.py r2 > gpurun_out/summ/summarise.log 2>&1
for f in gpurun_out/prof_*.ncu-rep; do k=$(basename $f .ncu-rep); k=${k#prof_}; python profiles/ncu_sass_hist.py $f 30 > gpurun_out/summ/r2_${k}_sass_hist.txt 2>&1; done
for f in gpurun_out/prof_*.ncu-rep; do case $f in *css_perm3*|*css_mds_codes*) ;; *) rm -f $f;; esac; done
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2_1gpu.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, "e2e", d["e2e"]["value"], "cpu", d.get("cpu_baseline", {}).get("value"))
print({k: round(v["ms_per_launch"], 3) for k, v in d["kernels"].items()})
print("fet", d["fet"]["value"], d["fet"]["e2e"]["value"], "tables", d["fet_tables"]["value"], "large", d["large_cohort"]["value"], d["large_cohort"].get("cpu_baseline"))
r = json.load(open("gpurun_out/bench_r2_reference_arm.json")); print("reference arm", r.get("value"), r.get("cpu_baseline"))
PY
