#!/bin/bash
# round 2: the strong-scaled bench on N GPUs of one box under torchrun (usage: run_r2_multi.sh N), and the sharded-equals-single check
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_r2_${N}gpu.json 2> gpurun_out/bench_r2_${N}gpu.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_r2_${N}gpu.err
python - $N <<'PY'
import json, sys
N = sys.argv[1]
try:
    d = json.load(open("gpurun_out/bench_r2_%sgpu.json" % N))
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "scaling")}, "e2e", d["e2e"]["value"], d.get("replicas"))
    print("fet", d["fet"]["value"], d["fet"]["e2e"]["value"], d["fet"]["e2e_int8"]["value"])
    print("tables", d["fet_tables"]["value"], d["fet_tables"]["ms_per_step"], d["fet_tables"]["e2e"]["value"])
    print("large", d["large_cohort"]["value"], d["large_cohort"]["ms_per_step"], d["large_cohort"]["e2e"]["value"])
except Exception as e:
    print("parse failed", e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tests/multi_gpu_check.py 2>&1 | tail -3
