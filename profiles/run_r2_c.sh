#!/bin/bash
# round 2, GPU call C (2 GPUs): the strong-scaled bench under torchrun, and the sharded-equals-single check
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2c_bench_2gpu.json 2> gpurun_out/r2c_bench_2gpu.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r2c_bench_2gpu.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/r2c_bench_2gpu.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "scaling")}, d["e2e"]["value"], d.get("replicas"))
    print("fet", d["fet"]["value"], d["fet"]["e2e"]["value"], d["fet"]["e2e_int8"]["value"])
    print("tables", d["fet_tables"]["value"], d["fet_tables"]["e2e"]["value"])
    print("large", d["large_cohort"]["value"], d["large_cohort"]["e2e"]["value"])
except Exception as e:
    print("parse failed", e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/multi_gpu_check.py 2>&1 | tail -3
