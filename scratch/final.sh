#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo bench rc=$?
python bench.py --impl reference > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err; echo ref rc=$?
