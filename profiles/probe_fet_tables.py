"""configs[3] tables (row sums U{20..500}, log-space walk) through fpt_fet_tables: the launch ncu captures for fet_score in log mode"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fpt_b200.api as api, fpt_b200.synth as synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4_000_000
T = synth.coverage_tables(3, n, 20, 500)
for _ in range(2):
    t = time.time(); s = api.fet_tables(T); dt = time.time() - t
print("tables", n, "sec", round(dt, 3), "mean -log10 P", float(s.mean()))
