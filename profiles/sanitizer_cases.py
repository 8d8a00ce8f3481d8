"""Workload for compute-sanitizer (memcheck / racecheck / synccheck): every kernel family once, small enough to finish under
the tools' slowdown. No oracle, no assertions on values beyond "scored something": the tool's report is the result.

    compute-sanitizer --tool racecheck python profiles/sanitizer_cases.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fpt_b200.api as api          # noqa: E402
import fpt_b200.synth as synth      # noqa: E402

which = sys.argv[1:] or ["small", "chunked", "smacof", "fallbacks", "large"]
if "small" in which:                 # m = 40: tridiag + eigvec + perm2 (u8 IMMA surrogate), FET count/score/window
    ch = synth.chromosome(7, 60000, 1500, 20, 20)
    s, d, w = api.fet_scan(ch["acodes"], ch["bcodes"], ch["pos"], 20, 20, 60000, 2500, 500, 0.95, seed=1)
    c, p, w2 = api.css_scan(ch["acodes"], ch["bcodes"], ch["pos"], 20, 20, 60000, 2500, 500, 1000, 1000, mds=0, seed=1)
    api.set_perm_mode(True)
    c, p, w2 = api.css_scan(ch["acodes"], ch["bcodes"], ch["pos"], 20, 20, 60000, 2500, 500, 20, 300, mds=0, seed=1)
    api.set_perm_mode(False)
    T = synth.coverage_tables(3, 20000, 20, 500)
    api.fet_tables(T)
    print("small ok", int(w.sum()), int(w2.sum()))
if "chunked" in which:               # upload cut into 8 chunks, windows released as their SNPs arrive
    os.environ["FPT_UPLOAD_CHUNK_BYTES"] = "4096"
    ch = synth.chromosome(321, 300000, 10000, 9, 8)
    av, bv, apos, bpos = synth.reference_layout(ch)
    s, d, w = api.fet_scan(av, bv, ch["pos"], 9, 8, 300000, 2500, 500, 0.95, seed=77)
    c, p, w2 = api.css_scan(ch["acodes"], ch["bcodes"], ch["pos"], 9, 8, 300000, 2500, 500, 5, 60, mds=0, seed=77)
    del os.environ["FPT_UPLOAD_CHUNK_BYTES"]
    print("chunked ok", int(w.sum()), int(w2.sum()))
if "smacof" in which:
    ch = synth.chromosome(41, 30000, 750, 20, 20)
    for mds in (1, 2):
        c, p, w = api.css_scan(ch["acodes"], ch["bcodes"], ch["pos"], 20, 20, 30000, 2500, 500, 10, 100, mds=mds, seed=5)
    print("smacof ok", int(w.sum()))
if "fallbacks" in which:             # 72: one warp per CTA; 130: gather surrogate; 290: Lanczos + general permutation kernel
    for a, b in ((36, 36), (70, 60), (150, 140)):
        ch = synth.chromosome(100 + a, 12000, 500, a, b)
        c, p, w = api.css_scan(ch["acodes"], ch["bcodes"], ch["pos"], a, b, 12000, 3000, 1500, 5, 60, mds=0, seed=3)
    api.set_perm_large_kernel(0)
    ch = synth.chromosome(300, 100000, 300, 150, 150)
    c, p, w = api.css_scan(ch["acodes"], ch["bcodes"], ch["pos"], 150, 150, 100000, 50000, 50000, 50, 200, mds=0, seed=3)
    api.set_perm_large_kernel(1)
    print("fallbacks ok", int(w.sum()))
if "large" in which:                 # 500+500: Lanczos, observed scores, tcgen05 permutation kernel over three batches of 128
    ch = synth.chromosome(500, 100000, 340, 500, 500)
    c, p, w = api.css_scan(ch["acodes"], ch["bcodes"], ch["pos"], 500, 500, 100000, 50000, 50000, 300, 300, mds=0, seed=4)
    ch = synth.chromosome(501, 100000, 340, 130, 171)       # m = 301: K / N padding, unequal groups
    c, p, w = api.css_scan(ch["acodes"], ch["bcodes"], ch["pos"], 130, 171, 100000, 50000, 50000, 300, 300, mds=0, seed=4)
    print("large ok", int(w.sum()))
