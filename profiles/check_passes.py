"""Large-cohort code route: one scan of a whole configs[4] chromosome (several GEMM -> Lanczos passes inside the library) must
equal, bit for bit, the same chromosome scanned as separate window ranges that each fit one pass (random streams are keyed by
the global window index). Prints the per-kernel device time of the whole-chromosome scan. Then, in the same process (one torch
import), the bench's own smoke run: `bench.py --small --skip-cpu`.
usage: python profiles/check_passes.py [windows=2600] > gpurun_out/check_passes.log"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
T0 = time.time()


def log(*a):
    print("[%6.1f s]" % (time.time() - T0), *a, flush=True)


import numpy as np
import torch
import bench
import fpt_b200.api as api
from fpt_b200 import _lib

lib = _lib.load()
log("imports done")
nwin = int(sys.argv[1]) if len(sys.argv) > 1 else 2600
dev = torch.device("cuda", 0)
L = bench.LARGE
dpos, da, db, regend, nsnp = bench.large_chromosome(torch, dev, nwin)
pos, a, b = (x.cpu().numpy() for x in (dpos, da, db))
del dpos, da, db
log("chromosome of %d windows, %d SNPs generated" % (nwin, nsnp))


def scan(w0, w1):
    return api.css_scan(a, b, pos, L["asize"], L["bsize"], regend, L["wsize"], L["wstep"], L["mcr"], L["mcr"], mds=0, seed=7,
                        window_begin=w0, window_end=w1)


scan(0, 8)                                               # warm-up: workspace, tables
lib.fpt_profile_enable(1)
bench.profile_json(lib)
t = time.time()
s, p, w = scan(0, nwin)
dt = time.time() - t
prof = bench.profile_json(lib)
lib.fpt_profile_enable(0)
log("whole chromosome: %.3f s end to end from pageable arrays, scored %d of %d" % (dt, int((w != 0).sum()), nwin))
log("kernel ms per chromosome:", json.dumps({k: round(v["ms"], 3) for k, v in prof.items()}),
    "scopes:", json.dumps({k: v["launches"] for k, v in prof.items()}))
cuts = list(range(0, nwin, 700)) + [nwin]                # ranges of <= 700 windows: one pass each
ok = True
for w0, w1 in zip(cuts, cuts[1:]):
    s1, p1, w1_ = scan(w0, w1)
    same = (np.array_equal(s[w0:w1], s1, equal_nan=True) and np.array_equal(p[w0:w1], p1, equal_nan=True)
            and np.array_equal(w[w0:w1], w1_))
    ok &= bool(same)
    log("windows [%d, %d): %s" % (w0, w1, "identical" if same else "DIFFERENT"))
log("PASSES", "OK" if ok else "MISMATCH", "| finite scores", int(np.isfinite(s).sum()), "| p < 0.05:", int((p < 0.05).sum()),
    "| score checksum %.17g" % float(np.nansum(s)))
del a, b, pos
torch.cuda.empty_cache()

if os.environ.get("FPT_SKIP_SMALL_BENCH") != "1":
    sys.argv = ["bench.py", "--small", "--skip-cpu", "--steps", "1", "--warmup", "1"]
    log("bench.py --small --skip-cpu")
    bench.main()
    log("bench main returned")
sys.exit(0 if ok else 1)
