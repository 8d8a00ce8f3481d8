#!/usr/bin/env python
"""Opcode histogram (warp instructions executed, stall samples) of the SASS page of an .ncu-rep: python profiles/ncu_sass_hist.py rep [n]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS, iE, iN, iT = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
ex, sm = collections.Counter(), collections.Counter()
tot = tots = 0
for r in rows[2:]:
    if len(r) <= iE: continue
    ops = r[iS].split()
    op = ops[1] if ops and ops[0].startswith("@") else (ops[0] if ops else "?")
    op = op.split(".")[0]
    e, s = int(r[iE] or 0), int(r[iN] or 0)
    ex[op] += e; sm[op] += s; tot += e; tots += s
print("total warp instructions", tot, "samples", tots, "static instructions", len(rows) - 2)
for op, e in ex.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print("%-10s %12d %5.1f%%   samples %5.1f%%" % (op, e, 100.0 * e / tot, 100.0 * sm[op] / max(tots, 1)))
