#!/usr/bin/env python
"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled into oracle/_ref (oracle/Makefile).
Run in the build container (needs /root/reference); the fixtures travel, the reference does not.

    python tests/golden/make_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import checkers  # noqa: E402
from checkers import RefMatrix, dptr, iptr  # noqa: E402


def genotypes(rng, nsnp, size, p=(0.45, 0.3, 0.23, 0.02)):
    return rng.choice(np.array([3.0, -3.0, 0.0, -10000.0]), size=nsnp * size, p=p)


def fet_fixture():
    rf = checkers.load_ref_fet()
    rng = np.random.default_rng(20261018)
    out = {}
    # (1) 4000 random tables with N <= 67 and their two-tailed P from the reference's fet()
    T = []
    while len(T) < 4000:
        n = int(rng.integers(0, 68))
        cuts = np.sort(rng.integers(0, n + 1, size=3))
        T.append((cuts[0], cuts[1] - cuts[0], cuts[2] - cuts[1], n - cuts[2]))
    T = np.array(T, dtype=np.int32)
    tmp = (C.c_int * 4)()
    P = np.array([rf.fet((C.c_int * 4)(*[int(v) for v in t]), tmp) for t in T])
    out["tables"], out["tables_p"] = T, P
    # (2) genotype rows -> fetcount tables
    asize, bsize, nsnp = 13, 9, 64
    av, bv = genotypes(rng, nsnp, asize), genotypes(rng, nsnp, bsize, (0.3, 0.45, 0.23, 0.02))
    cnt = np.zeros((nsnp, 4), dtype=np.int32)
    for k in range(nsnp):
        f = (C.c_int * 4)()
        rf.fetcount(f, dptr(av), dptr(bv), k, asize, bsize)
        cnt[k] = list(f)
    out["count_av"], out["count_bv"], out["count_tables"] = av, bv, cnt
    # (3) one window end to end (percentile + bootstrap sigma) from explicit nrand48 states
    asize = bsize = 20
    wins = []
    for npos, state in ((37, 0x1234ABCD5678), (5, 42), (2, 7), (120, 0xFFFFFFFFFFFF)):
        a, b = genotypes(rng, npos, asize), genotypes(rng, npos, bsize, (0.3, 0.45, 0.23, 0.02))
        res = np.zeros(2)
        f, t = (C.c_int * 4)(), (C.c_int * 4)()
        samples, stds, fets = np.zeros(npos), np.zeros(100), np.zeros(npos)
        rf.fisher_exact_test(dptr(res), dptr(a), dptr(b), asize, bsize, npos, f, t, dptr(samples), dptr(stds), 100, dptr(fets),
                             checkers.state_to_ushort3(state), 0.95)
        wins.append((a, b, state, res.copy()))
    for i, (a, b, state, res) in enumerate(wins):
        out["win%d_av" % i], out["win%d_bv" % i], out["win%d_state" % i], out["win%d_res" % i] = a, b, np.uint64(state), res
    out["nwin"] = len(wins)
    # (4) a serial `compute` scan (scores are deterministic, stddev is seeded from time(NULL) and left out)
    regend, wsize, wstep, nsnp = 40000, 2500, 500, 700
    pos = np.sort(rng.choice(regend, size=nsnp, replace=False)).astype(np.int32)
    a, b = genotypes(rng, nsnp, asize), genotypes(rng, nsnp, bsize, (0.3, 0.45, 0.23, 0.02))
    apos, bpos = np.repeat(pos, asize).astype(np.int32), np.repeat(pos, bsize).astype(np.int32)
    s, d = np.zeros(regend // wstep), np.zeros(regend // wstep)
    with checkers.silence_stdout():
        rf.compute(dptr(a), dptr(b), iptr(apos), iptr(bpos), 0, regend, wsize, wstep, a.size, b.size, 0.95, dptr(s), dptr(d))
    out.update(scan_av=a, scan_bv=b, scan_pos=pos, scan_geom=np.array([regend, wsize, wstep, asize, bsize]), scan_scores=s)
    np.savez_compressed(os.path.join(HERE, "fet_golden.npz"), **out)


def css_fixture():
    rc = checkers.load_ref_css()
    rng = np.random.default_rng(20261019)
    out = {}
    asize, bsize, npos = 8, 7, 40
    m = asize + bsize
    f = np.clip(rng.beta(0.5, 0.5, npos), 0.05, 0.95)
    enc = np.array([3.0, 0.0, -3.0])
    av = enc[rng.binomial(2, f[:, None], size=(npos, asize))].ravel().copy()
    bv = enc[rng.binomial(2, (1 - f)[:, None], size=(npos, bsize))].ravel().copy()
    D = RefMatrix(m, m)
    rc.compare_all(dptr(av), dptr(bv), asize, bsize, npos, D.pp)
    out.update(av=av, bv=bv, shape=np.array([asize, bsize, npos]), compare_all=D.a.copy())
    keep = rc.fill_averages(D.pp, m)
    out.update(fill_keep=np.int32(keep), filled=D.a.copy())
    X, B, Z, T, L, Q = RefMatrix(m, 2), RefMatrix(m, m), RefMatrix(m, m), RefMatrix(m, m), RefMatrix(2, 2), RefMatrix(m, 2)
    rc.cmds(D.pp, X.pp, 2, m, B.pp, Z.pp, T.pp, L.pp, Q.pp)
    out["cmds_X"] = X.a.copy()
    dist = RefMatrix(m, m)
    rc.calc_dist(X.pp, dist.pp, m)
    at, bt = np.arange(asize, dtype=np.int32), np.arange(asize, m, dtype=np.int32)
    score = rc.css(dist.pp, iptr(at), iptr(bt), asize, bsize)
    out.update(cmds_dist=dist.a.copy(), cmds_score=np.float64(score))
    state = 0x0BADC0FFEE11
    tracks = np.arange(m, dtype=np.int32)
    p = rc.significance_treshold(dist.pp, iptr(tracks), asize, bsize, score, 10, 500, checkers.state_to_ushort3(state))
    out.update(perm_state=np.uint64(state), perm_p=np.float64(p), perm_tracks_after=tracks.copy())
    Xs, Zs, Bs, Ds = RefMatrix(m, 2, X.a.copy()), RefMatrix(m, 2), RefMatrix(m, m), RefMatrix(m, m)
    sigma = rc.smacof(D.pp, m, 2, Xs.pp, Zs.pp, Bs.pp, Ds.pp, 300, 1e-6)
    out.update(smacof_from_cmds_X=Xs.a.copy(), smacof_from_cmds_sigma=np.float64(sigma))
    checkers.seed48(state)
    Xr, Rr = RefMatrix(m, 2), RefMatrix(m, 2)
    rc.smacof_runs(D.pp, m, 2, Xr.pp, Zs.pp, Bs.pp, Ds.pp, Rr.pp, 300, 4, 1e-6)
    out["smacof_runs_X"] = Xr.a.copy()
    # a serial `compute` scan, classical MDS: the score column is deterministic
    asize = bsize = 10
    regend, wsize, wstep, nsnp = 30000, 2500, 500, 600
    pos = np.sort(rng.choice(regend, size=nsnp, replace=False)).astype(np.int32)
    f = np.clip(rng.beta(0.5, 0.5, nsnp), 0.05, 0.95)
    a = enc[rng.binomial(2, f[:, None], size=(nsnp, asize))].ravel().copy()
    b = enc[rng.binomial(2, np.where((pos // 5000) % 2 == 1, 1 - f, f)[:, None], size=(nsnp, bsize))].ravel().copy()
    apos, bpos = np.repeat(pos, asize).astype(np.int32), np.repeat(pos, bsize).astype(np.int32)
    s, pp = np.zeros(regend // wstep), np.zeros(regend // wstep)
    with checkers.silence_stdout():
        rc.compute(dptr(a), dptr(b), iptr(apos), iptr(bpos), 0, regend, wsize, wstep, a.size, b.size, 5, 50, 0, 0, dptr(s), dptr(pp))
    out.update(scan_av=a, scan_bv=b, scan_pos=pos, scan_geom=np.array([regend, wsize, wstep, asize, bsize]), scan_scores=s,
               scan_scored=(pp != 0))
    np.savez_compressed(os.path.join(HERE, "css_golden.npz"), **out)


def vcf_fixture():
    """small.vcf (synthetic, two chromosomes, phased and unphased calls, missing calls, extra FORMAT slots) and the GTrack
    text the reference's own converter (tools/VCFConvert.py, Python 2: its print statements are rewritten as calls, nothing
    else is touched) makes of it for two populations; one requested individual is absent from the header."""
    import re
    src = open("/root/reference/tools/VCFConvert.py").read()
    src = re.sub(r"^(\s*)print (.*?);?\s*$", r"\1print(\2)", src, flags=re.M)
    ns = {"__name__": "vcfconvert_reference"}
    exec(compile(src, "VCFConvert.py", "exec"), ns)
    rng = np.random.default_rng(11)
    names = ["ind%02d" % i for i in range(9)]
    gts = ["0/0", "0|0", "0/1", "1/0", "0|1", "1|0", "1/1", "1|1", "./.", ".|."]
    lines = ["##fileformat=VCFv4.1", "##source=fpt_b200 tests/golden/make_golden.py",
             "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(names)]
    for chrom, n in (("chrI", 40), ("chrII", 25)):
        for pos in np.sort(rng.choice(20000, size=n, replace=False)):
            calls = ["%d:%s:%d" % (rng.integers(1, 60), gts[rng.integers(0, len(gts))], rng.integers(0, 99)) for _ in names]
            lines.append("%s\t%d\t.\tA\tC\t50\tPASS\t.\tDP:GT:GQ\t%s" % (chrom, pos, "\t".join(calls)))
    vcf = "\n".join(lines) + "\n"
    open(os.path.join(HERE, "small.vcf"), "w").write(vcf)
    pops = {"A": ["ind00", "ind02", "ind04", "ind07", "nobody"], "B": ["ind08", "ind01", "ind03", "ind05"]}
    for tag, pop in pops.items():
        with checkers.silence_stdout():
            text = ns["addHeader"]("test") + ns["convertToGtrackFile"](vcf, list(pop), "test")
        open(os.path.join(HERE, "small_pop%s.gtrack" % tag), "w").write(text)


if __name__ == "__main__":
    if os.path.exists("/root/reference/tools/VCFConvert.py"):
        vcf_fixture()
    if not checkers.ref_available():
        raise SystemExit("oracle/_ref is not built: run `make -C oracle ref` where /root/reference exists")
    fet_fixture()
    css_fixture()
    for fn in ("fet_golden.npz", "css_golden.npz"):
        print(fn, os.path.getsize(os.path.join(HERE, fn)), "bytes")
