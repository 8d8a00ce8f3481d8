"""Synthetic genotype data of SURVEY.md section 8(d): sorted unique positions, ancestral minor-allele
frequency f ~ Beta(0.5, 0.5) clipped to [0.02, 0.98], population B flipped to 1-f inside planted divergent
segments, diploid genotype ~ Binomial(2, f), code 3 / 0 / -3 for 0 / 1 / 2 minor alleles, 2 % missing (-10000
in the reference layout, -128 in the compact int8 layout)."""
import numpy as np

MISSING_F64 = -10000.0
MISSING_I8 = -128


def chromosome(seed, length, nsnp, asize, bsize, wstep=500, missing=0.02, planted_every=200, planted_len=20):
    """Returns dict(pos[int32 nsnp], acodes[int8 nsnp*asize], bcodes[int8 nsnp*bsize])."""
    rng = np.random.Generator(np.random.PCG64(seed))
    if nsnp > length:
        raise ValueError("more SNPs than positions")
    if nsnp * 8 > length:          # dense: exact sampling without replacement
        pos = np.sort(rng.choice(length, size=nsnp, replace=False)).astype(np.int32)
    else:                          # sparse: draw, dedupe, top up
        pos = np.unique(rng.integers(0, length, size=int(nsnp * 1.05) + 16))
        while pos.size < nsnp:
            pos = np.unique(np.concatenate([pos, rng.integers(0, length, size=nsnp - pos.size + 16)]))
        pos = np.sort(rng.choice(pos, size=nsnp, replace=False)).astype(np.int32)
    f = np.clip(rng.beta(0.5, 0.5, size=nsnp), 0.02, 0.98)
    win = pos // wstep
    planted = (win // planted_len) % planted_every == planted_every - 1
    fb = np.where(planted, 1.0 - f, f)
    enc = np.array([3, 0, -3], dtype=np.int8)
    ga = enc[rng.binomial(2, f[:, None], size=(nsnp, asize))]
    gb = enc[rng.binomial(2, fb[:, None], size=(nsnp, bsize))]
    if missing > 0:
        ga[rng.random(ga.shape) < missing] = MISSING_I8
        gb[rng.random(gb.shape) < missing] = MISSING_I8
    return {"pos": pos, "acodes": ga.reshape(-1), "bcodes": gb.reshape(-1), "asize": asize, "bsize": bsize,
            "length": int(length)}


def reference_layout(chrom):
    """compact dict -> (avals, bvals, apos, bpos) as the reference's Cython functions take them."""
    def vals(c):
        v = c.astype(np.float64)
        v[c == MISSING_I8] = MISSING_F64
        return v
    pos = chrom["pos"]
    return (vals(chrom["acodes"]), vals(chrom["bcodes"]),
            np.repeat(pos, chrom["asize"]).astype(np.int32), np.repeat(pos, chrom["bsize"]).astype(np.int32))


def coverage_tables(seed, n, lo=20, hi=500):
    """BASELINE config 4: direct 2x2 tables with row sums U{lo..hi} and diverging allele frequencies."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n1 = rng.integers(lo, hi + 1, size=n)
    n2 = rng.integers(lo, hi + 1, size=n)
    f = np.clip(rng.beta(0.5, 0.5, size=n), 0.02, 0.98)
    fb = np.where(rng.random(n) < 0.05, 1.0 - f, np.clip(f + rng.normal(0, 0.05, size=n), 0.01, 0.99))
    a = rng.binomial(n1, f)
    c = rng.binomial(n2, fb)
    return np.stack([a, n1 - a, c, n2 - c], axis=1).astype(np.int32)


def chromosome_fast(seed, length, nsnp, asize, bsize, wstep=500, missing=0.02, planted_every=200, planted_len=20):
    """Same shape and statistics as chromosome() but ~30x faster (benchmark inputs): allele frequencies are
    quantised to 1/256 and each genotype is the sum of two byte-threshold Bernoulli draws."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pos = np.unique(rng.integers(0, length, size=int(nsnp * 1.02) + 64))
    while pos.size < nsnp:
        pos = np.unique(np.concatenate([pos, rng.integers(0, length, size=nsnp - pos.size + 64)]))
    if pos.size > nsnp:
        pos = np.sort(rng.choice(pos, size=nsnp, replace=False))
    pos = pos.astype(np.int32)
    f = np.clip(rng.beta(0.5, 0.5, size=nsnp), 0.02, 0.98)
    planted = ((pos // wstep) // planted_len) % planted_every == planted_every - 1
    fb = np.where(planted, 1.0 - f, f)
    enc = np.array([3, 0, -3], dtype=np.int8)

    def draw(freq, size):
        thr = np.round(freq * 256.0).astype(np.int16)[:, None, None]
        u = rng.integers(0, 256, size=(nsnp, size, 2), dtype=np.uint8).astype(np.int16)
        g = enc[(u < thr).sum(axis=2)]
        if missing > 0:
            g[rng.integers(0, 256, size=(nsnp, size), dtype=np.uint8) < int(round(missing * 256))] = MISSING_I8
        return g.reshape(-1)

    return {"pos": pos, "acodes": draw(f, asize), "bcodes": draw(fb, bsize), "asize": asize, "bsize": bsize,
            "length": int(length)}
