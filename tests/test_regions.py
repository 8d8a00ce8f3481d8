"""Region calling: our host functions against a literal Python-3 transliteration of the loops in
tools/SignificantCSSRegions.py:97-153 and tools/FilterFisherScores.py:84-114 (the transliteration below IS the checker for
this piece: the reference tools are Python 2 + HyperBrowser and cannot be imported)."""
import numpy as np
import pytest


def ref_css_regions(addr, scores, p, window_size, chrlen, fdr=None, num_top=None):
    addrs = np.array(addr)
    if fdr is not None:
        psorted = np.argsort(p)[::-1]
        k = float(len(p)); n = k; testp = 0
        for pi in psorted:
            if p[pi] <= k / n * fdr:
                testp = p[pi]
                break
            k -= 1
        if k == 0:
            return None
        filtered = addrs[p <= testp]
    else:
        scoresorted = np.argsort(scores)[::-1]
        filtered = addrs[scores >= scores[scoresorted[num_top - 1]]]
    return ref_merge(filtered, window_size, chrlen)


def ref_merge(filtered, window_size, chrlen):
    out = []
    curchrom, start, end, prev = "", "", 2 ** 63, -1000000.
    for addr in filtered:
        al = addr.split("\t")
        if al[0] != curchrom or int(al[1]) - window_size > prev:
            if curchrom != "":
                newend = prev + window_size if prev + window_size < end else end
                out.append(start + "\t" + str(newend))
            start = addr
            curchrom = al[0]
            end = int(chrlen[curchrom]) - 1
        prev = int(addr.split("\t")[1])
    newend = prev + window_size if prev + window_size < end else end
    out.append(start + "\t" + str(newend))
    return out


def _windows(rng, nchrom=3, nwin=400, wstep=500):
    chroms, starts = [], []
    chrlen = {}
    for c in range(nchrom):
        name = "chr%d" % (c + 1)
        keep = np.sort(rng.choice(nwin, size=int(nwin * 0.8), replace=False))
        chroms += [name] * keep.size
        starts += list(keep * wstep)
        chrlen[name] = nwin * wstep + int(rng.integers(0, 400))
    return chroms, np.array(starts), chrlen


@pytest.mark.parametrize("seed", range(6))
def test_css_regions_match_transliteration(seed):
    from fpt_b200 import regions
    rng = np.random.default_rng(seed)
    chroms, starts, chrlen = _windows(rng)
    n = len(chroms)
    scores = rng.normal(size=n)
    p = rng.uniform(size=n) ** (3 if seed % 2 else 1)
    p[rng.choice(n, size=n // 10, replace=False)] = 1.0 / 1001
    addr = ["%s\t%d" % (c, s) for c, s in zip(chroms, starts)]
    for kw in (dict(fdr=0.05), dict(fdr=0.5), dict(fdr=1e-6), dict(num_top=25), dict(num_top=1)):
        want = ref_css_regions(addr, scores, p, 2500, chrlen, **kw)
        got = regions.css_significant_regions(chroms, starts, scores, p, 2500, chrlen, **kw)
        if want is None:
            assert got == []
        else:
            assert ["%s\t%d\t%d" % g for g in got] == want


def test_merge_rule_and_clamp():
    from fpt_b200.regions import merge_windows
    L = {"a": 10000, "b": 3000}
    # gap of exactly window_size continues a segment (strict `>`), a larger gap starts a new one; end clamps to chrLen-1
    got = merge_windows(["a", "a", "a", "a", "b"], [0, 2500, 5001, 9000, 500], 2500, L)
    assert got == [("a", 0, 5000), ("a", 5001, 7501), ("a", 9000, 9999), ("b", 500, 2999)]
    assert merge_windows([], [], 2500, L) == []


def test_fet_limit_pieces():
    from fpt_b200 import regions
    rng = np.random.default_rng(1)
    x = rng.gamma(2.0, 1.0, size=5000)
    assert abs(regions.cmedian(x) - np.median(x)) < 5 * (x.max() - x.min()) / 1000      # within a few bins of the true median
    assert regions.scoreatpercentile(np.arange(11.0), 75.0) == 7.5
    assert regions.scoreatpercentile([3.0, 1.0, 2.0], 50.0) == 2.0
    sd = rng.uniform(0.1, 0.4, size=5000)
    lim = regions.fet_limit(x, sd, 0.999, 75.0)
    assert lim == pytest.approx(regions.cmedian(x) + 3.090232306167813 * np.percentile(sd, 75.0), rel=1e-12)
    chroms, starts, chrlen = _windows(rng, nchrom=2, nwin=3125)
    n = len(chroms)
    got = regions.fet_significant_regions(chroms, starts, x[:n], sd[:n], 2500, chrlen)
    mask = x[:n] >= regions.fet_limit(x[:n], sd[:n], 0.999, 75.0)
    addr = np.array(["%s\t%d" % (c, s) for c, s in zip(chroms, starts)])
    assert ["%s\t%d\t%d" % g for g in got] == ref_merge(addr[mask], 2500, chrlen)


def test_scan_to_windows_skips_zero_scores():
    from fpt_b200.regions import scan_to_windows
    c, s, sc, se = scan_to_windows("chr1", 500, np.array([0.0, 1.5, 0.0, -0.0, 2.0]), np.array([9.0, 0.1, 9.0, 9.0, 0.2]))
    assert c == ["chr1", "chr1"] and list(s) == [500, 2000] and list(sc) == [1.5, 2.0] and list(se) == [0.1, 0.2]
