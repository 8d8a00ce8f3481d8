"""Significant-region calling on the per-window outputs of the two scans (host side, no GPU involved) — SURVEY.md
section 8(f) row 1. Restates the decision logic of the reference's two HyperBrowser tools so that "significant-region
calls" can be compared bit for bit:

* ``tools/SignificantCSSRegions.py:97-153``  — Benjamini-Hochberg step over the permutation p-values (or top-N scores),
  then merging of neighbouring significant windows into segments
* ``tools/FilterFisherScores.py:84-114``     — FET limit = cmedian(scores) + Phi^-1(q) * percentile(stddevs), same merge

Windows are given in file order: grouped by chromosome, ascending start (the order the runner tools write them,
tools/ClusterSeparationScore.py:190-205, tools/FisherExactTestSNPTool.py:176-189, which also skip windows with score 0).
"""
import numpy as np


def merge_windows(chroms, starts, window_size, chrom_len):
    """tools/SignificantCSSRegions.py:137-153 (identical loop in FilterFisherScores.py:97-114).
    A new segment starts when the chromosome changes or ``start - window_size > previous start``; a segment ends at
    ``previous start + window_size`` clamped to ``chrom_len[chrom] - 1``. Returns [(chrom, start, end)]."""
    out = []
    cur, seg_start, end, prev = None, None, None, -1000000
    for c, s in zip(chroms, starts):
        s = int(s)
        if c != cur or s - window_size > prev:
            if cur is not None:
                out.append((cur, seg_start, min(prev + window_size, end)))
            cur, seg_start, end = c, s, int(chrom_len[c]) - 1
        prev = s
    if cur is not None:
        out.append((cur, seg_start, min(prev + window_size, end)))
    return out


def css_fdr_threshold(p, fdr):
    """Benjamini-Hochberg step of tools/SignificantCSSRegions.py:104-114: walk the p-values from the largest down, k from n
    down, and stop at the first p <= k/n * FDR. Returns that p (the windows with p <= it are significant) or None."""
    p = np.asarray(p, dtype=np.float64)
    n = float(p.size)
    k = n
    for i in np.argsort(p)[::-1]:
        if p[i] <= k / n * fdr:
            return float(p[i])
        k -= 1
    return None


def css_significant_mask(scores, p, fdr=None, num_top=None):
    """boolean mask of significant windows: FDR filtering (SignificantCSSRegions.py:104-123) or the ``num_top`` best
    scores (:125-127; ties with the num_top-th score are included, as there)"""
    scores = np.asarray(scores, dtype=np.float64)
    p = np.asarray(p, dtype=np.float64)
    if (fdr is None) == (num_top is None):
        raise ValueError("give exactly one of fdr / num_top")
    if fdr is not None:
        t = css_fdr_threshold(p, fdr)
        return np.zeros(p.size, dtype=bool) if t is None else p <= t
    order = np.argsort(scores)[::-1]
    return scores >= scores[order[num_top - 1]]


def css_significant_regions(chroms, starts, scores, p, window_size, chrom_len, fdr=None, num_top=None):
    mask = css_significant_mask(scores, p, fdr=fdr, num_top=num_top)
    chroms = np.asarray(chroms, dtype=object)
    return merge_windows(chroms[mask], np.asarray(starts)[mask], window_size, chrom_len)


def cmedian(a, numbins=1000):
    """The histogram-interpolated median the reference calls as ``stats.cmedian`` (tools/FilterFisherScores.py:84): old
    scipy (< 0.13) binned the data into ``numbins`` equal bins spanning [min - w/2, max + w/2], located the bin holding the
    n/2-th value in the cumulative histogram and interpolated linearly inside it. Modern scipy no longer ships it, so the
    published algorithm is restated here (PARITY UNPINNED: no old scipy in this image to run against)."""
    a = np.ravel(np.asarray(a, dtype=np.float64))
    n = float(a.size)
    amin, amax = a.min(), a.max()
    est = (amax - amin) / float(numbins - 1)
    binsize = (amax - amin + est) / float(numbins)
    hist, bins = np.histogram(a, numbins, range=(amin - binsize * 0.5, amax + binsize * 0.5))
    binsize = bins[1] - bins[0]
    cum = np.cumsum(hist)
    cfbin = int(np.searchsorted(cum, n / 2.0))
    lrl = bins[cfbin]
    below = 0.0 if cfbin == 0 else float(cum[cfbin - 1])
    return lrl + ((n / 2.0 - below) / float(hist[cfbin])) * binsize


def scoreatpercentile(a, per):
    """scipy.stats.scoreatpercentile as the reference uses it (FilterFisherScores.py:85): linear interpolation between
    the order statistics at index per/100 * (n-1)"""
    v = np.sort(np.ravel(np.asarray(a, dtype=np.float64)))
    idx = per / 100.0 * (v.size - 1)
    i = int(idx)
    if i == idx:
        return float(v[i])
    return float(v[i] + (v[i + 1] - v[i]) * (idx - i))


def fet_limit(scores, stddevs, normquantile, percentile):
    """tools/FilterFisherScores.py:84-87: limit = cmedian(scores) + Phi^-1(normquantile) * scoreatpercentile(stddevs, percentile)"""
    from scipy.stats import norm
    return cmedian(scores) + float(norm.ppf(normquantile)) * scoreatpercentile(stddevs, percentile)


def fet_significant_regions(chroms, starts, scores, stddevs, window_size, chrom_len, normquantile=0.999, percentile=75.0):
    """windows with score >= limit, merged (FilterFisherScores.py:88-114; defaults from :40-48)"""
    scores = np.asarray(scores, dtype=np.float64)
    mask = scores >= fet_limit(scores, stddevs, normquantile, percentile)
    chroms = np.asarray(chroms, dtype=object)
    return merge_windows(chroms[mask], np.asarray(starts)[mask], window_size, chrom_len)


def scan_to_windows(chrom, wstep, scores, second):
    """per-chromosome output arrays -> the rows the runner tools write: start = index * wstep, windows with score == 0 skipped
    (tools/FisherExactTestSNPTool.py:181-189, tools/ClusterSeparationScore.py:199-205). Returns (chroms, starts, scores, second)."""
    scores = np.asarray(scores)
    keep = np.nonzero(scores != 0)[0]
    return [chrom] * keep.size, keep * wstep, scores[keep], np.asarray(second)[keep]
