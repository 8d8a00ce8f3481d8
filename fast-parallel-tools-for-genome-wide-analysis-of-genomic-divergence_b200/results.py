"""Result files of the two scans and of the region callers — the text the reference's HyperBrowser
tools write and read back.

Reference (paths relative to /root/reference/tools/):
  FisherExactTestSNPTool.py:162-189   "#seqid\\tstart\\tscore\\tstddev", one line per window with score != 0,
                                      start = index * wStep, numbers written with Python 2 `str()`
  ClusterSeparationScore.py:180-205   the same with "#seqid\\tstart\\tscore\\tp"
  SignificantCSSRegions.py:86-88,155-175  reading a result file back (`preProcessPvalues`)
  SignificantCSSRegions.py:129-153    segments GTrack header + merged regions
  FilterFisherScores.py:73-80,97-114  the same for FET (one more header line)
"""
import math

import numpy as np

from . import regions

FET_HEADER = "#seqid\tstart\tscore\tstddev\n"
CSS_HEADER = "#seqid\tstart\tscore\tp\n"
CSS_REGION_HEADER = ("##gtrack version: 1.0\n##track type: segments\n##uninterrupted data lines: true\n"
                     "##no overlapping elements: true\n###seqid\tstart\tend\n")
FET_REGION_HEADER = ("##gtrack version: 1.0\n##track type: segments\n##uninterrupted data lines: true\n"
                     "##sorted elements: false\n##no overlapping elements: true\n###seqid\tstart\tend\n")


def str_py2(x):
    """Python 2 `str()` of a float / numpy.float64, which is what the reference's writers emit: 12 significant digits
    (`%.12g`), with ".0" appended when the result looks like an integer."""
    x = float(x)
    if math.isnan(x):
        return "nan"
    if math.isinf(x):
        return "inf" if x > 0 else "-inf"
    s = "%.12g" % x
    if "." not in s and "e" not in s:
        s += ".0"
    return s


def str_exact(x):
    """shortest string that reads back to the same float64 (Python 3 `repr`)"""
    return repr(float(x))


def format_windows(chrom, wstep, scores, second, number=str_py2):
    """the data lines of one chromosome (windows with score == 0 are skipped, as the tools do). `number` formats the two
    float columns: `str_py2` reproduces the reference's files, `str_exact` keeps every bit."""
    scores = np.asarray(scores)
    second = np.asarray(second)
    return "".join("%s\t%s\t%s\t%s\n" % (str(chrom), int(i) * wstep, number(scores[i]), number(second[i]))
                   for i in np.nonzero(scores != 0)[0])


def write_scan(path_or_file, kind, per_chromosome, wstep, number=str_py2):
    """whole result file. `kind` is "fet" or "css"; `per_chromosome` yields (chrom, scores, second) in genome order."""
    header = {"fet": FET_HEADER, "css": CSS_HEADER}[kind]
    fh = open(path_or_file, "w") if isinstance(path_or_file, str) else path_or_file
    try:
        fh.write(header)
        for chrom, scores, second in per_chromosome:
            fh.write(format_windows(chrom, wstep, scores, second, number))
    finally:
        if isinstance(path_or_file, str):
            fh.close()


def read_scan(text):
    """`preProcessPvalues` of the region tools: -> (chroms, starts, score column, second column). Lines that are empty or
    start with '#' are skipped; columns are tab-separated."""
    chroms, starts, c2, c3 = [], [], [], []
    for line in text.split("\n"):
        if (not line) or line[0] == "#":
            continue
        cols = line.split("\t")
        chroms.append(cols[0])
        starts.append(int(cols[1]))
        c2.append(float(cols[2]))
        c3.append(float(cols[3]))
    return chroms, np.asarray(starts, dtype=np.int64), np.asarray(c2, dtype=np.float64), np.asarray(c3, dtype=np.float64)


def format_regions(segments, header):
    """merged (chrom, start, end) segments -> segments GTrack text"""
    return header + "".join("%s\t%s\t%s\n" % (c, s, e) for c, s, e in segments)


def significant_css_regions_file(result_text, window_size, chrom_len, fdr=None, num_top=None):
    """tools/SignificantCSSRegions.py `execute`: result file text -> output file text ("NONE found" when the
    Benjamini-Hochberg walk accepts nothing, :116-120)"""
    chroms, starts, scores, p = read_scan(result_text)
    if fdr is not None and regions.css_fdr_threshold(p, fdr) is None:
        return "NONE found"
    return format_regions(regions.css_significant_regions(chroms, starts, scores, p, window_size, chrom_len, fdr=fdr, num_top=num_top),
                          CSS_REGION_HEADER)


def filter_fisher_scores_file(result_text, window_size, chrom_len, normquantile, percentile):
    """tools/FilterFisherScores.py `execute`: result file text -> output file text"""
    chroms, starts, scores, stddevs = read_scan(result_text)
    return format_regions(regions.fet_significant_regions(chroms, starts, scores, stddevs, window_size, chrom_len,
                                                          normquantile=normquantile, percentile=percentile), FET_REGION_HEADER)
