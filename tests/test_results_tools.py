"""Result writers, per-chromosome drivers and Statistic stand-ins. On CPU the drop-in scorers are
replaced by oracle-backed stand-ins so that the host logic around them is exercised; tests/test_gpu_parity.py runs the
same pipeline through the CUDA path."""
import os
from collections import OrderedDict

import numpy as np
import pytest

import checkers
import fpt_b200.fisher_cython as fc
import fpt_b200.fisher_cython_parallel as fcp
import fpt_b200.css_cython as cc
import fpt_b200.css_cython_parallel as ccp
from checkers import dptr, iptr
from fpt_b200 import ingest, regions, results, stat_shims, synth, tools

SEED = 20261018


def test_str_py2():
    cases = {2.0: "2.0", 0.5: "0.5", 1.0 / 3: "0.333333333333", 1e-5: "1e-05", 123456789012.0: "123456789012.0",
             1234567890123.0: "1.23456789012e+12", 0.1 + 0.2: "0.3", 2.5e-10: "2.5e-10", -1.0: "-1.0", 100.0: "100.0",
             1e16: "1e+16", 3.141592653589793: "3.14159265359"}
    for x, want in cases.items():
        assert results.str_py2(x) == want, (x, results.str_py2(x))
        assert results.str_py2(np.float64(x)) == want
    assert results.str_py2(float("nan")) == "nan" and results.str_py2(float("inf")) == "inf" and results.str_py2(-float("inf")) == "-inf"


def test_format_and_read_back():
    scores = np.array([0.0, 1.5, 0.0, 2.0, 1.0 / 3])
    second = np.array([9.0, 0.25, 9.0, 0.0, 1e-7])
    text = results.FET_HEADER + results.format_windows("chrI", 500, scores, second)
    assert text == "#seqid\tstart\tscore\tstddev\nchrI\t500\t1.5\t0.25\nchrI\t1500\t2.0\t0.0\nchrI\t2000\t0.333333333333\t1e-07\n"
    chroms, starts, s, t = results.read_scan(text)
    assert chroms == ["chrI"] * 3 and starts.tolist() == [500, 1500, 2000]
    assert s.tolist() == [1.5, 2.0, 0.333333333333] and t.tolist() == [0.25, 0.0, 1e-07]
    exact = results.format_windows("chrI", 500, scores, second, number=results.str_exact)
    assert results.read_scan(exact)[2].tolist() == [1.5, 2.0, 1.0 / 3]


def test_region_files():
    lines = ["chrI\t%d\t%s\t%s" % (i * 500, s, p) for i, (s, p) in enumerate(
        [(3.0, 0.001), (3.1, 0.001), (1.0, 0.6), (1.0, 0.7), (1.0, 0.5), (1.0, 0.5), (1.0, 0.5), (1.0, 0.5), (4.0, 0.002)])]
    text = results.CSS_HEADER + "\n".join(lines) + "\n"
    out = results.significant_css_regions_file(text, 2500, {"chrI": 6000}, fdr=0.05)
    assert out == results.CSS_REGION_HEADER + "chrI\t0\t3000\nchrI\t4000\t5999\n"
    assert results.significant_css_regions_file(text, 2500, {"chrI": 6000}, fdr=1e-9) == "NONE found"
    top = results.significant_css_regions_file(text, 2500, {"chrI": 100000}, num_top=1)
    assert top == results.CSS_REGION_HEADER + "chrI\t4000\t6500\n"
    fet = results.FET_HEADER + "".join("c\t%d\t%s\t%s\n" % (i * 500, 1.0 + (i == 7) * 50, 0.1) for i in range(40))
    out = results.filter_fisher_scores_file(fet, 2500, {"c": 10**6}, 0.999, 75.0)
    assert out == results.FET_REGION_HEADER + "c\t3500\t6000\n"


def test_parse_analysis_def():
    kw, name = stat_shims.parse_analysis_def("Dummy: dummy name ([wStep=500] [wSize=2500] [func=False] [mds=0] [mcT=10] [mcR=200000])-> CategoryClusterSeparationStat")
    assert name == "CategoryClusterSeparationStat"
    assert kw == {"wStep": "500", "wSize": "2500", "func": "False", "mds": "0", "mcT": "10", "mcR": "200000"}


@pytest.fixture()
def oracle_dropins(monkeypatch):
    """oracle-backed stand-ins with the drop-in argument lists"""
    o = checkers.load_oracle()

    def fet(threaded):
        def f(avals, bvals, apos, bpos, regstart, regend, wsize, wstep, alen, blen, perc, scores, stddev):
            o.fpt_oracle_fet_scan(dptr(avals), dptr(bvals), iptr(apos), iptr(bpos), regstart, regend, wsize, wstep, alen, blen,
                                  perc, dptr(scores), dptr(stddev), threaded, SEED)
        return f

    def css(threaded):
        def f(avals, bvals, apos, bpos, regstart, regend, wsize, wstep, alen, blen, treshold, runs, drosophila, mds, scores, p):
            o.fpt_oracle_css_scan(dptr(avals), dptr(bvals), iptr(apos), iptr(bpos), regstart, regend, wsize, wstep, alen, blen,
                                  treshold, runs, drosophila, mds, dptr(scores), dptr(p), threaded, SEED)
        return f

    monkeypatch.setattr(fcp, "fisher_exact_tester", fet(1))
    monkeypatch.setattr(fc, "fisher_exact_tester", fet(0))
    monkeypatch.setattr(ccp, "cluster_separation_scorer", css(1))
    monkeypatch.setattr(cc, "cluster_separation_scorer", css(0))
    return o


def _genome(nchrom=3, length=60000, nsnp=1500, asize=6, bsize=5):
    ta, tb, lens = OrderedDict(), OrderedDict(), OrderedDict()
    for c in range(nchrom):
        ch = synth.chromosome(100 + c, length, nsnp, asize, bsize)
        comp = ingest.CompactChromosome(ch["pos"], ch["acodes"].reshape(-1, asize), ch["bcodes"].reshape(-1, bsize))
        a, b = comp.reference_layout()
        name = "chr%d" % (c + 1)
        lens[name] = length
        if c != 1:                       # chromosome 2 has no data in either track: "skipping chr"
            ta[name], tb[name] = a, b
    return ta, tb, lens


def test_stat_shims_compute(oracle_dropins):
    ta, tb, lens = _genome()
    reg = stat_shims.GenomeRegion("chr1", 0, 60000)
    st = stat_shims.FisherExactScoreStatUnsplittable(reg, ta["chr1"], tb["chr1"], wSize="2500", wStep="500", percentile="0.95")
    scores, stddev = st.getResult()
    assert scores.shape == (120,) and stddev.shape == (120,) and (scores != 0).sum() >= 100
    want_s, want_d = np.zeros(120), np.zeros(120)
    a, b = ta["chr1"], tb["chr1"]
    oracle_dropins.fpt_oracle_fet_scan(dptr(a.vals), dptr(b.vals), iptr(a.starts), iptr(b.starts), 0, 60000, 2500, 500, a.vals.size,
                                       b.vals.size, 0.95, dptr(want_s), dptr(want_d), 1, SEED)
    assert np.array_equal(scores, want_s) and np.array_equal(stddev, want_d)
    # empty track: the {"0": 0} sentinel of FisherExactScoreStat.py:36-37
    empty = ingest.Population(np.zeros(0, np.int32), np.zeros(0))
    assert stat_shims.FisherExactScoreStatUnsplittable(reg, empty, empty, wSize="2500", wStep="500", percentile="0.95").getResult() == {"0": 0}
    cs = stat_shims.CategoryClusterSeparationStatUnsplittable(reg, ta["chr1"], tb["chr1"], wSize="2500", wStep="500", mcT="5", mcR="40",
                                                              func="False", mds="0")
    s, p = cs.getResult()
    assert s.shape == (120,) and (s != 0).sum() >= 100 and np.all((p[s != 0] > 0) & (p[s != 0] <= 1))


def test_tools_write_whole_genome(oracle_dropins):
    ta, tb, lens = _genome()
    text = tools.fisher_exact_test_snp_tool(ta, tb, lens, 2500, 500, 0.95)
    lines = text.split("\n")
    assert lines[0] == "#seqid\tstart\tscore\tstddev" and lines[-1] == ""
    chroms, starts, s, d = results.read_scan(text)
    assert sorted(set(chroms)) == ["chr1", "chr3"] and chroms == sorted(chroms)
    assert np.all(s != 0) and np.all(starts % 500 == 0)
    # threaded semantics: 120 windows >= 103, so the pthreads scan scores them; serial agrees on the values
    serial = tools.fisher_exact_test_snp_tool(ta, tb, lens, 2500, 500, 0.95, parallel=False)
    cs, ss, s2, _ = results.read_scan(serial)
    common = dict(zip(zip(chroms, starts.tolist()), s.tolist()))
    assert all(common.get(k, v) == v for k, v in zip(zip(cs, ss.tolist()), s2.tolist()))
    css = tools.cluster_separation_score_tool(ta, tb, lens, compare=False, mds=0, mc_treshold=5, mc_runs=40)
    assert css.startswith("#seqid\tstart\tscore\tp\n")
    c3, st3, s3, p3 = results.read_scan(css)
    assert len(c3) >= 200 and np.all((p3 > 0) & (p3 <= 1))
    # downstream region callers accept what the tools wrote
    out = results.significant_css_regions_file(css, 2500, lens, num_top=5)
    assert out.startswith(results.CSS_REGION_HEADER) and out.count("\n") >= 6
    out = results.filter_fisher_scores_file(text, 2500, lens, 0.9, 75.0)
    assert out.startswith(results.FET_REGION_HEADER)


def test_tools_round_robin_over_ranks(oracle_dropins):
    ta, tb, lens = _genome()
    whole = tools.fisher_exact_test_snp_tool(ta, tb, lens, 2500, 500, 0.95)
    analysis = "Dummy: dummy name ([wStep=500] [wSize=2500] [percentile=0.95])-> FisherExactScoreStat"
    parts = [tools.run_manual(ta, tb, analysis, lens, rank=r, world=2) for r in range(2)]
    assert list(parts[0]) == ["chr1:1-60000", "chr3:1-60000"] and list(parts[1]) == ["chr2:1-60000"]
    merged = tools.run_manual(ta, tb, analysis, lens, rank=0, world=2, gather=lambda mine: parts)
    assert tools._write(merged, results.FET_HEADER, 500, results.str_py2) == whole
