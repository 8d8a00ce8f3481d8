"""Python-3 stand-ins for the HyperBrowser Statistic classes that call the scans.

The reference's `FisherExactScoreStatUnsplittable._compute` (statistics/FisherExactScoreStat.py:28-61) and
`CategoryClusterSeparationStatUnsplittable._compute` (statistics/CategoryClusterSeparationStat.py:31-80) are written
against `gold.statistic.Statistic.MultipleRawDataStatistic`, which is not part of the reference repository. These
classes keep what `_compute` relies on — `self._region` (start/end), `self._children[k].getResult()` (objects with
`startsAsNumpyArray` / `valsAsNumpyArray`), `self._kwArgs` (strings, as the analysis-definition parser delivers them) —
and restate the two `_compute` bodies for Python 3 (`regend // wStep` where Python 2 wrote `regend/wStep`).
"""
import numpy

from . import css_cython, css_cython_parallel, fisher_cython, fisher_cython_parallel


class GenomeRegion:
    def __init__(self, chr, start, end, genome=None):
        self.genome, self.chr, self.start, self.end = genome, chr, int(start), int(end)

    def __str__(self):
        return "%s:%d-%d" % (self.chr, self.start + 1, self.end)


class RawDataStat:
    """child statistic whose result is one population's track view on the region"""

    def __init__(self, track_view):
        self._tv = track_view

    def getResult(self):
        return self._tv


class MultipleRawDataStatistic:
    def __init__(self, region, track, track2, **kwArgs):
        self._region = region
        self._children = [RawDataStat(track), RawDataStat(track2)]
        self._kwArgs = kwArgs
        self._result = None

    def getResult(self):
        if self._result is None:
            self._result = self._compute()
        return self._result


class FisherExactScoreStatUnsplittable(MultipleRawDataStatistic):
    """kwArgs: wSize, wStep, percentile; `parallel=False` selects the serial import the reference keeps commented out
    (FisherExactScoreStat.py:18-19)"""
    parallel = True

    def _compute(self):
        reg = self._region
        groupA = self._children[0].getResult()
        groupB = self._children[1].getResult()
        apos = groupA.startsAsNumpyArray()
        if len(apos) == 0:
            return {"0": 0}
        bpos = groupB.startsAsNumpyArray()
        avals = groupA.valsAsNumpyArray()
        bvals = groupB.valsAsNumpyArray()
        regstart, regend = reg.start, reg.end
        wSize = int(self._kwArgs["wSize"])
        wStep = int(self._kwArgs["wStep"])
        alen, blen = avals.shape[0], bvals.shape[0]
        perc = float(self._kwArgs["percentile"])
        num_win = regend // wStep
        scores = numpy.zeros(num_win)
        stddev = numpy.zeros(num_win)
        tester = (fisher_cython_parallel if self.parallel else fisher_cython).fisher_exact_tester
        tester(avals, bvals, apos, bpos, regstart, regend, wSize, wStep, alen, blen, perc, scores, stddev)
        return scores, stddev


class CategoryClusterSeparationStatUnsplittable(MultipleRawDataStatistic):
    """kwArgs: wSize, wStep, mcT, mcR, func ("True" selects the frequency metric), mds (0 classical, 1 SMACOF,
    2 classical then SMACOF)"""
    parallel = True

    def _compute(self):
        reg = self._region
        groupA = self._children[0].getResult()
        groupB = self._children[1].getResult()
        apos = groupA.startsAsNumpyArray()
        if len(apos) == 0:
            return {"0": 0}
        bpos = groupB.startsAsNumpyArray()
        avals = groupA.valsAsNumpyArray()
        bvals = groupB.valsAsNumpyArray()
        regstart, regend = reg.start, reg.end
        wSize = int(self._kwArgs["wSize"])
        wStep = int(self._kwArgs["wStep"])
        alen, blen = avals.shape[0], bvals.shape[0]
        treshold = int(self._kwArgs["mcT"])
        runs = int(self._kwArgs["mcR"])
        drosophila = 1 if str(self._kwArgs["func"]) == "True" else 0
        mds = int(self._kwArgs["mds"])
        num_win = regend // wStep
        scores = numpy.zeros(num_win)
        p = numpy.zeros(num_win)
        scorer = (css_cython_parallel if self.parallel else css_cython).cluster_separation_scorer
        scorer(avals, bvals, apos, bpos, regstart, regend, wSize, wStep, alen, blen, treshold, runs, drosophila, mds, scores, p)
        return scores, p


def parse_analysis_def(text):
    """"Dummy: dummy name ([wStep=500] [wSize=2500] [percentile=0.95])-> FisherExactScoreStat" -> (kwArgs of strings,
    statistic name); the form the runner tools build (FisherExactTestSNPTool.py:166, ClusterSeparationScore.py:187)"""
    import re
    kw = dict(re.findall(r"\[(\w+)=([^\]]*)\]", text))
    return kw, text.rsplit("->", 1)[1].strip()
