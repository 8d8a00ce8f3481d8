"""Per-chromosome drivers: what the HyperBrowser runner tools do around the Statistic classes.

Reference (paths relative to /root/reference/tools/):
  FisherExactTestSNPTool.py:133-196   parse the choices, run FisherExactScoreStat over every chromosome of the genome
                                      (`reg = "*"`, `bins = "*"`: one region [0, chrLen) per chromosome), write the result file
  ClusterSeparationScore.py:144-212   the same for CategoryClusterSeparationStat

`run_manual` plays GalaxyInterface.runManual for the two statistics: one GenomeRegion per chromosome, results keyed by
region. With `world > 1` the chromosomes are dealt round-robin over the ranks (no data-path collective; rank 0 gathers the
per-chromosome arrays), which is how a multi-GPU node runs a whole genome through the unchanged per-chromosome call.
"""
from collections import OrderedDict

from . import results, stat_shims

STATS = {"FisherExactScoreStat": stat_shims.FisherExactScoreStatUnsplittable,
         "CategoryClusterSeparationStat": stat_shims.CategoryClusterSeparationStatUnsplittable}


def run_manual(tracks_a, tracks_b, analysis_def, chrom_len, parallel=True, rank=0, world=1, gather=None):
    """tracks_a / tracks_b: dict chromosome -> object with startsAsNumpyArray / valsAsNumpyArray (ingest.Population).
    chrom_len: OrderedDict chromosome -> length, in genome order (GenomeInfo.getChrList / getChrLen).
    Returns OrderedDict "chr:1-len" -> {"Result": (scores, second)} or {} for chromosomes without data (the tools print
    "skipping chr" for those). `gather(obj)` must all-gather a Python object over the ranks when world > 1."""
    kw, name = stat_shims.parse_analysis_def(analysis_def)
    cls = STATS[name]
    mine = OrderedDict()
    for k, (chrom, length) in enumerate(chrom_len.items()):
        if k % world != rank:
            continue
        region = stat_shims.GenomeRegion(chrom, 0, int(length))
        if chrom not in tracks_a or chrom not in tracks_b:
            mine[str(region)] = {}
            continue
        stat = cls(region, tracks_a[chrom], tracks_b[chrom], **kw)
        stat.parallel = parallel
        res = stat.getResult()
        mine[str(region)] = {} if isinstance(res, dict) else {"Result": res}
    if world == 1 or gather is None:         # no gather: this rank's chromosomes only
        return mine
    parts = gather(mine)
    out = OrderedDict()
    for chrom, length in chrom_len.items():
        key = str(stat_shims.GenomeRegion(chrom, 0, int(length)))
        for part in parts:
            if key in part:
                out[key] = part[key]
    return out


def _write(result, header, wstep, number):
    text = [header]
    for key, r in result.items():
        chrom = str(key).split(":")[0]
        if "Result" not in r:
            continue
        text.append(results.format_windows(chrom, wstep, r["Result"][0], r["Result"][1], number))
    return "".join(text)


def fisher_exact_test_snp_tool(tracks_a, tracks_b, chrom_len, window_size=2500, window_step=500, percentile=0.95,
                               parallel=True, number=results.str_py2, **dist):
    """FisherExactTestSNPTool.execute (tabular output): -> result file text"""
    analysis = "Dummy: dummy name ([wStep=%g] [wSize=%g] [percentile=%g])-> FisherExactScoreStat" % (window_step, window_size, percentile)
    return _write(run_manual(tracks_a, tracks_b, analysis, chrom_len, parallel=parallel, **dist), results.FET_HEADER, window_step, number)


def cluster_separation_score_tool(tracks_a, tracks_b, chrom_len, compare=False, mds=0, window_size=2500, window_step=500,
                                  mc_treshold=10, mc_runs=200000, parallel=True, number=results.str_py2, **dist):
    """ClusterSeparationScore.execute (tabular output): -> result file text. `compare` True = "average of frequencies"
    metric; mds 0 / 1 / 2 = classical / SMACOF / classical+SMACOF (ClusterSeparationScore.py:166-173)"""
    analysis = ("Dummy: dummy name ([wStep=%g] [wSize=%s] [func=%s] [mds=%s] [mcT=%s] [mcR=%s])-> CategoryClusterSeparationStat"
                % (window_step, window_size, compare, mds, mc_treshold, mc_runs))
    return _write(run_manual(tracks_a, tracks_b, analysis, chrom_len, parallel=parallel, **dist), results.CSS_HEADER, window_step, number)
