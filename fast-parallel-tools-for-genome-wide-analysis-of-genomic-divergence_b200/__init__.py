"""fpt_b200 — B200 (sm_100a) implementation of the reference's two genome-wide divergence scans.

Modules mirror the reference's Cython extension modules, so its Statistic classes stay drop-ins:

==============================  ==========================================================================
``fisher_cython_parallel``      ``fisher_exact_tester`` -> threaded-scan semantics
                                (statistics/fisher/fisher_cython_parallel.pyx:14)
``fisher_cython``               ``fisher_exact_tester`` -> serial-scan semantics (fisher_cython.pyx:10)
``css_cython_parallel``         ``cluster_separation_scorer`` (statistics/css/css_cython_parallel.pyx:14)
``css_cython``                  ``cluster_separation_scorer`` (css_cython.pyx:10)
``api``                         extended entry points: compact inputs, window ranges, probes, device API
``sharding``                    contiguous genome-range sharding over the GPUs of one node
``synth``                       synthetic genotype generator of SURVEY.md section 8(d)
``regions``                     significant-region calling (tools/SignificantCSSRegions.py, tools/FilterFisherScores.py)
``ingest``                      VCF / GTrack text -> genotype arrays (tools/VCFConvert.py), native scanner
``results``                     result-file writers / readers of the runner tools, Python-2 number formatting
``stat_shims``                  Python-3 stand-ins for FisherExactScoreStat / CategoryClusterSeparationStat `_compute`
``tools``                       per-chromosome drivers (tools/FisherExactTestSNPTool.py, tools/ClusterSeparationScore.py)
==============================  ==========================================================================

All compute goes through ``libfpt_b200.so`` (hand-written CUDA behind the C ABI of ``include/fpt_b200.h``).
There is no CPU fallback: importing works anywhere, calling a compute function without the built library
or without a CUDA device raises ``FptError``.
"""
__version__ = "0.1.0"
