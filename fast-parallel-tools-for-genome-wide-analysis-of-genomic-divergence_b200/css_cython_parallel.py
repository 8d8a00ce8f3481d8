"""Drop-in for the reference's ``quick.statistic.css_cython_parallel`` extension module
(statistics/css/css_cython_parallel.pyx:14-15 -> ``threadcompute``, statistics/css/threadcss.c:52), as imported
by statistics/CategoryClusterSeparationStat.py:19. Runs on the GPU through ``fpt_css_threadcompute``."""
from . import _lib
from ._dropin import check_inputs, check_outputs


def cluster_separation_scorer(avals, bvals, apos, bpos, regstart, regend, wsize, wstep, alen, blen, treshold, runs,
                              drosophila, mds, scores, p):
    pa, pb, qa, qb = check_inputs(avals, bvals, apos, bpos, alen, blen)
    ps, pp = check_outputs(regend, wstep, scores=scores, p=p)
    lib = _lib.load()
    _lib.check(lib.fpt_css_threadcompute(pa, pb, qa, qb, int(regstart), int(regend), int(wsize), int(wstep), int(alen),
                                         int(blen), int(treshold), int(runs), int(drosophila), int(mds), ps, pp))
