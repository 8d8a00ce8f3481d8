"""Drop-in for the reference's ``quick.statistic.fisher_cython_parallel`` extension module
(statistics/fisher/fisher_cython_parallel.pyx:14-15 -> ``threadcompute``, statistics/fisher/threadfisher.c:47),
as imported by statistics/FisherExactScoreStat.py:19. Same name, same positional arguments, same in-place
outputs; the work runs on the GPU through ``fpt_fet_threadcompute`` of libfpt_b200.so."""
from . import _lib
from ._dropin import check_inputs, check_outputs


def fisher_exact_tester(avals, bvals, apos, bpos, regstart, regend, wsize, wstep, alen, blen, perc, scores, stddev):
    pa, pb, qa, qb = check_inputs(avals, bvals, apos, bpos, alen, blen)
    ps, pd = check_outputs(regend, wstep, scores=scores, stddev=stddev)
    lib = _lib.load()
    _lib.check(lib.fpt_fet_threadcompute(pa, pb, qa, qb, int(regstart), int(regend), int(wsize), int(wstep), int(alen),
                                         int(blen), float(perc), ps, pd))
