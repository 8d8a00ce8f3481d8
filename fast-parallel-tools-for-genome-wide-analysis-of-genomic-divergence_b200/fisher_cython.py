"""Drop-in for the reference's serial ``quick.statistic.fisher_cython`` module
(statistics/fisher/fisher_cython.pyx:10-11 -> ``compute``, statistics/fisher/cFisher.c:38): every window of
the region is visited, including regions too short for the threaded scan."""
from . import _lib
from ._dropin import check_inputs, check_outputs


def fisher_exact_tester(avals, bvals, apos, bpos, regstart, regend, wsize, wstep, alen, blen, perc, scores, stddev):
    pa, pb, qa, qb = check_inputs(avals, bvals, apos, bpos, alen, blen)
    ps, pd = check_outputs(regend, wstep, scores=scores, stddev=stddev)
    lib = _lib.load()
    _lib.check(lib.fpt_fet_compute(pa, pb, qa, qb, int(regstart), int(regend), int(wsize), int(wstep), int(alen),
                                   int(blen), float(perc), ps, pd))
