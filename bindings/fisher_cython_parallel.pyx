# cython: language_level=2
# Drop-in for the reference's statistics/fisher/fisher_cython_parallel.pyx (same def signature, lines 14-15 there):
# the extern moves from threadfisher.h:33 `threadcompute` to fpt_b200.h `fpt_fet_threadcompute`, nothing else changes.
cimport numpy as np
import numpy as np

cdef extern from "fpt_b200.h":
     int fpt_fet_threadcompute(double *avals, double *bvals, int *apos, int *bpos, int regstart, int regend, int wsize, int wstep, int alen, int blen, double perc, double *scores, double *stddev)
     const char *fpt_last_error()

def fisher_exact_tester(np.ndarray[np.float64_t, ndim=1] avals, np.ndarray[np.float64_t, ndim=1] bvals, np.ndarray[np.int32_t, ndim=1] apos, np.ndarray[np.int32_t, ndim=1] bpos, int regstart, int regend, int wsize, int wstep, int alen, int blen, double perc, np.ndarray[np.float64_t, ndim=1] scores, np.ndarray[np.float64_t, ndim=1] stddev):
    if fpt_fet_threadcompute(<double*> avals.data, <double*> bvals.data, <int*> apos.data, <int*> bpos.data, regstart, regend, wsize, wstep, alen, blen, perc, <double*> scores.data, <double*> stddev.data) != 0:
        raise RuntimeError(fpt_last_error().decode("utf-8", "replace"))
