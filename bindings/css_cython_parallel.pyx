# cython: language_level=2
# Drop-in for statistics/css/css_cython_parallel.pyx (lines 14-15 there): threadcss.h:36 `threadcompute` -> fpt_css_threadcompute
cimport numpy as np
import numpy as np

cdef extern from "fpt_b200.h":
     int fpt_css_threadcompute(double *avals, double *bvals, int *apos, int *bpos, int regstart, int regend, int wsize, int wstep, int alen, int blen, int treshold, int runs, int drosophila, int mds, double *scores, double *p)
     const char *fpt_last_error()

def cluster_separation_scorer(np.ndarray[np.float64_t, ndim=1] avals, np.ndarray[np.float64_t, ndim=1] bvals, np.ndarray[np.int32_t, ndim=1] apos, np.ndarray[np.int32_t, ndim=1] bpos, int regstart, int regend, int wsize, int wstep, int alen, int blen, int treshold, int runs, int drosophila, int mds, np.ndarray[np.float64_t, ndim=1] scores, np.ndarray[np.float64_t, ndim=1] p):
    if fpt_css_threadcompute(<double*> avals.data, <double*> bvals.data, <int*> apos.data, <int*> bpos.data, regstart, regend, wsize, wstep, alen, blen, treshold, runs, drosophila, mds, <double*> scores.data, <double*> p.data) != 0:
        raise RuntimeError(fpt_last_error().decode("utf-8", "replace"))
