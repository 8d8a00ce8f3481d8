/*
 * libfpt_fisher.so — the reference's literal FET entry points, so that the UNMODIFIED reference Cython modules
 * (statistics/fisher/fisher_cython_parallel.pyx:4-5,14-15 and fisher_cython.pyx:4-5,10-11) and their setup scripts link
 * against this library in place of cFisher.o + comparative.o + threadfisher.o:
 *
 *     void threadcompute(...)   fisher/threadfisher.h:33-34  ->  fpt_fet_threadcompute (libfpt_b200.so)
 *     void compute(...)         fisher/cFisher.h:11-12       ->  fpt_fet_compute
 *
 * The reference's functions return void; a failure (no CUDA device, mismatching positions, ...) is reported on stderr,
 * leaves the outputs as the caller zeroed them, and is kept for fpt_alias_status() / fpt_last_error().
 * No computation happens here.
 */
#include <stdio.h>

#include "../../../include/fpt_b200.h"

static int g_status = FPT_OK;

int fpt_alias_status(void) { return g_status; }

static void report(const char *what, int rc) {
    g_status = rc;
    if (rc != FPT_OK) fprintf(stderr, "libfpt_fisher: %s failed (%d): %s\n", what, rc, fpt_last_error());
}

void threadcompute(double *avals, double *bvals, int *apos, int *bpos, int regstart, int regend, int wsize, int wstep, int alen,
                   int blen, double perc, double *scores, double *stddev) {
    report("threadcompute", fpt_fet_threadcompute(avals, bvals, apos, bpos, regstart, regend, wsize, wstep, alen, blen, perc, scores, stddev));
}

void compute(double *avals, double *bvals, int *apos, int *bpos, int regstart, int regend, int wsize, int wstep, int alen, int blen,
             double perc, double *scores, double *stddev) {
    report("compute", fpt_fet_compute(avals, bvals, apos, bpos, regstart, regend, wsize, wstep, alen, blen, perc, scores, stddev));
}
