/*
 * libfpt_css.so — the reference's literal CSS entry points, so that the UNMODIFIED reference Cython modules
 * (statistics/css/css_cython_parallel.pyx:4-5,14-15 and css_cython.pyx:4-5,10-11) and their setup scripts link against
 * this library in place of css.o + comparative.o + threadcss.o and GSL:
 *
 *     void threadcompute(...)   css/threadcss.h:36-37  ->  fpt_css_threadcompute (libfpt_b200.so)
 *     void compute(...)         css/css.h:10-11        ->  fpt_css_compute
 *
 * (FET and CSS use the same two names with different argument lists, as in the reference, hence two libraries.)
 * The reference's functions return void; a failure is reported on stderr, leaves the outputs as the caller zeroed
 * them, and is kept for fpt_alias_status() / fpt_last_error(). No computation happens here.
 */
#include <stdio.h>

#include "../../../include/fpt_b200.h"

static int g_status = FPT_OK;

int fpt_alias_status(void) { return g_status; }

static void report(const char *what, int rc) {
    g_status = rc;
    if (rc != FPT_OK) fprintf(stderr, "libfpt_css: %s failed (%d): %s\n", what, rc, fpt_last_error());
}

void threadcompute(double *avals, double *bvals, int *apos, int *bpos, int regstart, int regend, int wsize, int wstep, int alen,
                   int blen, int treshold, int runs, int drosophila, int mds, double *scores, double *p) {
    report("threadcompute", fpt_css_threadcompute(avals, bvals, apos, bpos, regstart, regend, wsize, wstep, alen, blen, treshold, runs,
                                                  drosophila, mds, scores, p));
}

void compute(double *avals, double *bvals, int *apos, int *bpos, int regstart, int regend, int wsize, int wstep, int alen, int blen,
             int treshold, int runs, int drosophila, int mdsalg, double *scores, double *p) {
    report("compute", fpt_css_compute(avals, bvals, apos, bpos, regstart, regend, wsize, wstep, alen, blen, treshold, runs, drosophila,
                                      mdsalg, scores, p));
}
