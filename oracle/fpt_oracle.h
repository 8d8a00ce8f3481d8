/*
 * fpt_oracle.h — CPU restatement of the reference's FET and CSS hot paths.
 *
 * TEST INFRASTRUCTURE. Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may load this
 * library; the product (libfpt_b200.so) never links, imports or falls back to it.
 *
 * Every function cites the reference code it restates (paths relative to
 * /root/reference/statistics/). Pinned against (a) the known-answer vectors of
 * fisher/testFisher.c and css/testcss.c and (b) the unmodified reference compiled into
 * oracle/_ref/ (tests/test_oracle_*.py).
 *
 * Two places where this restatement deliberately defines behaviour the reference leaves
 * undefined (both documented in DESIGN.md):
 *   - random streams are keyed per window (fpt_oracle_window_state) instead of time(NULL);
 *     inside a window the stream is the reference's own nrand48/drand48 LCG, consumed in the
 *     reference's order, so a window agrees bit-for-bit with the reference's per-window
 *     functions started from the same 48-bit state.
 *   - FET tables outside the reference's u64-binomial domain (N > 67 or numerator overflow)
 *     are evaluated by the same tail walk in log space ("log mode").
 */
#ifndef FPT_ORACLE_H
#define FPT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- random streams */
#define FPT_STREAM_RESAMPLE 0   /* bootstrap indices (FET) / label shuffles (CSS): nrand48 draws */
#define FPT_STREAM_INIT     1   /* SMACOF random starts: drand48 draws                          */

uint64_t fpt_oracle_window_state(uint64_t seed, int64_t window, int stream);
long     fpt_oracle_nrand48(uint64_t *state);               /* glibc nrand48 on a packed 48-bit state */
double   fpt_oracle_drand48(uint64_t *state);               /* glibc drand48 on a packed 48-bit state */
long     fpt_oracle_randint(long n, uint64_t *state);       /* fisher/cFisher.c:547-554, css/css.c:675-690 */

/* ---------------------------------------------------------------- windows */
/* which window indices a scan visits; mode 0 = serial `compute`, 1 = pthreads `threadcompute` */
int64_t fpt_oracle_window_count(int regend, int wsize, int wstep);
int     fpt_oracle_window_scheduled(int64_t w, int regend, int wsize, int wstep, int threaded);
/* [left,right) SNP range of window w over unique sorted positions; comparative.c:49-71 */
void    fpt_oracle_window_bounds(const int32_t *pos, int64_t nsnp, int64_t w, int wsize, int wstep,
                                 int64_t *left, int64_t *right);

/* ---------------------------------------------------------------- FET */
void     fpt_oracle_fetcount(const double *avals, const double *bvals, int64_t snp, int asize, int bsize, int f[4]);
uint64_t fpt_oracle_binomial(uint64_t n, uint64_t k);
int      fpt_oracle_fet_exact_domain(const int f[4]);
double   fpt_oracle_fet_exact(const int f[4]);              /* two-tailed P, reference arithmetic */
double   fpt_oracle_fet_neglog10(const int f[4]);           /* per-SNP score: exact domain or log mode */
double   fpt_oracle_fet_neglog10_logmode(const int f[4]);   /* force log mode (for cross-checks) */
double   fpt_oracle_percentile(double *x, int n, double q); /* sorts x in place */
void     fpt_oracle_fet_window(double *snp_scores, int npos, double perc, int nsamples, uint64_t state,
                               double out[2]);
int      fpt_oracle_fet_scan(const double *avals, const double *bvals, const int32_t *apos, const int32_t *bpos,
                             int regstart, int regend, int wsize, int wstep, int alen, int blen, double perc,
                             double *scores, double *stddev, int threaded, uint64_t seed);
/* per-SNP outputs for parity tests: tables[4*nsnp], neglog10p[nsnp] */
int      fpt_oracle_fet_per_snp(const double *avals, const double *bvals, int64_t nsnp, int asize, int bsize,
                                int32_t *tables, double *neglog10p);
void     fpt_oracle_fet_tables(const int32_t *tables, int64_t n, double *neglog10p);

/* ---------------------------------------------------------------- CSS */
void   fpt_oracle_compare_all(const double *avals, const double *bvals, int asize, int bsize, int npos, double *D);
void   fpt_oracle_compare_freq(const double *avals, const double *bvals, int npos, double *D);
int    fpt_oracle_fill_averages(double *D, int m);
void   fpt_oracle_cmds(const double *D, int m, double *X, double evals[3]);
double fpt_oracle_smacof(const double *delta, int m, double *X, int max_iters, double eps, int *iters);
double fpt_oracle_smacof_runs(const double *delta, int m, double *X, int max_iters, int n_init, double eps,
                              uint64_t *state);
void   fpt_oracle_calc_dist(const double *X, int m, double *dist);
double fpt_oracle_css(const double *dist, int m, const int *atracks, const int *btracks, int asize, int bsize);
double fpt_oracle_significance(const double *dist, int m, int *tracks, int asize, int bsize, double score,
                               int treshold, int runs, uint64_t *state, int *hits_out, int *n_out);
uint64_t fpt_oracle_lcg_skip(uint64_t state, uint64_t n);
double fpt_oracle_significance_indep(const double *dist, int m, int asize, int bsize, double score, int treshold,
                                     int runs, uint64_t state, int *hits_out, int *n_out);
void   fpt_oracle_set_perm_mode(int chain);   /* scans: 0 = independent shuffles (default), 1 = chained labels */
/* one window end to end; returns the score or -1 (discarded); *p_out untouched when discarded */
double fpt_oracle_css_window(const double *avals, const double *bvals, int asize, int bsize, int npos,
                             int drosophila, int mds, int treshold, int runs, uint64_t state_perm,
                             uint64_t state_init, double *p_out, double *X_out, double evals_out[3]);
int    fpt_oracle_css_scan(const double *avals, const double *bvals, const int32_t *apos, const int32_t *bpos,
                           int regstart, int regend, int wsize, int wstep, int alen, int blen, int treshold,
                           int runs, int drosophila, int mds, double *scores, double *p, int threaded,
                           uint64_t seed);

#ifdef __cplusplus
}
#endif
#endif
