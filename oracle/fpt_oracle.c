/*
 * fpt_oracle.c — plain-C restatement of the reference's FET and CSS hot paths (see fpt_oracle.h).
 * TEST INFRASTRUCTURE ONLY — never linked into or called from the product library.
 *
 * Reference paths are relative to /root/reference/statistics/.
 */
#define _GNU_SOURCE
#include "fpt_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================= random streams
 * glibc's nrand48/drand48 are the 48-bit LCG X <- (0x5DEECE66D X + 0xB) mod 2^48; nrand48 returns the
 * top 31 bits, drand48 returns X / 2^48. The reference draws bootstrap indices and label shuffles
 * with nrand48 on a private state (fisher/cFisher.c:547-554, css/css.c:675-690) and SMACOF starts
 * with drand48 (css/css.c:863-864). The state is kept packed in one uint64_t here.
 */
#define LCG_A 0x5DEECE66DULL
#define LCG_C 0xBULL
#define MASK48 0xFFFFFFFFFFFFULL

uint64_t fpt_oracle_window_state(uint64_t seed, int64_t window, int stream) {
    /* splitmix64 finaliser over (seed, window, stream); our own keying, the reference uses time(NULL) */
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL * (uint64_t)(4 * (uint64_t)window + (uint64_t)stream + 1);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return z & MASK48;
}

static inline uint64_t lcg_step(uint64_t *s) {
    *s = (*s * LCG_A + LCG_C) & MASK48;
    return *s;
}

long fpt_oracle_nrand48(uint64_t *state) { return (long)(lcg_step(state) >> 17); }

double fpt_oracle_drand48(uint64_t *state) { return ldexp((double)lcg_step(state), -48); }

long fpt_oracle_randint(long n, uint64_t *state) {
    /* rejection above RAND_MAX - (RAND_MAX+1) % n, then modulo: cFisher.c:547-554 */
    const long rmax = 2147483647L;
    long limit = rmax - (rmax + 1) % n;
    long r = fpt_oracle_nrand48(state);
    while (r > limit) r = fpt_oracle_nrand48(state);
    return r % n;
}

/* ======================================================================= windows
 * Serial scan: fisher/cFisher.c:81-99, css/css.c:117-135 — window w spans [w*wstep, w*wstep+wsize],
 * both ends inclusive (comparative.c:58-65), visited while start + wsize <= regend + wstep.
 * Threaded scan: fisher/threadfisher.c:55-58,191-238 — tasks 0..num_tasks (inclusive) of 100 windows,
 * num_tasks = (regend/wstep - 3)/100, nothing at all when num_tasks == 0, last stop clipped to
 * regend + wstep.
 */
int64_t fpt_oracle_window_count(int regend, int wsize, int wstep) {
    int64_t lim = (int64_t)regend + wstep - wsize;
    if (lim < 0) return 0;
    return lim / wstep + 1;
}

int fpt_oracle_window_scheduled(int64_t w, int regend, int wsize, int wstep, int threaded) {
    int64_t start = w * (int64_t)wstep;
    if (w < 0 || start + wsize > (int64_t)regend + wstep) return 0;
    if (!threaded) return 1;
    int num_tasks = (regend / wstep - 3) / 100;
    if (num_tasks <= 0) return 0;
    if (w / 100 <= num_tasks) return 1;        /* inside one of the tasks' own 100 windows */
    /* beyond the last task: only reached when that task's stop was clipped up to regend + wstep */
    int64_t stop = ((int64_t)num_tasks + 1) * 100 * (int64_t)wstep + (wsize - wstep);
    return stop >= regend;
}

void fpt_oracle_window_bounds(const int32_t *pos, int64_t nsnp, int64_t w, int wsize, int wstep,
                              int64_t *left, int64_t *right) {
    /* comparative.c:58-65: left = first pos >= start, right = first pos > stop */
    int64_t start = w * (int64_t)wstep, stop = start + wsize;
    int64_t lo = 0, hi = nsnp;
    while (lo < hi) { int64_t mid = (lo + hi) / 2; if (pos[mid] < start) lo = mid + 1; else hi = mid; }
    *left = lo;
    hi = nsnp;
    while (lo < hi) { int64_t mid = (lo + hi) / 2; if (pos[mid] <= stop) lo = mid + 1; else hi = mid; }
    *right = lo;
}

static int population_size(const int32_t *pos, int len) {
    /* comparative.c:25-34, but bounded by the array length (the reference runs off the end, Q9) */
    int n = 0;
    while (n < len && pos[n] == pos[0]) n++;
    return n;
}

/* ======================================================================= FET */

void fpt_oracle_fetcount(const double *avals, const double *bvals, int64_t snp, int asize, int bsize, int f[4]) {
    /* cFisher.c:208-238: row = population, column = allele (3 major, -3 minor) */
    int maj = 0, min = 0;
    for (int i = 0; i < asize; i++) {
        double v = avals[snp * asize + i];
        if (v == 3) maj++; else if (v == -3) min++;
    }
    f[0] = maj; f[1] = min;
    maj = 0; min = 0;
    for (int i = 0; i < bsize; i++) {
        double v = bvals[snp * bsize + i];
        if (v == 3) maj++; else if (v == -3) min++;
    }
    f[2] = maj; f[3] = min;
}

static uint64_t gcd64(uint64_t x, uint64_t y) {
    while (y) { uint64_t t = x % y; x = y; y = t; }
    return x;
}

uint64_t fpt_oracle_binomial(uint64_t n, uint64_t k) {
    /* cFisher.c:256-284 (multiplicative formula, gcd rescue near overflow, 0 = unavoidable overflow) */
    if (k == 0 || k == n) return 1;
    if (k == 1) return n;
    if (k > n) return 0;
    if (k > n / 2) k = n - k;
    uint64_t r = 1;
    for (uint64_t i = 1; i <= k; i++, n--) {
        if (r >= ULONG_MAX / n) {
            uint64_t g = gcd64(n, i), nr = n / g, ir = i / g;
            g = gcd64(r, ir); r /= g; ir /= g;
            if (r >= ULONG_MAX / nr) return 0;
            r = r * nr / ir;
        } else {
            r = r * n / i;
        }
    }
    return r;
}

/* cFisher.c:327-346: cells in clockwise order a,b,d,c; rotate so the first minimum leads */
static void rotate_min_first(int f[4]) {
    int cw[4] = { f[0], f[1], f[3], f[2] };
    int at = 0;
    for (int i = 1; i < 4; i++) if (cw[i] < cw[at]) at = i;
    f[0] = cw[at & 3]; f[1] = cw[(at + 1) & 3]; f[3] = cw[(at + 2) & 3]; f[2] = cw[(at + 3) & 3];
}

/* cFisher.c:357-390: the most extreme table on the other side, margins kept */
static void opposite_extreme(int f[4]) {
    int R1 = f[0] + f[1], R2 = f[2] + f[3], C1 = f[0] + f[2], C2 = f[1] + f[3];
    int m1 = R1;                                   /* first minimum of R1,R2,C1,C2 (value only) */
    if (R2 < m1) m1 = R2;
    if (C1 < m1) m1 = C1;
    if (C2 < m1) m1 = C2;
    if (R1 <= R2 && C1 <= C2)      { f[0] = m1 - f[0]; f[1] = R1 - f[0]; f[2] = C1 - f[0]; f[3] = C2 - f[1]; }
    else if (R1 <= R2 && C2 <= C1) { f[1] = m1 - f[1]; f[0] = R1 - f[1]; f[3] = C2 - f[1]; f[2] = C1 - f[0]; }
    else if (R1 >= R2 && C1 <= C2) { f[2] = m1 - f[2]; f[0] = C1 - f[2]; f[3] = R2 - f[2]; f[1] = R1 - f[0]; }
    else                           { f[3] = m1 - f[3]; f[1] = C2 - f[3]; f[2] = R2 - f[3]; f[0] = R1 - f[1]; }
}

#define FET_EXACT_MAX_N 67

/* point probability the way cFisher.c:473-483 computes it; *ok = 0 when the u64 arithmetic overflows */
static double point_prob_exact(int a, int b, int c, int d, int *ok) {
    uint64_t x = fpt_oracle_binomial((uint64_t)(a + b), (uint64_t)a);
    uint64_t y = fpt_oracle_binomial((uint64_t)(c + d), (uint64_t)c);
    uint64_t z = fpt_oracle_binomial((uint64_t)(a + b + c + d), (uint64_t)(a + c));
    if (!x || !y || !z || (unsigned __int128)x * y > (unsigned __int128)UINT64_MAX) { *ok = 0; return 0.0; }
    double nom = (double)(x * y), denom = (double)z;
    return nom / denom;
}

int fpt_oracle_fet_exact_domain(const int f[4]) {
    /* N <= 67 keeps every binomial inside u64 (Q2); the numerator product is checked on both
       tables whose point probability the reference evaluates (observed and opposite extreme). */
    int n = f[0] + f[1] + f[2] + f[3];
    if (n > FET_EXACT_MAX_N) return 0;
    int g[4] = { f[0], f[1], f[2], f[3] };
    int R1 = g[0] + g[1], R2 = g[2] + g[3], C1 = g[0] + g[2], C2 = g[1] + g[3];
    int ok = 1;
    rotate_min_first(g);
    (void)point_prob_exact(g[0], g[1], g[2], g[3], &ok);
    if (!ok) return 0;
    if (R1 == R2 || C1 == C2) return 1;
    g[1] += g[0]; g[2] += g[0]; g[3] -= g[0]; g[0] = 0;      /* end of the first tail */
    opposite_extreme(g);
    rotate_min_first(g);
    (void)point_prob_exact(g[0], g[1], g[2], g[3], &ok);
    return ok;
}

double fpt_oracle_fet_exact(const int fin[4]) {
    /* cFisher.c:405-455, same operation order (ratio first, then times the running term) */
    int f[4] = { fin[0], fin[1], fin[2], fin[3] };
    int R1 = f[0] + f[1], R2 = f[2] + f[3], C1 = f[0] + f[2], C2 = f[1] + f[3];
    int ok = 1;
    rotate_min_first(f);
    double P0 = point_prob_exact(f[0], f[1], f[2], f[3], &ok);
    double P = P0, P1 = P0;
    while (f[0] > 0) {
        f[1]++; f[2]++;
        P1 = (1.0 * f[0] * f[3]) / (double)(f[1] * f[2]) * P1;
        P += P1;
        f[0]--; f[3]--;
    }
    if (R1 == R2 || C1 == C2) {
        P = 2 * P;
    } else {
        opposite_extreme(f);
        rotate_min_first(f);
        double P2 = point_prob_exact(f[0], f[1], f[2], f[3], &ok);
        while (P2 < P0) {
            P += P2;
            if (f[1] == 0 || f[2] == 0) break;
            f[0]++; f[3]++;
            P2 = (1.0 * f[1] * f[2]) / (double)(f[0] * f[3]) * P2;
            f[1]--; f[2]--;
        }
    }
    if (P > 1) P = 1;
    return ok ? P : NAN;
}

/* ---- log mode: the same walk with every term expressed relative to the observed table's P0.
 * lf(k) = lgamma(k+1). Terms below 2^-60 of the running sum cannot change it any more (the ratios
 * only shrink along a tail), so the first tail stops there; the second tail starts at the first
 * table whose probability is within 2^-60 of P0 (found by bisection on the concave log-pmf) and
 * includes a table while its relative weight is < 1 - 1e-10 (strict `P2 < P0` with a tie guard).
 */
#define FET_LOG_SKIP   41.58883083359672     /* 60 ln 2 */
#define FET_TIE_GUARD  1e-10
#define FET_TINY       0x1p-60

static double lfact(int k) { return lgamma((double)k + 1.0); }

static double log_point_prob(int a, int b, int c, int d) {
    return ((((lfact(a + b) + lfact(c + d)) + lfact(a + c)) + lfact(b + d)) - lfact(a + b + c + d))
           - (((lfact(a) + lfact(b)) + lfact(c)) + lfact(d));
}

double fpt_oracle_fet_neglog10_logmode(const int fin[4]) {
    int f[4] = { fin[0], fin[1], fin[2], fin[3] };
    int R1 = f[0] + f[1], R2 = f[2] + f[3], C1 = f[0] + f[2], C2 = f[1] + f[3];
    rotate_min_first(f);
    double lp0 = log_point_prob(f[0], f[1], f[2], f[3]);
    double S = 1.0, u = 1.0;
    while (f[0] > 0) {
        f[1]++; f[2]++;
        u = ((double)f[0] * (double)f[3]) / ((double)f[1] * (double)f[2]) * u;
        S += u;
        f[0]--; f[3]--;
        if (u < S * FET_TINY) {                    /* the rest of the tail is a no-op */
            f[1] += f[0]; f[2] += f[0]; f[3] -= f[0]; f[0] = 0;
            break;
        }
    }
    if (R1 == R2 || C1 == C2) {
        S = 2 * S;
    } else {
        opposite_extreme(f);
        rotate_min_first(f);
        int K = f[1] < f[2] ? f[1] : f[2];         /* steps available before b or c runs out */
        int n = f[0] + f[1] + f[2] + f[3];
        /* inward steps up to the mode of cell a: floor((a+c+1)(a+b+1)/(n+2)) - a */
        int64_t mode = ((int64_t)(f[0] + f[2] + 1) * (int64_t)(f[0] + f[1] + 1)) / (int64_t)(n + 2);
        int khi = (int)(mode - f[0]);
        if (khi < 0) khi = 0;
        if (khi > K) khi = K;
        int k = 0;
        double lu = log_point_prob(f[0], f[1], f[2], f[3]) - lp0;
        if (lu < -FET_LOG_SKIP) {
            int lo = 0, hi = khi;                  /* first k with log weight >= -FET_LOG_SKIP */
            while (lo < hi) {
                int mid = (lo + hi) / 2;
                double l = log_point_prob(f[0] + mid, f[1] - mid, f[2] - mid, f[3] + mid) - lp0;
                if (l < -FET_LOG_SKIP) lo = mid + 1; else hi = mid;
            }
            k = lo;
            f[0] += k; f[1] -= k; f[2] -= k; f[3] += k;
            lu = log_point_prob(f[0], f[1], f[2], f[3]) - lp0;
        }
        double u2 = exp(lu);
        while (u2 < 1.0 - FET_TIE_GUARD) {
            S += u2;
            if (f[1] == 0 || f[2] == 0) break;
            f[0]++; f[3]++;
            u2 = ((double)f[1] * (double)f[2]) / ((double)f[0] * (double)f[3]) * u2;
            f[1]--; f[2]--;
        }
    }
    double lp = lp0 + log(S);
    if (lp >= 0.0) return -0.0;                    /* P clamps to 1 and -1.0*log10(1) is -0.0 (Q4) */
    return -(lp * 0.43429448190325182765);         /* log10(e) */
}

double fpt_oracle_fet_neglog10(const int f[4]) {
    /* cFisher.c:180-184: fetscores[i] = -1.0*log10(fet(f)) */
    if (fpt_oracle_fet_exact_domain(f)) return -1.0 * log10(fpt_oracle_fet_exact(f));
    return fpt_oracle_fet_neglog10_logmode(f);
}

void fpt_oracle_fet_tables(const int32_t *tables, int64_t n, double *neglog10p) {
    for (int64_t i = 0; i < n; i++) {
        int f[4] = { tables[4 * i], tables[4 * i + 1], tables[4 * i + 2], tables[4 * i + 3] };
        neglog10p[i] = fpt_oracle_fet_neglog10(f);
    }
}

int fpt_oracle_fet_per_snp(const double *avals, const double *bvals, int64_t nsnp, int asize, int bsize,
                           int32_t *tables, double *neglog10p) {
    for (int64_t k = 0; k < nsnp; k++) {
        int f[4];
        fpt_oracle_fetcount(avals, bvals, k, asize, bsize, f);
        if (tables) for (int j = 0; j < 4; j++) tables[4 * k + j] = f[j];
        if (neglog10p) neglog10p[k] = fpt_oracle_fet_neglog10(f);
    }
    return 0;
}

static int cmp_double(const void *pa, const void *pb) {
    double a = *(const double *)pa, b = *(const double *)pb;
    return (a > b) - (a < b);
}

static double percentile_of_sorted(const double *x, int n, double q) {
    /* cFisher.c:141-143. The reference reads x[idx+1] even when it lies one past the data (n == 1 or
       q == 1, Q8) and multiplies it by delta == 0; here that term is dropped instead of read. */
    int idx = (int)((n - 1) * q);
    double delta = (n - 1) * q - idx;
    if (idx + 1 >= n) return (1 - delta) * x[idx];
    return (1 - delta) * x[idx] + delta * x[idx + 1];
}

double fpt_oracle_percentile(double *x, int n, double q) {
    qsort(x, (size_t)n, sizeof(double), cmp_double);
    return percentile_of_sorted(x, n, q);
}

void fpt_oracle_fet_window(double *snp_scores, int npos, double perc, int nsamples, uint64_t state, double out[2]) {
    /* cFisher.c:188-194 + 562-597: percentile of the window, then sigma of `nsamples` bootstrap
       percentiles; resampling draws from the already-sorted scores; std/mean sum from the top index
       down (cFisher.c:492-518). */
    out[0] = fpt_oracle_percentile(snp_scores, npos, perc);
    double *sample = (double *)malloc((size_t)npos * sizeof(double));
    double *reps = (double *)malloc((size_t)nsamples * sizeof(double));
    for (int s = 0; s < nsamples; s++) {
        for (int i = npos; i--;) sample[i] = snp_scores[fpt_oracle_randint(npos, &state)];
        reps[s] = fpt_oracle_percentile(sample, npos, perc);
    }
    double mu = 0;
    for (int i = nsamples; i--;) mu += reps[i];
    mu /= nsamples;
    double var = 0;
    for (int i = nsamples; i--;) var += (reps[i] - mu) * (reps[i] - mu);
    out[1] = sqrt(var / nsamples);
    free(sample); free(reps);
}

int fpt_oracle_fet_scan(const double *avals, const double *bvals, const int32_t *apos, const int32_t *bpos,
                        int regstart, int regend, int wsize, int wstep, int alen, int blen, double perc,
                        double *scores, double *stddev, int threaded, uint64_t seed) {
    (void)regstart;                                              /* ignored by the reference too (Q5) */
    int asize = population_size(apos, alen), bsize = population_size(bpos, blen);
    if (asize <= 0 || bsize <= 0) return -1;
    int64_t nsnp = alen / asize;
    if (blen / bsize != nsnp) return -2;
    int32_t *pos = (int32_t *)malloc((size_t)(nsnp ? nsnp : 1) * sizeof(int32_t));
    for (int64_t k = 0; k < nsnp; k++) {
        pos[k] = apos[k * asize];
        if (bpos[k * bsize] != pos[k]) { free(pos); return -2; }
    }
    double *snp = (double *)malloc((size_t)(nsnp ? nsnp : 1) * sizeof(double));
    fpt_oracle_fet_per_snp(avals, bvals, nsnp, asize, bsize, NULL, snp);
    int64_t nwin = fpt_oracle_window_count(regend, wsize, wstep), nout = regend / wstep;
    double *buf = NULL; int64_t cap = 0;
    for (int64_t w = 0; w < nwin && w < nout; w++) {
        if (!fpt_oracle_window_scheduled(w, regend, wsize, wstep, threaded)) continue;
        int64_t l, r;
        fpt_oracle_window_bounds(pos, nsnp, w, wsize, wstep, &l, &r);
        int npos = (int)(r - l);
        if (npos <= 0) continue;
        if (npos > cap) { cap = npos; buf = (double *)realloc(buf, (size_t)cap * sizeof(double)); }
        memcpy(buf, snp + l, (size_t)npos * sizeof(double));
        double out[2];
        fpt_oracle_fet_window(buf, npos, perc, 100, fpt_oracle_window_state(seed, w, FPT_STREAM_RESAMPLE), out);
        scores[w] = out[0];
        stddev[w] = out[1];
    }
    free(buf); free(snp); free(pos);
    return 0;
}

/* ======================================================================= CSS */

void fpt_oracle_compare_all(const double *avals, const double *bvals, int asize, int bsize, int npos, double *D) {
    /* css.c:277-327: D[i][j] = number of SNPs where i and j are opposite homozygotes (3 vs -3);
       diagonal left untouched; individuals 0..asize-1 are group A, asize..m-1 group B */
    int m = asize + bsize;
    for (int i = 0; i < m; i++) {
        for (int j = 0; j < i; j++) {
            const double *vi = i < asize ? avals + i : bvals + (i - asize);
            const double *vj = j < asize ? avals + j : bvals + (j - asize);
            int si = i < asize ? asize : bsize, sj = j < asize ? asize : bsize;
            double count = 0;
            for (int k = 0; k < npos; k++) {
                double x = vi[(size_t)k * si], y = vj[(size_t)k * sj];
                if ((x == 3 && y == -3) || (x == -3 && y == 3)) count++;
            }
            D[i * m + j] = count;
            D[j * m + i] = count;
        }
    }
}

void fpt_oracle_compare_freq(const double *avals, const double *bvals, int npos, double *D) {
    /* css.c:245-264: m = 2, mean absolute minor-allele-frequency difference, summed from the top down */
    double s = 0;
    for (int i = npos; i--;) s += fabs(avals[i] - bvals[i]);
    s /= npos;
    D[1] = s; D[2] = s;
}

int fpt_oracle_fill_averages(double *D, int m) {
    /* css.c:337-366: blanks (< 1e-5, diagonal included) get sum/m^2; discard when blanks > m*m/2 */
    int blanks = 0, total = m * m;
    double sum = 0;
    for (int i = m; i--;)
        for (int j = m; j--;) {
            if (D[i * m + j] < 0.00001) blanks++; else sum += D[i * m + j];
        }
    double avg = sum / total;
    if (blanks > total / 2) return 0;
    for (int i = 0; i < total; i++) if (D[i] < 0.00001) D[i] = avg;
    return 1;
}

/* cyclic Jacobi on a dense symmetric matrix; V's columns are the eigenvectors */
static void jacobi_eig(double *A, double *V, int n) {
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) V[i * n + j] = i == j;
    for (int sweep = 0; sweep < 100; sweep++) {
        double off = 0, dia = 0;
        for (int i = 0; i < n; i++) {
            dia += A[i * n + i] * A[i * n + i];
            for (int j = i + 1; j < n; j++) off += A[i * n + j] * A[i * n + j];
        }
        /* quadratic convergence; the rounding floor of off/dia sits near 1e-29, so 1e-26 is reached one sweep after 1e-13 */
        if (off <= 1e-300 || off <= 1e-26 * (dia + off)) break;
        for (int p = 0; p < n - 1; p++)
            for (int q = p + 1; q < n; q++) {
                double apq = A[p * n + q];
                if (apq == 0.0) continue;
                double th = (A[q * n + q] - A[p * n + p]) / (2 * apq);
                double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1));
                double c = 1 / sqrt(t * t + 1), s = t * c;
                for (int k = 0; k < n; k++) {
                    double x = A[k * n + p], y = A[k * n + q];
                    A[k * n + p] = c * x - s * y; A[k * n + q] = s * x + c * y;
                }
                for (int k = 0; k < n; k++) {
                    double x = A[p * n + k], y = A[q * n + k];
                    A[p * n + k] = c * x - s * y; A[q * n + k] = s * x + c * y;
                }
                for (int k = 0; k < n; k++) {
                    double x = V[k * n + p], y = V[k * n + q];
                    V[k * n + p] = c * x - s * y; V[k * n + q] = s * x + c * y;
                }
            }
    }
}


/* Householder tridiagonalisation + implicit-shift QL for cohorts where cyclic Jacobi would take minutes (m > 128).
   Same contract as jacobi_eig: eigenvalues on A's diagonal, eigenvectors in V's columns. The rotations are applied to the
   transposed eigenvector matrix so that the inner loops run over contiguous memory. */
static void tridiag_ql_eig(double *A, double *V, int n) {
    double *d = (double *)calloc((size_t)n + 1, sizeof(double)), *e = (double *)calloc((size_t)n + 1, sizeof(double));
    double *Q = (double *)malloc((size_t)n * n * sizeof(double));   /* rows = basis vectors (Q^T) */
    double *u = (double *)malloc((size_t)n * sizeof(double)), *w = (double *)malloc((size_t)n * sizeof(double));
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) Q[(size_t)i * n + j] = i == j;
    /* reduce from the top: step k annihilates A[k+2.., k] with H = I - beta u u' acting on indices k+1..n-1 */
    for (int k = 0; k + 2 < n; k++) {
        double sigma = 0;
        for (int i = k + 2; i < n; i++) sigma += A[(size_t)i * n + k] * A[(size_t)i * n + k];
        double x0 = A[(size_t)(k + 1) * n + k];
        if (sigma == 0.0) continue;
        double norm = sqrt(x0 * x0 + sigma), alpha = x0 > 0 ? -norm : norm;
        u[k + 1] = x0 - alpha;
        for (int i = k + 2; i < n; i++) u[i] = A[(size_t)i * n + k];
        double beta = 1.0 / (norm * (norm + fabs(x0)));           /* 2 / u'u */
        /* w = beta A u ; K = beta/2 u'w ; w -= K u ; A -= u w' + w u' (trailing block) */
        for (int i = k + 1; i < n; i++) {
            double acc = 0; const double *row = A + (size_t)i * n;
            for (int j = k + 1; j < n; j++) acc += row[j] * u[j];
            w[i] = beta * acc;
        }
        double kk = 0;
        for (int i = k + 1; i < n; i++) kk += u[i] * w[i];
        kk *= 0.5 * beta;
        for (int i = k + 1; i < n; i++) w[i] -= kk * u[i];
        for (int i = k + 1; i < n; i++) {
            double *row = A + (size_t)i * n; const double ui = u[i], wi = w[i];
            for (int j = k + 1; j < n; j++) row[j] -= ui * w[j] + wi * u[j];
        }
        A[(size_t)(k + 1) * n + k] = A[(size_t)k * n + k + 1] = alpha;
        for (int i = k + 2; i < n; i++) A[(size_t)i * n + k] = A[(size_t)k * n + i] = 0.0;
        /* Q^T <- Q^T H: every row r of Q gets r -= beta (r.u) u */
        for (int r = 0; r < n; r++) {
            double *row = Q + (size_t)r * n; double acc = 0;
            for (int j = k + 1; j < n; j++) acc += row[j] * u[j];
            acc *= beta;
            for (int j = k + 1; j < n; j++) row[j] -= acc * u[j];
        }
    }
    /* Q currently holds the product H1 H2 ... as a matrix whose COLUMNS are the tridiagonal basis; transpose into rows */
    for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) {
        double t = Q[(size_t)i * n + j]; Q[(size_t)i * n + j] = Q[(size_t)j * n + i]; Q[(size_t)j * n + i] = t;
    }
    for (int i = 0; i < n; i++) { d[i] = A[(size_t)i * n + i]; e[i] = i + 1 < n ? A[(size_t)(i + 1) * n + i] : 0.0; }
    /* implicit QL; rotation (i, i+1) mixes rows i and i+1 of Q */
    for (int l = 0; l < n; l++) {
        for (int iter = 0; iter < 300; iter++) {
            int mm = l;
            for (; mm + 1 < n; mm++) {
                double dd = fabs(d[mm]) + fabs(d[mm + 1]);
                if (fabs(e[mm]) <= 2.220446049250313e-16 * dd) break;
            }
            if (mm == l) break;
            double g = (d[l + 1] - d[l]) / (2.0 * e[l]), r = hypot(g, 1.0);
            g = d[mm] - d[l] + e[l] / (g + (g >= 0 ? fabs(r) : -fabs(r)));
            double s = 1.0, c = 1.0, p = 0.0; int i;
            for (i = mm - 1; i >= l; i--) {
                double f = s * e[i], b = c * e[i];
                r = hypot(f, g); e[i + 1] = r;
                if (r == 0.0) { d[i + 1] -= p; e[mm] = 0.0; break; }
                s = f / r; c = g / r; g = d[i + 1] - p;
                r = (d[i] - g) * s + 2.0 * c * b; p = s * r; d[i + 1] = g + p; g = c * r - b;
                double *r0 = Q + (size_t)i * n, *r1 = Q + (size_t)(i + 1) * n;
                for (int k = 0; k < n; k++) { double f1 = r1[k]; r1[k] = s * r0[k] + c * f1; r0[k] = c * r0[k] - s * f1; }
            }
            if (r == 0.0 && i >= l) continue;
            d[l] -= p; e[l] = g; e[mm] = 0.0;
        }
    }
    for (int i = 0; i < n; i++) {
        A[(size_t)i * n + i] = d[i];
        for (int j = 0; j < n; j++) V[(size_t)j * n + i] = Q[(size_t)i * n + j];
    }
    free(d); free(e); free(Q); free(u); free(w);
}

void fpt_oracle_cmds(const double *D, int m, double *X, double evals[3]) {
    /* css.c:505-560: B = -1/2 Z (D.D) Z with Z = I - 11'/m, formed by the same two products
       (Z*(D.D) then *Z); the two largest eigenvalues by value; X = Q sqrt(L), no guard on L < 0 */
    size_t mm = (size_t)m * m;
    double *S = (double *)malloc(mm * sizeof(double)), *T = (double *)malloc(mm * sizeof(double));
    double *B = (double *)malloc(mm * sizeof(double)), *V = (double *)malloc(mm * sizeof(double));
    double zo = -1.0 / m, zd = (m - 1) / (m * 1.0);
    for (size_t i = 0; i < mm; i++) S[i] = D[i] * D[i];
    /* k-outer accumulation: every T[i][j] still sums k = 0..m-1 in order (bit-identical), the loads are contiguous */
    for (int i = 0; i < m; i++) {
        double *Ti = T + (size_t)i * m;
        for (int j = 0; j < m; j++) Ti[j] = 0;
        for (int k = 0; k < m; k++) {
            const double z = i == k ? zd : zo; const double *Sk = S + (size_t)k * m;
            for (int j = 0; j < m; j++) Ti[j] += z * Sk[j];
        }
    }
    for (int i = 0; i < m; i++)
        for (int j = 0; j < m; j++) {
            double acc = 0;
            for (int k = 0; k < m; k++) acc += T[i * m + k] * (k == j ? zd : zo);
            B[i * m + j] = acc * -0.5;
        }
    if (m > 128) tridiag_ql_eig(B, V, m); else jacobi_eig(B, V, m);
    int i1 = -1, i2 = -1, i3 = -1;                  /* top three by value */
    for (int i = 0; i < m; i++) {
        double e = B[i * m + i];
        if (i1 < 0 || e > B[i1 * m + i1]) { i3 = i2; i2 = i1; i1 = i; }
        else if (i2 < 0 || e > B[i2 * m + i2]) { i3 = i2; i2 = i; }
        else if (i3 < 0 || e > B[i3 * m + i3]) { i3 = i; }
    }
    double l1 = B[i1 * m + i1], l2 = i2 >= 0 ? B[i2 * m + i2] : 0.0, l3 = i3 >= 0 ? B[i3 * m + i3] : 0.0;
    /* css.c:558 takes these square roots unguarded (negative eigenvalue -> NaN coordinates). Deviation, shared with the
       product: an eigenvalue that is zero up to rounding (rank-deficient B) is clamped to zero instead of letting the
       sign of its last bit decide between 0 and NaN. */
    if (l2 < 0 && -l2 <= 1e-13 * fabs(l1)) l2 = 0.0;
    double s1 = sqrt(l1), s2 = sqrt(l2);
    for (int j = 0; j < m; j++) {
        X[2 * j] = V[j * m + i1] * s1;
        X[2 * j + 1] = i2 >= 0 ? V[j * m + i2] * s2 : 0.0;
    }
    if (evals) { evals[0] = l1; evals[1] = l2; evals[2] = l3; }
    free(S); free(T); free(B); free(V);
}

void fpt_oracle_calc_dist(const double *X, int m, double *dist) {
    /* css.c:573-587 */
    for (int i = m; i--;) {
        dist[i * m + i] = 0;
        for (int j = i; j--;) {
            double dx = X[2 * i] - X[2 * j], dy = X[2 * i + 1] - X[2 * j + 1];
            double d = sqrt(dx * dx + dy * dy);
            dist[i * m + j] = d; dist[j * m + i] = d;
        }
    }
}

static double stress_of(const double *delta, const double *dist, int m) {
    /* css.c:767-777: sum over i > j, both indices counting down */
    double s = 0;
    for (int i = m; i--;)
        for (int j = i; j--;) {
            double e = dist[i * m + j] - delta[i * m + j];
            s += e * e;
        }
    return s;
}

static double smacof_core(const double *delta, int m, double *X, int max_iters, double eps, int *iters, double *min_margin);

double fpt_oracle_smacof(const double *delta, int m, double *X, int max_iters, double eps, int *iters) {
    return smacof_core(delta, m, X, max_iters, eps, iters, NULL);
}

/* the same, also reporting how close the stopping rule came to deciding the other way: the smallest
   |(sigma_prev - sigma) - eps| over the decisions taken (test diagnostics: a start that differs in its last bits
   can only change the iteration count of a window whose margin is of the size of the induced stress change) */
double fpt_oracle_smacof_margin(const double *delta, int m, double *X, int max_iters, double eps, int *iters, double *min_margin) {
    return smacof_core(delta, m, X, max_iters, eps, iters, min_margin);
}

static double smacof_core(const double *delta, int m, double *X, int max_iters, double eps, int *iters, double *min_margin) {
    /* css.c:907-938 with guttman_transform css.c:811-836: b_ij = -delta_ij/d_ij (0 when d_ij < 1e-5),
       b_ii = -sum_j b_ij accumulated with j counting down, X <- (B Z)/m with the product summed
       over ascending k (dgemm), loop while first pass or (drop > eps and k <= max_iters) */
    size_t mm = (size_t)m * m;
    double *dist = (double *)malloc(mm * sizeof(double)), *B = (double *)malloc(mm * sizeof(double));
    double *Z = (double *)malloc((size_t)m * 2 * sizeof(double));
    memcpy(Z, X, (size_t)m * 2 * sizeof(double));
    fpt_oracle_calc_dist(X, m, dist);
    double sigma = stress_of(delta, dist, m), prev = 0;
    int k = 0;
    double margin = INFINITY;
    for (;;) {
        if (k > 0) {
            if (k <= max_iters && fabs((prev - sigma) - eps) < margin) margin = fabs((prev - sigma) - eps);
            if (!((prev - sigma) > eps && k <= max_iters)) break;
        }
        prev = sigma;
        k++;
        for (int i = m; i--;) {
            double d = 0;
            for (int j = m; j--;) {
                if (i == j) continue;
                double b = dist[i * m + j] < 0.00001 ? 0.0 : -1 * delta[i * m + j] / dist[i * m + j];
                B[i * m + j] = b;
                d += b;
            }
            B[i * m + i] = -1 * d;
        }
        for (int i = 0; i < m; i++) {
            double x = 0, y = 0;
            for (int j = 0; j < m; j++) { x += B[i * m + j] * Z[2 * j]; y += B[i * m + j] * Z[2 * j + 1]; }
            X[2 * i] = x / m; X[2 * i + 1] = y / m;
        }
        fpt_oracle_calc_dist(X, m, dist);
        sigma = stress_of(delta, dist, m);
        memcpy(Z, X, (size_t)m * 2 * sizeof(double));
    }
    if (iters) *iters = k;
    if (min_margin) *min_margin = margin;
    free(dist); free(B); free(Z);
    return sigma;
}

double fpt_oracle_smacof_runs(const double *delta, int m, double *X, int max_iters, int n_init, double eps,
                              uint64_t *state) {
    /* css.c:852-884: n_init uniform(0,1) starts (x then y per individual), keep the lowest stress.
       The reference only accepts a run below its 99999 sentinel and otherwise leaves stale
       coordinates in X (Q12); here the first run is always accepted. */
    double *cand = (double *)malloc((size_t)m * 2 * sizeof(double));
    double best = INFINITY;
    for (int r = 0; r < n_init; r++) {
        for (int i = 0; i < m; i++) { cand[2 * i] = fpt_oracle_drand48(state); cand[2 * i + 1] = fpt_oracle_drand48(state); }
        double s = fpt_oracle_smacof(delta, m, cand, max_iters, eps, NULL);
        if (r == 0 || s < best) { memcpy(X, cand, (size_t)m * 2 * sizeof(double)); best = s; }
    }
    free(cand);
    return best;
}

double fpt_oracle_css(const double *dist, int m, const int *at, const int *bt, int asize, int bsize) {
    /* css.c:608-647: mean between-group distance minus (asize+bsize) times the within-group terms,
       which only visit ADJACENT pairs in track order; all sums count down; integer divisors */
    double bet = 0;
    for (int i = asize; i--;)
        for (int j = bsize; j--;) bet += dist[at[i] * m + bt[j]];
    bet = bet / (asize * bsize);
    double wa = 0, wb = 0;
    if (asize > 1) {
        for (int i = asize - 1; i--;) wa += dist[at[i] * m + at[i + 1]];
        wa = wa / (asize * asize * (asize - 1));
    }
    if (bsize > 1) {
        for (int i = bsize - 1; i--;) wb += dist[bt[i] * m + bt[i + 1]];
        wb = wb / (bsize * bsize * (bsize - 1));
    }
    return bet - (asize + bsize) * (wa + wb);
}

double fpt_oracle_significance(const double *dist, int m, int *tracks, int asize, int bsize, double score,
                               int treshold, int runs, uint64_t *state, int *hits_out, int *n_out) {
    /* css.c:727-752 with the Fisher-Yates shuffle of css.c:700-706 applied to the persistent array */
    int hits = 0, n = 0;
    while (hits < treshold && n < runs) {
        for (int i = m - 1; i > 0; i--) {
            int r = (int)fpt_oracle_randint(i + 1, state);
            int t = tracks[i]; tracks[i] = tracks[r]; tracks[r] = t;
        }
        if (fpt_oracle_css(dist, m, tracks, tracks + asize, asize, bsize) >= score) hits++;
        n++;
    }
    if (hits_out) *hits_out = hits;
    if (n_out) *n_out = n;
    return (hits + 1) * 1.0 / (n + 1);
}

/* state after n further draws of the LCG (square-and-multiply on the affine map) */
uint64_t fpt_oracle_lcg_skip(uint64_t s, uint64_t n) {
    uint64_t ra = 1, rc = 0, ba = LCG_A, bc = LCG_C;
    while (n) {
        if (n & 1) { rc = (ba * rc + bc) & MASK48; ra = (ba * ra) & MASK48; }
        bc = (ba * bc + bc) & MASK48;
        ba = (ba * ba) & MASK48;
        n >>= 1;
    }
    return (ra * s + rc) & MASK48;
}

/* permutation test in "independent" mode (the product's default): permutation k is the reference's Fisher-Yates
   shuffle (css.c:700-706) of FRESH identity labels with the window's nrand48 stream positioned k*(m-1) draws in;
   scoring, hit counting, early stop and p exactly as css.c:727-752 */
double fpt_oracle_significance_indep(const double *dist, int m, int asize, int bsize, double score, int treshold,
                                     int runs, uint64_t state, int *hits_out, int *n_out) {
    int hits = 0, n = 0;
    int *tracks = (int *)malloc((size_t)m * sizeof(int));
    while (hits < treshold && n < runs) {
        uint64_t st = fpt_oracle_lcg_skip(state, (uint64_t)n * (uint64_t)(m - 1));
        for (int i = 0; i < m; i++) tracks[i] = i;
        for (int i = m - 1; i > 0; i--) {
            int r = (int)fpt_oracle_randint(i + 1, &st);
            int t = tracks[i]; tracks[i] = tracks[r]; tracks[r] = t;
        }
        if (fpt_oracle_css(dist, m, tracks, tracks + asize, asize, bsize) >= score) hits++;
        n++;
    }
    free(tracks);
    if (hits_out) *hits_out = hits;
    if (n_out) *n_out = n;
    return (hits + 1) * 1.0 / (n + 1);
}

static int g_perm_chain = 0;    /* 0 = independent shuffles (product default), 1 = the reference's chained label array */
void fpt_oracle_set_perm_mode(int chain) { g_perm_chain = chain != 0; }

double fpt_oracle_css_window(const double *avals, const double *bvals, int asize, int bsize, int npos,
                             int drosophila, int mds, int treshold, int runs, uint64_t state_perm,
                             uint64_t state_init, double *p_out, double *X_out, double evals_out[3]) {
    /* css.c:181-223 then threadcss.c:264-270 */
    int m = asize + bsize;
    size_t mm = (size_t)m * m;
    double *D = (double *)calloc(mm, sizeof(double)), *dist = (double *)malloc(mm * sizeof(double));
    double *X = (double *)calloc((size_t)m * 2, sizeof(double));
    int *tracks = (int *)malloc((size_t)m * sizeof(int));
    double score = -1;
    if (drosophila) fpt_oracle_compare_freq(avals, bvals, npos, D);
    else fpt_oracle_compare_all(avals, bvals, asize, bsize, npos, D);
    if (fpt_oracle_fill_averages(D, m)) {
        double ev[3] = { 0, 0, 0 };
        if (mds == 0) fpt_oracle_cmds(D, m, X, ev);
        else if (mds == 1) fpt_oracle_smacof_runs(D, m, X, 300, 4, 0.000001, &state_init);
        else { fpt_oracle_cmds(D, m, X, ev); fpt_oracle_smacof(D, m, X, 300, 0.000001, NULL); }
        fpt_oracle_calc_dist(X, m, dist);
        for (int i = 0; i < m; i++) tracks[i] = i;
        score = fpt_oracle_css(dist, m, tracks, tracks + asize, asize, bsize);
        if (p_out) *p_out = g_perm_chain
            ? fpt_oracle_significance(dist, m, tracks, asize, bsize, score, treshold, runs, &state_perm, NULL, NULL)
            : fpt_oracle_significance_indep(dist, m, asize, bsize, score, treshold, runs, state_perm, NULL, NULL);
        if (X_out) memcpy(X_out, X, (size_t)m * 2 * sizeof(double));
        if (evals_out) memcpy(evals_out, ev, sizeof ev);
    }
    free(D); free(dist); free(X); free(tracks);
    return score;
}

int fpt_oracle_css_scan(const double *avals, const double *bvals, const int32_t *apos, const int32_t *bpos,
                        int regstart, int regend, int wsize, int wstep, int alen, int blen, int treshold,
                        int runs, int drosophila, int mds, double *scores, double *p, int threaded,
                        uint64_t seed) {
    (void)regstart;
    int asize = population_size(apos, alen), bsize = population_size(bpos, blen);
    if (asize <= 0 || bsize <= 0) return -1;
    int64_t nsnp = alen / asize;
    if (blen / bsize != nsnp) return -2;
    int32_t *pos = (int32_t *)malloc((size_t)(nsnp ? nsnp : 1) * sizeof(int32_t));
    for (int64_t k = 0; k < nsnp; k++) {
        pos[k] = apos[k * asize];
        if (bpos[k * bsize] != pos[k]) { free(pos); return -2; }
    }
    int64_t nwin = fpt_oracle_window_count(regend, wsize, wstep), nout = regend / wstep;
    for (int64_t w = 0; w < nwin && w < nout; w++) {
        if (!fpt_oracle_window_scheduled(w, regend, wsize, wstep, threaded)) continue;
        int64_t l, r;
        fpt_oracle_window_bounds(pos, nsnp, w, wsize, wstep, &l, &r);
        int npos = (int)(r - l);
        if (npos <= 0) continue;
        double pv = 0;
        double s = fpt_oracle_css_window(avals + l * asize, bvals + l * bsize, asize, bsize, npos, drosophila, mds,
                                         treshold, runs, fpt_oracle_window_state(seed, w, FPT_STREAM_RESAMPLE),
                                         fpt_oracle_window_state(seed, w, FPT_STREAM_INIT), &pv, NULL, NULL);
        if (s != -1) { scores[w] = s; p[w] = pv; }
    }
    free(pos);
    return 0;
}
