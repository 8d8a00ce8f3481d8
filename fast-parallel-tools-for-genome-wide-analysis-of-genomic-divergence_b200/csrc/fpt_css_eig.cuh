/*
 * fpt_css_eig.cuh — classical MDS of one window by ONE WARP (reference: cmds, css/css.c:505-560, which
 * calls GSL's dense symmetric eigensolver and keeps the two largest eigenpairs).
 *
 * The reference needs only the two largest eigenvalues BY VALUE and their vectors, so a full
 * diagonalisation is two orders of magnitude more work than necessary. Per window this file does:
 *
 * Two kernels (the first needs the m x m matrix in shared memory, the second only O(m) vectors and therefore runs
 * with twice the resident warps — it is a chain of dependent fp64 operations — and each fits the instruction cache):
 *
 *   1. pairwise opposite-homozygote counts from the window's bit-plane slab, fill_averages,
 *      double centring  B = -1/2 (S - r 1' - 1 r' + g),  S = D.D                     (O(m^2))
 *   2. Householder reduction of B to tridiagonal T (LAPACK dsytd2 scheme, lower)       (4/3 m^3 flops)
 *   3. the two (optionally three) largest eigenvalues of T by Sturm-count multisection, both searched
 *      at once, 16 lanes each                                                          (O(m) per probe)
 *   4. their eigenvectors by inverse iteration on T (tridiagonal LU with partial pivoting, lanes 0/1)
 *   5. back-transformation through the stored reflectors, X = [v1 v2] diag(sqrt l1, sqrt l2)
 *
 * No __syncthreads anywhere: each warp of the CTA owns its own window and its own slice of shared
 * memory. The symmetric matrix is kept as its packed lower triangle (half the shared memory of a square
 * array, hence more resident warps for a latency-bound kernel, and only one triangle to update).
 */
#ifndef FPT_CSS_EIG_CUH
#define FPT_CSS_EIG_CUH

#include "fpt_rt.cuh"

FPT_D double fpt_warp_sum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FPT_FULL_MASK, v, o);
    return v;
}
FPT_D long long fpt_warp_sum_i64(long long v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FPT_FULL_MASK, v, o);
    return v;
}
FPT_D double fpt_warp_max(double v) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FPT_FULL_MASK, v, o));
    return v;
}
FPT_D double fpt_warp_min(double v) {
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FPT_FULL_MASK, v, o));
    return v;
}

/* per-warp shared-memory work area */
struct FptEigWork {
    double *A;        /* symmetric matrix, lower triangle packed by rows: element (i, j <= i) at i(i+1)/2 + j */
    double *d, *e;    /* tridiagonal: m diagonal, m-1 off-diagonal (e2 overwrites nothing: kept in tau2) */
    double *tau;      /* m reflector scales */
    double *pv, *wv;  /* m each: Householder work vectors */
    double *y;        /* 2 x m eigenvectors */
    double *lu;       /* 2 x 3 x m tridiagonal solver arrays */
    unsigned *wbuf;   /* wch x 2 x m bit-plane words */
    int wch;
};

FPT_HD int fpt_tri(int i) { return (i * (i + 1)) >> 1; }

/* e / m for 0 <= e < m*m, m <= 4096, without an integer division: magic = ceil(2^32 / m) */
FPT_HD unsigned fpt_div_magic(int m) { return (unsigned)((0x100000000ULL + (unsigned)m - 1) / (unsigned)m); }
FPT_D int fpt_fastdiv(int e, unsigned magic) { return (int)__umulhi((unsigned)e, magic); }

/* 1/q to full double precision from the single-precision reciprocal and two Newton steps (|q| within float range) */
FPT_D double fpt_fast_rcp(double q) {
    double r = (double)(1.0f / (float)q);
    r = fma(r, fma(-q, r, 1.0), r);
    r = fma(r, fma(-q, r, 1.0), r);
    return r;
}

/* ---- step 1a: D (packed lower triangle, zero diagonal) from the bit-planes of SNPs [l, r); warp-level twin of
   fpt_css_counts. Lane = column j of row i, so the plane words of a row are read as consecutive addresses. */
FPT_D void fpt_warp_counts(const unsigned *__restrict__ planes, int m, int l, int r, const FptEigWork &w) {
    const int lane = threadIdx.x & 31;
    const unsigned magic = fpt_div_magic(m);
    const int w0 = l >> 5, w1 = (r - 1) >> 5;
    #pragma unroll 1
    for (int e = lane; e < fpt_tri(m); e += 32) w.A[e] = 0.0;
    #pragma unroll 1
    for (int wc = w0; wc <= w1; wc += w.wch) {
        const int nw = min(w.wch, w1 - wc + 1);
        __syncwarp();
        #pragma unroll 1
        for (int e = lane; e < nw * 2 * m; e += 32) {
            const int ww = wc + (e >= 2 * m ? fpt_fastdiv(e, magic) >> 1 : 0);
            unsigned mask = 0xffffffffu;
            if (ww == w0) mask &= 0xffffffffu << (l & 31);
            if (ww == w1) mask &= 0xffffffffu >> (31 - ((r - 1) & 31));
            w.wbuf[e] = planes[(size_t)wc * 2 * m + e] & mask;
        }
        __syncwarp();
        #pragma unroll 1
        for (int i = 1; i < m; i++) {
            #pragma unroll 1
            for (int j = lane; j < i; j += 32) {
                int cnt = 0;
                #pragma unroll 1
                for (int q = 0; q < nw; q++) {
                    const unsigned *row = w.wbuf + (size_t)q * 2 * m;
                    cnt += __popc(row[i] & row[m + j]) + __popc(row[m + i] & row[j]);
                }
                w.A[fpt_tri(i) + j] += (double)cnt;
            }
        }
    }
    __syncwarp();
}

/* ---- step 1b: fill_averages (css.c:337-366) on the packed triangle; returns 1 to keep the window. Every off-diagonal
   entry stands for two of the m*m entries the reference visits, the diagonal (always blank before filling) for one. */
FPT_D int fpt_warp_fill(int m, const FptEigWork &w) {
    const int lane = threadIdx.x & 31, mm = m * m;
    long long blanks = 0;
    double sum = 0.0;
    #pragma unroll 1
    for (int i = 1; i < m; i++)
        #pragma unroll 1
        for (int j = lane; j < i; j += 32) {
            const double v = w.A[fpt_tri(i) + j];
            if (v < 0.00001) blanks += 2; else sum += 2.0 * v;
        }
    blanks = fpt_warp_sum_i64(blanks) + m;
    sum = fpt_warp_sum(sum);
    if (blanks > (long long)(mm / 2)) return 0;
    const double avg = __ddiv_rn(sum, (double)mm);
    #pragma unroll 1
    for (int e = lane; e < fpt_tri(m); e += 32) if (w.A[e] < 0.00001) w.A[e] = avg;
    __syncwarp();
    return 1;
}

/* number of eigenvalues of the tridiagonal (d, e2 = e^2) below x = sign changes of the Sturm sequence p_0 = 1, p_1 = d_0 - x,
   p_i+1 = (d_i - x) p_i - e_i-1^2 p_i-1 (the leading principal minors of T - xI). The determinant form needs two multiply-adds
   and a sign test per step where the pivot form q_i = p_i / p_i-1 (LAPACK dstebz) needs a reciprocal — ncu of the eigenvector
   kernel: 36 instructions per step in the pivot form, 47 % of the kernel. Works on the matrix scaled to unit norm (|d - x| <= 2.3,
   e^2 <= 1: growth below 3.3 per step); the pair is rescaled every eight steps, which keeps it inside the fp64 range for any
   sequence that does not lose 200 decades within eight steps. A minor that is exactly zero counts as a sign change and
   continues as -1e-30 of its predecessor (the pivot form's pivmin). */
FPT_D int fpt_sturm_count(const double *d, const double *e2, int n, double x, double pivmin) {
    (void)pivmin;
    double pm = 1.0, p = d[0] - x;
    if (p == 0.0) p = -1e-30;
    int cnt = (int)((unsigned)__double2hiint(p) >> 31);
    int i = 1;
    #pragma unroll 1
    while (i < n) {
        const int stop = min(n, i + 8);
        #pragma unroll 8
        for (; i < stop; i++) {
            double pn = fma(d[i] - x, p, -(e2[i - 1] * pm));
            if (pn == 0.0) pn = -1e-30 * p;
            cnt += (int)(((unsigned)__double2hiint(pn) ^ (unsigned)__double2hiint(p)) >> 31);
            pm = p; p = pn;
        }
        const double ap = fmax(fabs(p), fabs(pm));
        if (ap > 1e100) { p *= 1e-100; pm *= 1e-100; }
        else if (ap < 1e-100) { p *= 1e100; pm *= 1e100; }
    }
    return cnt;
}

/* start of column k of the reflector store: strictly-lower part of the matrix, column by column, so that a reflector is one
   contiguous run (coalesced for lane = row): column k holds rows k+1 .. m-1 */
FPT_HD int fpt_refl_col(int m, int k) { return (k * (2 * m - k - 1)) >> 1; }

/* ---- steps 1c-2 (phase A). A holds the filled D on entry. Leaves the tridiagonal (d, e) and reflector scales (tau) in the
   work area and writes the reflectors to `refl` (global, m(m-1)/2 doubles, layout fpt_refl_col). */
FPT_D void fpt_warp_tridiag(int m, const FptEigWork &w, double *__restrict__ refl) {
    const int lane = threadIdx.x & 31;
    double *A = w.A;
    if (m == 1) {
        if (lane == 0) { w.d[0] = 0.0; w.e[0] = 0.0; w.tau[0] = 0.0; }
        __syncwarp();
        return;
    }
    /* double centring on the packed triangle; a row sum visits j < i in the row, the diagonal, then j > i down the column */
    #pragma unroll 1
    for (int e = lane; e < fpt_tri(m); e += 32) { const double v = A[e]; A[e] = v * v; }
    __syncwarp();
    #pragma unroll 1
    for (int i = lane; i < m; i += 32) {
        double s = 0.0;
        const double *row = A + fpt_tri(i);
        #pragma unroll 1
        for (int j = 0; j <= i; j++) s += row[j];
        #pragma unroll 1
        for (int j = i + 1; j < m; j++) s += A[fpt_tri(j) + i];
        w.pv[i] = s / m;
    }
    __syncwarp();
    double g = 0.0;
    #pragma unroll 1
    for (int i = 0; i < m; i++) g += w.pv[i];
    g /= m;
    #pragma unroll 1
    for (int i = 0; i < m; i++) {
        const double ri = w.pv[i];
        double *row = A + fpt_tri(i);
        #pragma unroll 1
        for (int j = lane; j <= i; j += 32) row[j] = -0.5 * (((row[j] - ri) - w.pv[j]) + g);
    }
    __syncwarp();

    /* Householder tridiagonalisation (LAPACK dsytd2, lower): H_k = I - tau v v', v(k+1) = 1, v(k+2:) stored in column k.
       Only the lower triangle is referenced and updated. Lane = row. */
    #pragma unroll 1
    for (int k = 0; k + 2 < m; k++) {
        const double x0 = A[fpt_tri(k + 1) + k];
        double s2 = 0.0;
        #pragma unroll 1
        for (int i = k + 2 + lane; i < m; i += 32) { const double x = A[fpt_tri(i) + k]; s2 += x * x; }
        s2 = fpt_warp_sum(s2);
        if (lane == 0) w.d[k] = A[fpt_tri(k) + k];
        if (s2 == 0.0) {                                  /* column already tridiagonal: H = I */
            if (lane == 0) { w.e[k] = x0; w.tau[k] = 0.0; }
            __syncwarp();
            continue;
        }
        const double nrm = sqrt(x0 * x0 + s2);
        const double alpha = x0 >= 0.0 ? -nrm : nrm;
        const double tau = (alpha - x0) / alpha;
        const double scal = 1.0 / (x0 - alpha);
        double *vv = w.y;                                 /* the reflector as a contiguous vector (y is free until step 4) */
        #pragma unroll 1
        double *rcol = refl + fpt_refl_col(m, k) - (k + 1);   /* rcol[i] = v_i, i = k+1 .. m-1 */
        for (int i = k + 2 + lane; i < m; i += 32) { const double v = A[fpt_tri(i) + k] * scal; vv[i] = v; rcol[i] = v; }
        if (lane == 0) { vv[k + 1] = 1.0; rcol[k + 1] = 1.0; w.e[k] = alpha; w.tau[k] = tau; }
        __syncwarp();
        /* p = tau * A22 v, A22 symmetric and stored as its lower triangle: row part, then down the column (index grows by j+1) */
        double pvdot = 0.0;
        #pragma unroll 1
        for (int i = k + 1 + lane; i < m; i += 32) {
            const double *row = A + fpt_tri(i);
            /* two independent accumulators halve the dependent fp64 chain (summation order differs from a single running sum
               only in rounding, far below the 1e-5 tolerance on the scores) */
            double s = 0.0, s_b = 0.0;
            int j = k + 1;
            #pragma unroll 1
            for (; j + 1 <= i; j += 2) { s += row[j] * vv[j]; s_b += row[j + 1] * vv[j + 1]; }
            if (j <= i) s += row[j] * vv[j];
            int idx = fpt_tri(i + 1) + i;
            j = i + 1;
            #pragma unroll 1
            for (; j + 1 < m; j += 2) {
                const int idx2 = idx + j + 1;
                s += A[idx] * vv[j]; s_b += A[idx2] * vv[j + 1];
                idx = idx2 + j + 2;
            }
            if (j < m) s += A[idx] * vv[j];
            s += s_b;
            s *= tau;
            w.pv[i] = s;
            pvdot += s * vv[i];
        }
        pvdot = fpt_warp_sum(pvdot);
        const double K = -0.5 * tau * pvdot;
        #pragma unroll 1
        for (int i = k + 1 + lane; i < m; i += 32) w.wv[i] = w.pv[i] + K * vv[i];
        __syncwarp();
        /* A22 -= v w' + w v' on the lower triangle */
        #pragma unroll 1
        for (int i = k + 1 + lane; i < m; i += 32) {
            double *row = A + fpt_tri(i);
            const double vi = vv[i], wi = w.wv[i];
            #pragma unroll 2
            for (int j = k + 1; j <= i; j++) row[j] -= vi * w.wv[j] + wi * vv[j];
        }
        __syncwarp();
    }
    if (lane == 0) {
        w.d[m - 2] = A[fpt_tri(m - 2) + (m - 2)];
        w.e[m - 2] = A[fpt_tri(m - 1) + (m - 2)];
        w.d[m - 1] = A[fpt_tri(m - 1) + (m - 1)];
        if (m == 2) w.tau[0] = 0.0;
    }
    __syncwarp();
}

/* ---- steps 3-4: the two largest eigenpairs (and optionally the third eigenvalue) of the symmetric tridiagonal (w.d, w.e)
   of order m >= 2, by one warp. On return w.y[0..m) and w.y[m..2m) hold the unit eigenvectors; out1..out3 the eigenvalues
   (out3 = 0 unless want_third), tnorm_out the Gershgorin norm of T. Returns 0 when T is zero or not finite (nothing
   else is written then; tnorm_out tells which). Used by the Householder path below and by the Lanczos path for large
   cohorts (fpt_css_lanczos.cuh). Needs w.d, w.e, w.pv, w.wv, w.y (2m), w.lu (6m). */
FPT_D int fpt_warp_tri_eig(int m, const FptEigWork &w, int want_third, double &out1, double &out2, double &out3, double &tnorm_out) {
    const int lane = threadIdx.x & 31;
    /* Gershgorin bounds and the norm used to scale T to O(1): the searches below run on d/tnorm, (e/tnorm)^2 (kept in
       pv / wv), which keeps every Sturm pivot inside single-precision range for the fast reciprocal */
    double glo = 1e300, ghi = -1e300;
    #pragma unroll 1
    for (int i = lane; i < m; i += 32) {
        const double el = i > 0 ? fabs(w.e[i - 1]) : 0.0, er = i < m - 1 ? fabs(w.e[i]) : 0.0;
        glo = fmin(glo, w.d[i] - el - er);
        ghi = fmax(ghi, w.d[i] + el + er);
    }
    glo = fpt_warp_min(glo); ghi = fpt_warp_max(ghi);
    const double tnorm = fmax(fabs(glo), fabs(ghi));
    tnorm_out = tnorm;
    if (!(tnorm > 0.0) || !(tnorm < 1e300)) return 0;      /* B = 0 (all dissimilarities equal) or not finite */
    const double rnorm = 1.0 / tnorm;
    #pragma unroll 1
    for (int i = lane; i < m; i += 32) {
        w.pv[i] = w.d[i] * rnorm;
        if (i < m - 1) { const double es = w.e[i] * rnorm; w.wv[i] = es * es; }
    }
    __syncwarp();
    const double pivmin = 1e-30;
    const double pad = 4.0 * 2.220446049250313e-16 * m + 2.0 * pivmin;
    glo = glo * rnorm - pad; ghi = ghi * rnorm + pad;

    /* two largest eigenvalues at once: lanes 0-15 bracket index m-1, lanes 16-31 index m-2; 17-section.
       8 rounds shrink the bracket by 17^8 = 7e9 (1.4e-10 of the spectrum width); three inverse-iteration solves with that
       shift and the Rayleigh quotient of the converged vector supply the rest. */
    const int half = lane >> 4, hl = lane & 15;
    const int want = m - 1 - half;                         /* ascending index searched by this half-warp */
    double lo = glo, hi = ghi;
    #pragma unroll 1
    for (int round = 0; round < 8; round++) {
        const double x = lo + (hi - lo) * ((double)(hl + 1) / 17.0);
        const int flag = fpt_sturm_count(w.pv, w.wv, m, x, pivmin) >= want + 1;
        const unsigned bal = (__ballot_sync(FPT_FULL_MASK, flag) >> (16 * half)) & 0xffffu;
        int js = 16;                                       /* first probe already above the eigenvalue */
        #pragma unroll 1
        for (int b = 0; b < 16; b++) if ((bal >> b) & 1u) { js = b; break; }
        const double xl = __shfl_sync(FPT_FULL_MASK, x, 16 * half + (js > 0 ? js - 1 : 0));
        const double xh = __shfl_sync(FPT_FULL_MASK, x, 16 * half + (js < 16 ? js : 15));
        if (js > 0) lo = xl;
        if (js < 16) hi = xh;
    }
    const double lam_mine = 0.5 * (lo + hi);               /* scaled units */
    double lam1 = __shfl_sync(FPT_FULL_MASK, lam_mine, 0);
    double lam2 = __shfl_sync(FPT_FULL_MASK, lam_mine, 16);
    double lam3 = 0.0;
    if (want_third && m >= 3) {                            /* diagnostics only: 33-section, 7 rounds */
        double l3 = glo, h3 = ghi;
        #pragma unroll 1
        for (int round = 0; round < 7; round++) {
            const double x = l3 + (h3 - l3) * ((double)(lane + 1) / 33.0);
            const int flag = fpt_sturm_count(w.pv, w.wv, m, x, pivmin) >= m - 2;
            const unsigned bal = __ballot_sync(FPT_FULL_MASK, flag);
            int js = 32;
            #pragma unroll 1
            for (int b = 0; b < 32; b++) if ((bal >> b) & 1u) { js = b; break; }
            const double xl = __shfl_sync(FPT_FULL_MASK, x, js > 0 ? js - 1 : 0);
            const double xh = __shfl_sync(FPT_FULL_MASK, x, js < 32 ? js : 31);
            if (js > 0) l3 = xl;
            if (js < 32) h3 = xh;
        }
        lam3 = 0.5 * (l3 + h3) * tnorm;
    }

    /* eigenvectors of T (unscaled) by inverse iteration. Lane c (0, 1) owns vector c: T - lam I is factored ONCE
       (LAPACK dgttrf scheme: partial pivoting, multipliers and swap flags kept, pivots stored as reciprocals), then
       three solves; the second vector is kept orthogonal to the first; the eigenvalue is polished by the Rayleigh
       quotient of the converged vector. Per vector: 1/pivot, du, du2 in `lu`, the multipliers in pv / wv (a multiplier of an
       interchanged row is stored as f + 4). */
    const double epsT = 2.220446049250313e-16 * tnorm;
    {   /* set-up, half a warp per vector */
        const int c = lane >> 4, hl = lane & 15;
        const double lam = (c == 0 ? lam1 : lam2) * tnorm;
        double *dd = w.lu + (size_t)c * 3 * m, *du = dd + m, *du2 = du + m;
        double *dl = (c == 0 ? w.pv : w.wv);               /* the scaled copies are no longer needed */
        double *y = w.y + (size_t)c * m;
        __syncwarp();
        #pragma unroll 1
        for (int i = hl; i < m; i += 16) {
            dd[i] = w.d[i] - lam;
            if (i < m - 1) { du[i] = w.e[i]; dl[i] = w.e[i]; }
            du2[i] = 0.0;
            const unsigned hsh = ((unsigned)i * 2654435761u + (unsigned)c * 40503u + 12345u) >> 8;
            y[i] = ((double)(hsh & 0xffffu) / 65536.0) - 0.5 + (c == 0 ? 1.0 : 0.0);
        }
    }
    __syncwarp();
    if (lane < 2) {                                        /* the elimination itself is a recurrence: one lane per vector */
        const int c = lane;
        double *dd = w.lu + (size_t)c * 3 * m, *du = dd + m, *du2 = du + m;
        double *dl = (c == 0 ? w.pv : w.wv);
        #pragma unroll 1
        for (int i = 0; i < m - 1; i++) {
            if (fabs(dd[i]) >= fabs(dl[i])) {               /* no interchange */
                if (dd[i] == 0.0) dd[i] = epsT;
                const double f = dl[i] / dd[i];
                dl[i] = f;                                  /* |f| <= 1 */
                dd[i + 1] -= f * du[i];
            } else {                                        /* interchange rows i and i+1 */
                const double f = dd[i] / dl[i];
                dd[i] = dl[i];
                dl[i] = f + 4.0;                            /* |f| < 1: a stored value above 2 flags the interchange */
                const double t = du[i];
                du[i] = dd[i + 1];
                dd[i + 1] = t - f * dd[i + 1];
                if (i < m - 2) { du2[i] = du[i + 1]; du[i + 1] = -f * du[i + 1]; }
            }
        }
        if (dd[m - 1] == 0.0) dd[m - 1] = epsT;
    }
    __syncwarp();
    {   /* pivots are only ever divided by: reciprocals, half a warp per vector */
        double *dd = w.lu + (size_t)(lane >> 4) * 3 * m;
        #pragma unroll 1
        for (int i = lane & 15; i < m; i += 16) dd[i] = 1.0 / dd[i];
    }
    __syncwarp();
    /* at most three solves; the loop stops as soon as the Rayleigh quotients of both vectors have settled (a vector error e
       moves the quotient by ~e^2, so a change below 1e-12 of the norm means the previous iterate was already good to ~1e-6
       and the current one far better). With the bracket above two solves are the rule. */
    double rq_prev = 0.0;
    #pragma unroll 1
    for (int iter = 0; iter < 3; iter++) {
        if (lane < 2) {
            const int c = lane;
            const double *rdd = w.lu + (size_t)c * 3 * m, *du = rdd + m, *du2 = du + m;
            const double *dl = (c == 0 ? w.pv : w.wv);
            double *y = w.y + (size_t)c * m;
            #pragma unroll 1
            for (int i = 0; i < m - 1; i++) {              /* forward: replay interchanges and multipliers (dgtts2) */
                const double f = dl[i];
                if (f > 2.0) { const double t = y[i]; y[i] = y[i + 1]; y[i + 1] = t - (f - 4.0) * y[i]; }
                else y[i + 1] -= f * y[i];
            }
            y[m - 1] *= rdd[m - 1];
            if (m > 1) y[m - 2] = (y[m - 2] - du[m - 2] * y[m - 1]) * rdd[m - 2];
            #pragma unroll 1
            for (int i = m - 3; i >= 0; i--) y[i] = (y[i] - du[i] * y[i + 1] - du2[i] * y[i + 2]) * rdd[i];
        }
        __syncwarp();
        /* the solves above are recurrences (one lane each); everything else on the vectors is data-parallel: lanes 0-15 work
           on vector 0, lanes 16-31 on vector 1 */
        {
            const int c = lane >> 4, hl = lane & 15;
            double *y = w.y + (size_t)c * m;
            double mx = 0.0;
            #pragma unroll 1
            for (int i = hl; i < m; i += 16) mx = fmax(mx, fabs(y[i]));
            for (int o = 8; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(FPT_FULL_MASK, mx, o));
            const bool broke = !(mx > 0.0) || !(mx < 1e300);   /* overflow / breakdown: restart from a unit vector */
            if (broke) {
                #pragma unroll 1
                for (int i = hl; i < m; i += 16) y[i] = i == c ? 1.0 : 0.0;
                mx = 1.0;
            }
            __syncwarp();
            const double rm = 1.0 / mx;
            double nn = 0.0;
            #pragma unroll 1
            for (int i = hl; i < m; i += 16) { const double v = y[i] * rm; y[i] = v; nn += v * v; }
            for (int o = 8; o > 0; o >>= 1) nn += __shfl_xor_sync(FPT_FULL_MASK, nn, o);
            nn = 1.0 / sqrt(nn);
            #pragma unroll 1
            for (int i = hl; i < m; i += 16) y[i] *= nn;
        }
        __syncwarp();
        {   /* keep the second vector orthogonal to the first */
            double *y0 = w.y, *y1 = w.y + m;
            double dot = 0.0;
            #pragma unroll 1
            for (int i = lane; i < m; i += 32) dot += y0[i] * y1[i];
            dot = fpt_warp_sum(dot);
            double nn = 0.0;
            #pragma unroll 1
            for (int i = lane; i < m; i += 32) { const double v = y1[i] - dot * y0[i]; y1[i] = v; nn += v * v; }
            nn = fpt_warp_sum(nn);
            if (nn > 0.0) {
                nn = 1.0 / sqrt(nn);
                #pragma unroll 1
                for (int i = lane; i < m; i += 32) y1[i] *= nn;
            }
        }
        __syncwarp();
        /* Rayleigh quotients y'Ty of the unit vectors: the eigenvalues to working precision (half a warp per vector) */
        const int c = lane >> 4, hl = lane & 15;
        const double *y = w.y + (size_t)c * m;
        double rq = 0.0;
        #pragma unroll 1
        for (int i = hl; i < m; i += 16) {
            const double yi = y[i];
            rq += w.d[i] * yi * yi;
            if (i < m - 1) rq += 2.0 * w.e[i] * yi * y[i + 1];
        }
        for (int o = 8; o > 0; o >>= 1) rq += __shfl_xor_sync(FPT_FULL_MASK, rq, o);
        lam1 = __shfl_sync(FPT_FULL_MASK, rq, 0);
        lam2 = __shfl_sync(FPT_FULL_MASK, rq, 16);
        const int settled = iter > 0 && fabs(rq - rq_prev) <= 1e-12 * tnorm;
        rq_prev = rq;
        if (__all_sync(FPT_FULL_MASK, settled)) break;
    }
    out1 = lam1; out2 = lam2; out3 = lam3;
    return 1;
}

/* ---- steps 3-5 (phase B): the two largest eigenpairs of the tridiagonal in the work area (d, e, tau), back-transformed
   through the reflectors in `refl`; X (2m doubles) and evals3 are written by the warp. Needs no m x m storage. */
FPT_D void fpt_warp_eig(int m, const FptEigWork &w, const double *__restrict__ refl, double *X, double *evals3, int want_third) {
    const int lane = threadIdx.x & 31;
    if (m == 1) {
        if (lane == 0) { X[0] = 0.0; X[1] = 0.0; if (evals3) { evals3[0] = 0.0; evals3[1] = 0.0; evals3[2] = 0.0; } }
        __syncwarp();
        return;
    }
    double lam1 = 0.0, lam2 = 0.0, lam3 = 0.0, tnorm = 0.0;
    if (!fpt_warp_tri_eig(m, w, want_third, lam1, lam2, lam3, tnorm)) {
        const double v = tnorm == 0.0 ? 0.0 : tnorm - tnorm;   /* 0, or NaN when the input was not finite */
        #pragma unroll 1
        for (int j = lane; j < 2 * m; j += 32) X[j] = v;
        if (lane == 0 && evals3) { evals3[0] = v; evals3[1] = v; evals3[2] = v; }
        __syncwarp();
        return;
    }

    /* back-transform: z = H_0 H_1 ... H_{m-3} y, applied last reflector first */
    #pragma unroll 1
    for (int k = m - 3; k >= 0; k--) {
        const double tau = w.tau[k];
        if (tau == 0.0) continue;
        const double *rcol = refl + fpt_refl_col(m, k) - (k + 1);
        double s0 = 0.0, s1 = 0.0;
        #pragma unroll 1
        for (int i = k + 1 + lane; i < m; i += 32) {
            const double v = rcol[i];
            s0 += v * w.y[i]; s1 += v * w.y[m + i];
        }
        s0 = tau * fpt_warp_sum(s0); s1 = tau * fpt_warp_sum(s1);
        #pragma unroll 1
        for (int i = k + 1 + lane; i < m; i += 32) {
            const double v = rcol[i];
            w.y[i] -= s0 * v; w.y[m + i] -= s1 * v;
        }
        __syncwarp();
    }
    /* css.c:558 takes sqrt of the eigenvalues unguarded: a genuinely negative one gives NaN coordinates, as in the
       reference. An eigenvalue that is zero up to rounding (rank-deficient B, e.g. always with two tracks) is clamped
       to zero instead of letting the sign of the last bit decide between 0 and NaN. */
    if (lam1 < 0.0 && -lam1 <= 1e-13 * tnorm) lam1 = 0.0;
    if (lam2 < 0.0 && -lam2 <= 1e-13 * fmax(fabs(lam1), tnorm * 1e-3)) lam2 = 0.0;
    const double r1 = sqrt(lam1), r2 = sqrt(lam2);
    #pragma unroll 1
    for (int j = lane; j < m; j += 32) { X[2 * j] = w.y[j] * r1; X[2 * j + 1] = w.y[m + j] * r2; }
    if (lane == 0 && evals3) { evals3[0] = lam1; evals3[1] = lam2; evals3[2] = lam3; }
    __syncwarp();
}

/* work areas: phase A needs the packed matrix, phase B only O(m) vectors — which is why they are separate kernels:
   the eigen-solve is a chain of dependent fp64 operations and wants as many resident warps as possible */
FPT_HD size_t fpt_tridiag_work_bytes(int m, int wch) {
    /* packed matrix + three vectors (pv, wv, the contiguous reflector); d, e and tau go straight to global memory (one
       scalar store per step), and the bit-plane words of the counting stage live in the vectors' space, which is idle until
       the reduction starts. Shared memory per warp decides how many warps an SM holds, and this kernel runs on latency hiding. */
    const size_t vec = (size_t)3 * m * 8, words = (size_t)wch * 2 * m * 4;
    size_t bytes = (size_t)fpt_tri(m) * 8 + (words > vec ? words : vec);
    return (bytes + 15) & ~(size_t)15;
}
FPT_D FptEigWork fpt_tridiag_carve(unsigned char *base, int m, int wch) {
    FptEigWork w;
    double *p = (double *)base;
    w.A = p; p += (size_t)fpt_tri(m);
    w.wbuf = (unsigned *)p;                                /* aliases the vectors below */
    w.pv = p; p += m; w.wv = p; p += m;
    w.y = p; p += m;                                       /* the contiguous reflector (phase B has its own 2m) */
    w.d = 0; w.e = 0; w.tau = 0;                           /* set per window: the window's slot of the hand-over buffer */
    w.lu = 0;
    w.wch = wch;
    return w;
}
FPT_HD size_t fpt_eigvec_work_bytes(int m) { return (size_t)13 * m * 8; }
FPT_D FptEigWork fpt_eigvec_carve(unsigned char *base, int m) {
    FptEigWork w;
    double *p = (double *)base;
    w.A = 0;
    w.d = p; p += m; w.e = p; p += m; w.tau = p; p += m;
    w.pv = p; p += m; w.wv = p; p += m;
    w.y = p; p += 2 * m;
    w.lu = p; p += 6 * m;
    w.wbuf = 0; w.wch = 0;
    return w;
}

/* phase A kernel: one WARP per window. tri_out: 3m doubles per window (d, e, tau); refl_out: m(m-1)/2 per window. */
__global__ void __launch_bounds__(128)
fpt_css_tridiag_kernel(const unsigned *__restrict__ planes, const double *__restrict__ absdiff, int m,
                       const int *__restrict__ wleft, const int *__restrict__ wright, long long nwin, int wch,
                       double *__restrict__ tri_out, double *__restrict__ refl_out, unsigned char *__restrict__ status) {
    FPT_DYN_SMEM(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const FptEigWork w = fpt_tridiag_carve(smem + (size_t)warp * fpt_tridiag_work_bytes(m, wch), m, wch);
    const size_t nrefl = (size_t)m * (m - 1) / 2;
    #pragma unroll 1
    for (long long win = (long long)blockIdx.x * nwarp + warp; win < nwin; win += (long long)gridDim.x * nwarp) {
        const int l = wleft[win], r = wright[win];
        if (r <= l) { if (lane == 0) status[win] = 0; continue; }
        if (absdiff) {
            if (lane == 0) {                               /* compare_freq, css.c:245-264 */
                double s = 0.0;
                #pragma unroll 1
                for (int i = r; i-- > l;) s = __dadd_rn(s, absdiff[i]);
                s = __ddiv_rn(s, (double)(r - l));
                w.A[0] = 0.0; w.A[1] = s; w.A[2] = 0.0;          /* packed: (0,0), (1,0), (1,1) */
            }
            __syncwarp();
        } else {
            fpt_warp_counts(planes, m, l, r, w);
        }
        if (!fpt_warp_fill(m, w)) { if (lane == 0) status[win] = 1; __syncwarp(); continue; }
        FptEigWork ww = w;                                   /* d, e, tau: written in place in the hand-over buffer */
        double *t = tri_out + (size_t)win * 3 * m;
        ww.d = t; ww.e = t + m; ww.tau = t + 2 * m;
        fpt_warp_tridiag(m, ww, refl_out + (size_t)win * nrefl);
        if (lane == 0) {                                     /* entries the reduction does not produce */
            t[m + m - 1] = 0.0;
            t[2 * m + m - 1] = 0.0;
            if (m >= 2) t[2 * m + m - 2] = 0.0;
            status[win] = 2;
        }
        __syncwarp();
    }
}

/* phase B kernel: one WARP per kept window, O(m) shared memory per warp */
__global__ void __launch_bounds__(128)
fpt_css_eigvec_kernel(int m, long long nwin, const double *__restrict__ tri_in, const double *__restrict__ refl_in,
                      const unsigned char *__restrict__ status, double *__restrict__ Xout, double *__restrict__ evals_out) {
    FPT_DYN_SMEM(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const FptEigWork w = fpt_eigvec_carve(smem + (size_t)warp * fpt_eigvec_work_bytes(m), m);
    const size_t nrefl = (size_t)m * (m - 1) / 2;
    #pragma unroll 1
    for (long long win = (long long)blockIdx.x * nwarp + warp; win < nwin; win += (long long)gridDim.x * nwarp) {
        if (status[win] != 2) continue;
        const double *t = tri_in + (size_t)win * 3 * m;
        #pragma unroll 1
        for (int i = lane; i < m; i += 32) { w.d[i] = t[i]; w.e[i] = t[m + i]; w.tau[i] = t[2 * m + i]; }
        __syncwarp();
        fpt_warp_eig(m, w, refl_in + (size_t)win * nrefl, Xout + (size_t)win * 2 * m, evals_out ? evals_out + 3 * win : 0, evals_out != 0);
        __syncwarp();
    }
}

#endif
