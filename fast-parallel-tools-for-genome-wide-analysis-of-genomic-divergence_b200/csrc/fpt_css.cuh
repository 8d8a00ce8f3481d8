/*
 * fpt_css.cuh — Cluster-Separation-Score hot path on the device.
 *
 *   fpt_css_pack_kernel     genotype codes -> two bit-planes per individual ("is 3", "is -3"),
 *                           32 SNPs per word, word-major so a window is one contiguous slab
 *   fpt_css_dissimilarity   per window: pairwise opposite-homozygote counts (compare_all,
 *                           css/css.c:277-327) or frequency difference (compare_freq, css.c:245-264),
 *                           fill_averages (css.c:337-366); classical MDS (cmds, css.c:505-560) lives in
 *                           fpt_css_eig.cuh (one warp per window) and fpt_css_lanczos.cuh (large cohorts)
 *   fpt_css_smacof_kernel   per (window, start): SMACOF majorisation (smacof / guttman_transform /
 *                           stress, css.c:767-938) from a random start (smacof_runs, css.c:852-884)
 *                           or from the classical-MDS solution (css.c:216-217)
 *   fpt_css_pick_kernel     best of the random starts (css.c:876-881)
 *   fpt_css_perm_kernel     embedding distances (calc_dist, css.c:573-587), the score (css,
 *                           css.c:608-647) and the Monte-Carlo p-value (significance_treshold /
 *                           random_shuffle, css.c:700-752)
 *
 * Paths are relative to /root/reference/statistics/.
 */
#ifndef FPT_CSS_CUH
#define FPT_CSS_CUH

#include "fpt_rt.cuh"
#include "fpt_fet.cuh"     /* fpt_code_of / fpt_stage_codes */

/* window status */
#define FPT_WIN_EMPTY 0      /* not visited or no SNPs: outputs untouched */
#define FPT_WIN_DISCARDED 1  /* fill_averages said "more than half blank": scorer returns -1 */
#define FPT_WIN_SCORED 2

/* ============================================================================================
 * Packing. planes[(word*2 + plane)*m + individual], plane 0 = genotype 3, plane 1 = genotype -3;
 * bit b of word w is SNP 32*w + b; individuals 0..asize-1 are group A, asize..m-1 group B.
 * An opposite-homozygote count over a window is then popc(P_i & M_j) + popc(M_i & P_j) summed over
 * the window's words (first/last masked) — exact small integers, like the reference's count++.
 */
template <typename T>
__global__ void __launch_bounds__(256)
fpt_css_pack_kernel(const T *__restrict__ avals, const T *__restrict__ bvals, long long nsnp, int asize, int bsize,
                    int words_per_tile, unsigned *__restrict__ planes) {
    FPT_DYN_SMEM(smem);
    const int m = asize + bsize, tile = words_per_tile * 32;
    unsigned char *sa = smem;
    unsigned char *sb = smem + (((size_t)tile * asize + 15) & ~(size_t)15);
    const long long nwords = (nsnp + 31) >> 5;
    for (long long w0 = (long long)blockIdx.x * words_per_tile; w0 < nwords; w0 += (long long)gridDim.x * words_per_tile) {
        const long long t0 = w0 * 32;
        const int nt = (int)min((long long)tile, nsnp - t0);
        const int nw = (nt + 31) >> 5;
        fpt_stage_codes<T>(avals + t0 * asize, (long long)nt * asize, sa);
        fpt_stage_codes<T>(bvals + t0 * bsize, (long long)nt * bsize, sb);
        __syncthreads();
        for (int item = threadIdx.x; item < nw * m; item += blockDim.x) {
            const int w = item / m, i = item - w * m;
            const unsigned char *col = i < asize ? sa + i : sb + (i - asize);
            const int stride = i < asize ? asize : bsize;
            unsigned P = 0, M = 0;
            const int nb = min(32, nt - w * 32);
            for (int b = 0; b < nb; b++) {
                unsigned c = col[(size_t)(w * 32 + b) * stride];
                P |= (unsigned)(c == 1) << b;
                M |= (unsigned)(c == 2) << b;
            }
            planes[((size_t)(w0 + w) * 2 + 0) * m + i] = P;
            planes[((size_t)(w0 + w) * 2 + 1) * m + i] = M;
        }
        __syncthreads();
    }
}

/* drosophila metric (compare_freq): per-SNP |a - b| of two frequency tracks */
__global__ void fpt_css_absdiff_kernel(const double *__restrict__ a, const double *__restrict__ b, long long n,
                                       double *__restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        out[i] = fabs(__dsub_rn(a[i], b[i]));
}

/* ============================================================================================
 * Per-window dissimilarity matrix D (m x m doubles, row-major) into `D`; returns 1 to keep the
 * window, 0 to discard it. `wbuf` is shared scratch for wch*2*m words, `red` 33 doubles + 33 int64.
 */
struct FptCssScratch {
    double *red;            /* 33 doubles */
    long long *redi;        /* 33 int64 */
    unsigned *wbuf;         /* wch * 2 * m words */
    int wch;
    int *pairs;             /* SMACOF: pair index p -> (i << 16 | j), or NULL when the table does not fit shared memory */
};

FPT_D void fpt_css_counts(const unsigned *__restrict__ planes, int m, int l, int r, double *D, const FptCssScratch &sc) {
    const int w0 = l >> 5, w1 = (r - 1) >> 5, mm = m * m;
    for (int e = threadIdx.x; e < mm; e += blockDim.x) D[e] = 0.0;
    for (int wc = w0; wc <= w1; wc += sc.wch) {
        const int nw = min(sc.wch, w1 - wc + 1);
        __syncthreads();
        for (int e = threadIdx.x; e < nw * 2 * m; e += blockDim.x) {
            const int w = wc + e / (2 * m);
            unsigned mask = 0xffffffffu;
            if (w == w0) mask &= 0xffffffffu << (l & 31);
            if (w == w1) mask &= 0xffffffffu >> (31 - ((r - 1) & 31));
            sc.wbuf[e] = planes[(size_t)wc * 2 * m + e] & mask;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < mm; e += blockDim.x) {
            const int i = e / m, j = e - i * m;
            if (j < i) {
                int cnt = 0;
                for (int w = 0; w < nw; w++) {
                    const unsigned *row = sc.wbuf + (size_t)w * 2 * m;
                    cnt += __popc(row[i] & row[m + j]) + __popc(row[m + i] & row[j]);
                }
                D[e] += (double)cnt;
            }
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < mm; e += blockDim.x) {
        const int i = e / m, j = e - i * m;
        if (j > i) D[e] = D[j * m + i];
    }
    __syncthreads();
}

/* css.c:337-366: blanks (< 1e-5, diagonal included) become sum/m^2; discard when blanks > m*m/2 */
FPT_D int fpt_css_fill(double *D, int m, const FptCssScratch &sc) {
    const int mm = m * m;
    long long blanks = 0;
    double sum = 0.0;
    for (int e = threadIdx.x; e < mm; e += blockDim.x) {
        double v = D[e];
        if (v < 0.00001) blanks++; else sum += v;
    }
    blanks = fpt_block_sum_i64(blanks, sc.redi);
    sum = fpt_block_sum(sum, sc.red);
    if (blanks > (long long)(mm / 2)) return 0;
    const double avg = __ddiv_rn(sum, (double)mm);
    for (int e = threadIdx.x; e < mm; e += blockDim.x) if (D[e] < 0.00001) D[e] = avg;
    __syncthreads();
    return 1;
}

/* window -> filled dissimilarity matrix; handles both metrics. absdiff != NULL selects compare_freq (m = 2). */
FPT_D int fpt_css_dissimilarity(const unsigned *__restrict__ planes, const double *__restrict__ absdiff, int m, int l, int r,
                                double *D, const FptCssScratch &sc) {
    if (absdiff) {
        /* css.c:245-264: mean over the window, accumulated from the last SNP down */
        if (threadIdx.x == 0) {
            double s = 0.0;
            for (int i = r; i-- > l;) s = __dadd_rn(s, absdiff[i]);
            s = __ddiv_rn(s, (double)(r - l));
            D[0] = 0.0; D[1] = s; D[2] = s; D[3] = 0.0;
        }
        __syncthreads();
    } else {
        fpt_css_counts(planes, m, l, r, D, sc);
    }
    return fpt_css_fill(D, m, sc);
}

/* dynamic shared memory carve-up shared by the window kernels */
struct FptCssSmem {
    double *M0, *M1;          /* two m x m matrices (or global scratch when m is too large) */
    double *X, *Z, *Zold;     /* m x 2 each: X^k, its copy for the next transform, and X^(k-1) (exact stress of the previous step) */
    FptCssScratch sc;
};

/* doubles taken by the two matrices of a window: M0 is m x m, M1 is m x (m|1) (SMACOF keeps B with an odd leading dimension) */
FPT_HD size_t fpt_css_mats_doubles(int m) { return (size_t)2 * m * m + (size_t)m; }

FPT_D FptCssSmem fpt_css_carve(unsigned char *smem, int m, int wch, int mats_in_smem, double *gscratch) {
    FptCssSmem s;
    size_t off = 0;
    const size_t mm = (size_t)m * m;
    if (mats_in_smem) { s.M0 = (double *)(smem + off); off += mm * 8; s.M1 = (double *)(smem + off); off += (mm + m) * 8; }
    else { s.M0 = gscratch; s.M1 = gscratch + mm; }
    s.X = (double *)(smem + off); off += (size_t)2 * m * 8;
    s.Z = (double *)(smem + off); off += (size_t)2 * m * 8;
    s.Zold = (double *)(smem + off); off += (size_t)2 * m * 8;
    s.sc.red = (double *)(smem + off); off += 33 * 8;
    s.sc.redi = (long long *)(smem + off); off += 33 * 8;
    off = (off + 15) & ~(size_t)15;
    s.sc.pairs = 0;
    if (mats_in_smem) { s.sc.pairs = (int *)(smem + off); off += (((size_t)m * (m - 1) / 2) * 4 + 15) & ~(size_t)15; }
    s.sc.wbuf = (unsigned *)(smem + off);
    s.sc.wch = wch;
    return s;
}

FPT_HD size_t fpt_css_smem_bytes(int m, int wch, int mats_in_smem) {
    const size_t mm = (size_t)m * m;
    size_t off = (mats_in_smem ? (2 * mm + m) * 8 : 0) + (size_t)6 * m * 8 + 66 * 8;
    off = (off + 15) & ~(size_t)15;
    if (mats_in_smem) off += (((size_t)m * (m - 1) / 2) * 4 + 15) & ~(size_t)15;   /* SMACOF pair table */
    return off + (size_t)wch * 2 * m * 4;
}

/* ============================================================================================
 * SMACOF (css.c:907-938). Per iteration, the reference's sequence: B(Z) from the current distances
 * (b_ij = -delta_ij/d_ij, 0 when d_ij < 1e-5; b_ii = -sum_j b_ij accumulated with j counting down), X = B Z / m with the
 * products accumulated over ascending j (dgemm order), new distances, stress, Z = X; loop while first pass or
 * (stress drop > eps and k <= max_iters).
 *
 * One pass over the m(m-1)/2 pairs per iteration does all the pair work: distance of the NEW configuration (IEEE sqrt,
 * same expression as calc_dist), its contribution to the stress, and b_ij for the NEXT Guttman transform — distances and B
 * are symmetric, so each pair is evaluated once and mirrored, and the distance matrix itself is never stored. B lives in a
 * matrix with an odd leading dimension, so the row pass (thread = one row and coordinate, the reference's exact operation
 * order) reads it without bank conflicts. The stress is summed as a fixed-order parallel tree and, wherever the summation
 * order could change the stopping decision or the value returned, again as the reference's single running sum
 * (fpt_css_smacof): iteration counts and results are those of the reference's operation order.
 */
FPT_HD int fpt_css_ldb(int m) { return m | 1; }

FPT_D void fpt_pair_of(int p, int &i, int &j) {             /* p = i(i-1)/2 + j, 0 <= j < i */
    int r = (int)((1.0f + sqrtf(1.0f + 8.0f * (float)p)) * 0.5f);
    while ((r * (r - 1)) / 2 > p) r--;
    while (((r + 1) * r) / 2 <= p) r++;
    i = r; j = p - (r * (r - 1)) / 2;
}

/* stress of configuration X against delta; leaves B(X) (off-diagonal) in Bm */
FPT_D double fpt_css_pairs_pass(const double *X, const double *delta, double *Bm, int m, const FptCssScratch &sc) {
    const int npairs = (m * (m - 1)) >> 1, ldb = fpt_css_ldb(m);
    double part = 0.0;
    for (int p = threadIdx.x; p < npairs; p += blockDim.x) {
        int i, j;
        if (sc.pairs) { const int pk = sc.pairs[p]; i = pk >> 16; j = pk & 0xffff; }
        else fpt_pair_of(p, i, j);
        const double dx = __dsub_rn(X[2 * i], X[2 * j]), dy = __dsub_rn(X[2 * i + 1], X[2 * j + 1]);
        const double d = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        const double dl = delta[i * m + j];
        const double err = __dsub_rn(d, dl);
        part = __dadd_rn(part, __dmul_rn(err, err));
        const double b = d < 0.00001 ? 0.0 : __ddiv_rn(__dmul_rn(-1.0, dl), d);
        Bm[i * ldb + j] = b; Bm[j * ldb + i] = b;
    }
    return fpt_block_sum(part, sc.red);                     /* has the barriers that publish Bm */
}

/* The reference's stress (css.c:767-777): ONE running sum over the pairs, i = m-1..1 and j = i-1..0 — i.e. the pair index
   p = i(i-1)/2 + j counting down. Warp 0 evaluates it: the lanes compute the next 32 terms (same expressions as the pair pass,
   so every term has the bits the reference's term has), lane 0's accumulator adds them in order (the padding terms of the last
   group are +0.0, which leaves a non-negative sum unchanged). ~10 cycles per pair on one warp — used only where the stopping
   rule or the value handed to smacof_runs depends on the summation order (fpt_css_smacof below). All threads get the result. */
FPT_D double fpt_css_stress_reforder(const double *X, const double *delta, int m, const FptCssScratch &sc) {
    const int npairs = (m * (m - 1)) >> 1;
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        double acc = 0.0;
        for (int s0 = 0; s0 < npairs; s0 += 32) {
            const int sidx = s0 + lane;
            double term = 0.0;
            if (sidx < npairs) {
                const int p = npairs - 1 - sidx;
                int i, j;
                if (sc.pairs) { const int pk = sc.pairs[p]; i = pk >> 16; j = pk & 0xffff; }
                else fpt_pair_of(p, i, j);
                const double dx = __dsub_rn(X[2 * i], X[2 * j]), dy = __dsub_rn(X[2 * i + 1], X[2 * j + 1]);
                const double d = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                const double err = __dsub_rn(d, delta[i * m + j]);
                term = __dmul_rn(err, err);
            }
#pragma unroll
            for (int l = 0; l < 32; l++) acc = __dadd_rn(acc, __shfl_sync(FPT_FULL_MASK, term, l));
        }
        if (lane == 0) sc.red[32] = acc;
    }
    __syncthreads();
    const double r = sc.red[32];
    __syncthreads();
    return r;
}

/* Every iteration's stress is first summed as a fixed-order parallel tree (fpt_css_pairs_pass). The terms are the reference's,
   only the order of the additions differs, so tree sum t and running sum r of the same n non-negative terms satisfy
   |t - r| <= (n - 1 + depth) u S (1 + O(u)), u = 2^-53, depth = additions on the longest path of the tree. The stopping rule
   `sigma_prev - sigma > eps` (css.c:921) is decided on the tree sums when they clear eps by more than that bound for both
   stresses; otherwise both running sums are evaluated in the reference's order and decide — same iteration count as the
   reference, always. The stress returned (compared between starts by smacof_runs, css.c:876) is the running sum. */
#ifndef FPT_SMACOF_BOUND_SCALE
#define FPT_SMACOF_BOUND_SCALE 1.0      /* tests build the CPU emulation with 1e12 as well: every decision on the exact path */
#endif
FPT_D double fpt_css_smacof(const double *delta, double *Bm, int m, double *X, double *Z, double *Zold, int max_iters, double eps,
                            const FptCssScratch &sc, int *iters_out) {
    const int ldb = fpt_css_ldb(m);
    const int npairs = (m * (m - 1)) >> 1;
    const int depth = (npairs + (int)blockDim.x - 1) / (int)blockDim.x + 5 + (int)(blockDim.x >> 5);
    const double order_bound = FPT_SMACOF_BOUND_SCALE * ((double)npairs + (double)depth) * 1.1102230246251565e-16 * 1.001;
    for (int e = threadIdx.x; e < 2 * m; e += blockDim.x) Z[e] = X[e];
    __syncthreads();
    double sigma = fpt_css_pairs_pass(X, delta, Bm, m, sc), prev = 0.0;
    double sigma_r = 0.0, prev_r = 0.0;                 /* running-sum values, valid when the flags say so */
    bool have_r = false, have_prev_r = false;
    int k = 0;
    for (;;) {
        if (k > 0) {
            /* css.c:921: continue while (prev - sigma) > eps and k <= max_iters */
            if (k > max_iters) break;
            const double drop = __dsub_rn(prev, sigma);
            const double slack = order_bound * (prev + sigma) + 4.0 * 1.1102230246251565e-16 * fabs(drop);
            bool go;
            if (fabs(drop - eps) > slack && prev == prev && sigma == sigma) go = drop > eps;
            else {
                if (!have_prev_r) prev_r = fpt_css_stress_reforder(Zold, delta, m, sc);
                sigma_r = fpt_css_stress_reforder(X, delta, m, sc);
                have_r = true;
                go = __dsub_rn(prev_r, sigma_r) > eps;
            }
            if (!go) break;
        }
        prev = sigma; prev_r = sigma_r; have_prev_r = have_r; have_r = false;
        k++;
        for (int i = threadIdx.x; i < m; i += blockDim.x) {               /* b_ii = -sum_{j != i} b_ij, j counting down */
            double *brow = Bm + (size_t)i * ldb;
            double dsum = 0.0;
            for (int j = m; --j > i;) dsum = __dadd_rn(dsum, brow[j]);
            for (int j = i; j--;) dsum = __dadd_rn(dsum, brow[j]);
            brow[i] = __dmul_rn(-1.0, dsum);
        }
        for (int e = threadIdx.x; e < 2 * m; e += blockDim.x) Zold[e] = Z[e];
        __syncthreads();
        for (int e = threadIdx.x; e < 2 * m; e += blockDim.x) {           /* one (row, coordinate) per thread */
            const int i = e >> 1, c = e & 1;
            const double *brow = Bm + (size_t)i * ldb;
            double acc = 0.0;
            for (int j = 0; j < m; j++) acc = __dadd_rn(acc, __dmul_rn(brow[j], Z[2 * j + c]));
            X[e] = __ddiv_rn(acc, (double)m);
        }
        __syncthreads();
        sigma = fpt_css_pairs_pass(X, delta, Bm, m, sc);
        for (int e = threadIdx.x; e < 2 * m; e += blockDim.x) Z[e] = X[e];
        __syncthreads();
    }
    if (iters_out) *iters_out = k;
    return have_r ? sigma_r : fpt_css_stress_reforder(X, delta, m, sc);
}

/* one CTA per (window, start). nruns = 4 random starts (mds 1) or 1 start from Xin (mds 2). MATS_SMEM is a template parameter so
   that, where the two matrices live in shared memory (every cohort up to ~115 individuals), the compiler sees shared-memory pointers
   and emits LDS / STS: with the choice made at run time every access to delta and B was a generic LD / ST (ncu: 7 % of the kernel's
   instructions, each paying the address-space check). */
template <bool MATS_SMEM>
__global__ void __launch_bounds__(256)
fpt_css_smacof_kernel(const unsigned *__restrict__ planes, const double *__restrict__ absdiff, int m,
                      const int *__restrict__ wleft, const int *__restrict__ wright, long long wbase, long long nwin,
                      int wch, int mats_in_smem, double *__restrict__ gscratch, int nruns, int random_start, uint64_t seed,
                      const uint64_t *__restrict__ state_override, int max_iters, double eps,
                      const double *__restrict__ Xin, double *__restrict__ Xruns, double *__restrict__ sigma_runs,
                      int *__restrict__ iters_runs, unsigned char *__restrict__ status) {
    FPT_DYN_SMEM(smem);
    /* MATS_SMEM = true: the caller guarantees mats_in_smem != 0 and the constant is used; false: the run-time value decides */
    FptCssSmem s = fpt_css_carve(smem, m, wch, MATS_SMEM ? 1 : mats_in_smem, gscratch ? gscratch + (size_t)blockIdx.x * fpt_css_mats_doubles(m) : 0);
    if (s.sc.pairs) {                                     /* the pair decode is the same for every window and iteration */
        for (int p = threadIdx.x; p < (m * (m - 1)) >> 1; p += blockDim.x) { int i, j; fpt_pair_of(p, i, j); s.sc.pairs[p] = (i << 16) | j; }
        __syncthreads();
    }
    const long long nitems = nwin * nruns;
    for (long long it = blockIdx.x; it < nitems; it += gridDim.x) {
        const long long w = it / nruns;
        const int run = (int)(it - w * nruns);
        const int l = wleft[w], r = wright[w];
        if (r <= l) { if (threadIdx.x == 0 && run == 0 && random_start) status[w] = FPT_WIN_EMPTY; continue; }
        if (!random_start && status[w] != FPT_WIN_SCORED) continue;        /* cmds already discarded it */
        const int keep = fpt_css_dissimilarity(planes, absdiff, m, l, r, s.M0, s.sc);
        if (!keep) { if (threadIdx.x == 0 && run == 0) status[w] = FPT_WIN_DISCARDED; __syncthreads(); continue; }
        if (random_start) {
            /* css.c:861-865: x then y per individual from drand48; start `run` begins 2*m*run draws in */
            const uint64_t st0 = state_override ? state_override[w] : fpt_stream_state(seed, wbase + w, FPT_STREAM_INIT);
            for (int e = threadIdx.x; e < 2 * m; e += blockDim.x) {
                uint64_t st = fpt_lcg_skip(st0, (uint64_t)run * 2 * m + e);
                s.X[e] = fpt_drand48(st);
            }
        } else {
            for (int e = threadIdx.x; e < 2 * m; e += blockDim.x) s.X[e] = Xin[(size_t)w * 2 * m + e];
        }
        __syncthreads();
        int iters = 0;
        const double sigma = fpt_css_smacof(s.M0, s.M1, m, s.X, s.Z, s.Zold, max_iters, eps, s.sc, &iters);
        for (int e = threadIdx.x; e < 2 * m; e += blockDim.x) Xruns[(size_t)it * 2 * m + e] = s.X[e];
        if (threadIdx.x == 0) {
            sigma_runs[it] = sigma;
            if (iters_runs) iters_runs[it] = iters;
            if (run == 0) status[w] = FPT_WIN_SCORED;
        }
        __syncthreads();
    }
}

/* css.c:876-881: the lowest stress wins, the earliest on ties. (The reference additionally refuses
   every run at or above its 99999 sentinel and then keeps stale coordinates, Q12; not reproduced.) */
__global__ void fpt_css_pick_kernel(const double *__restrict__ Xruns, const double *__restrict__ sigma_runs, int m,
                                    int nruns, long long nwin, const unsigned char *__restrict__ status,
                                    double *__restrict__ Xout) {
    for (long long w = blockIdx.x; w < nwin; w += gridDim.x) {
        if (status[w] != FPT_WIN_SCORED) continue;
        int best = 0;
        for (int r = 1; r < nruns; r++) if (sigma_runs[w * nruns + r] < sigma_runs[w * nruns + best]) best = r;
        for (int e = threadIdx.x; e < 2 * m; e += blockDim.x)
            Xout[(size_t)w * 2 * m + e] = Xruns[((size_t)w * nruns + best) * 2 * m + e];
    }
}

/* ============================================================================================
 * Score and permutation test. One CTA per window; the embedding distances live in shared memory.
 *
 * css(): between-group mean minus (asize+bsize) x the within-group terms, which visit only ADJACENT
 * pairs in track order; every sum is a single running sum counting down, exactly as css.c:608-647,
 * so a permuted score compares against the observed one the way it does in the reference.
 *
 * significance_treshold(): the reference shuffles ONE persistent label array again and again
 * (Fisher-Yates, m-1 draws each) and stops as soon as `treshold` permuted scores reach the observed
 * one. Here permutation k of a chunk is generated independently from the stream position k*(m-1)
 * (skip-ahead) as the permutation it applies to positions, the chain is rebuilt with a prefix scan
 * under composition, and the early stop is a prefix count of hits — same labels, same hits, same n.
 */
template <typename TrackT>
FPT_D double fpt_css_score(const double *dist, int m, const TrackT *at, const TrackT *bt, int asize, int bsize) {
    double bet = 0.0;
    for (int i = asize; i--;) {
        const double *row = dist + (size_t)at[i] * m;
        for (int j = bsize; j--;) bet = __dadd_rn(bet, row[bt[j]]);
    }
    bet = __ddiv_rn(bet, (double)((long long)asize * bsize));
    double wa = 0.0, wb = 0.0;
    if (asize > 1) {
        for (int i = asize - 1; i--;) wa = __dadd_rn(wa, dist[(size_t)at[i] * m + at[i + 1]]);
        wa = __ddiv_rn(wa, (double)((long long)asize * asize * (asize - 1)));
    }
    if (bsize > 1) {
        for (int i = bsize - 1; i--;) wb = __dadd_rn(wb, dist[(size_t)bt[i] * m + bt[i + 1]]);
        wb = __ddiv_rn(wb, (double)((long long)bsize * bsize * (bsize - 1)));
    }
    return __dsub_rn(bet, __dmul_rn((double)(asize + bsize), __dadd_rn(wa, wb)));
}

#endif
