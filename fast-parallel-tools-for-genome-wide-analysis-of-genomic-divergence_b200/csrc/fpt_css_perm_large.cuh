/*
 * fpt_css_perm_large.cuh — the general permutation kernel: any cohort size, labels in 8 or 16 bits, distance matrix and
 * label rows in shared memory when they fit and in the CTA's global scratch otherwise. It serves the cohorts the
 * all-in-shared-memory kernel of fpt_css_perm.cuh cannot hold (m > 250, BASELINE configs[4] has m = 1000).
 *
 * Reference: calc_dist css/css.c:573-587, css css.c:608-647, significance_treshold / random_shuffle css.c:700-752
 * (paths relative to /root/reference/statistics/).
 *
 * Scoring a permutation in the reference's order costs asize x bsize dependent fp64 additions fed by scattered loads of an
 * m x m matrix (250 000 per permutation at m = 1000). With `qbits` > 0 the kernel uses the same exact integer surrogate as
 * the small-cohort kernel instead: distances quantised to q <= 2^qbits, the between-group sum as a u8 tensor-core product
 * (membership rows of 32 permutations x three base-256 digit matrices of q, both operands exact integers), the two
 * adjacent-pair sums as integer gathers, and the fp64 evaluation in reference order only for the rare permutation whose
 * surrogate lies within the proven error bound E of the observed score. Decisions — and therefore hits, the stopping
 * permutation and p — are identical to the exact evaluation.
 */
#ifndef FPT_CSS_PERM_LARGE_CUH
#define FPT_CSS_PERM_LARGE_CUH

#include "fpt_css.cuh"
#include "fpt_css_perm.cuh"

/* digit matrices and membership rows of the large-cohort surrogate: k padded to whole 32-wide MMA steps, plus 16 bytes so
   that consecutive rows start in different banks / sectors */
FPT_HD int fpt_perm_large_kpad(int m) { return ((m + 31) >> 5) << 5; }
FPT_HD int fpt_perm_large_stride(int m) { return fpt_perm_large_kpad(m) + 16; }
#define FPT_LARGE_DIGITS 4        /* base-256 digits of q: qbits <= 31 */
FPT_HD size_t fpt_perm_large_sur_scratch(int m) {            /* q (u32 m x m) + the digit matrices */
    const size_t nrows = (size_t)(((m + 7) >> 3) << 3);
    size_t b = (size_t)m * m * 4 + FPT_LARGE_DIGITS * nrows * fpt_perm_large_stride(m);
    return (b + 255) & ~(size_t)255;
}

/* sum over A'xB' of q for the 32 permutations of a warp (lane L owns permutation L), any m.
   ind: the warp's 32 membership rows (stride zs bytes, shared memory); qd: digit matrices [FPT_LARGE_DIGITS][nrows][qs] in
   global memory.
   Per 8-column tile and 32-deep k-step: 8 fragment loads of A (shared), 2 of B per digit (global, L1/L2), 2 MMAs per digit. */
FPT_D long long fpt_bet_mma_large(const unsigned char *ind, int zs, const unsigned char *__restrict__ qd, int qs, int m) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int ntiles = (m + 7) >> 3, nrows = ntiles << 3, ksteps = (m + 31) >> 5;
    const unsigned char *myrows = ind + (size_t)g * zs;
    const size_t dstride = (size_t)nrows * qs;
    long long sum[4] = { 0, 0, 0, 0 };                  /* masked row sums: tile 0 rows g, g+8; tile 1 rows g, g+8 */
    for (int nt = 0; nt < ntiles; nt++) {
        int c0[FPT_LARGE_DIGITS][4], c1[FPT_LARGE_DIGITS][4];
#pragma unroll
        for (int d = 0; d < FPT_LARGE_DIGITS; d++) {
#pragma unroll
            for (int i = 0; i < 4; i++) { c0[d][i] = 0; c1[d][i] = 0; }
        }
        const unsigned char *brow = qd + (size_t)(8 * nt + g) * qs + 4 * t;
        for (int ks = 0; ks < ksteps; ks++) {
            const unsigned char *ar = myrows + 32 * ks + 4 * t;
            unsigned a0[4], a1[4];
            a0[0] = *reinterpret_cast<const unsigned *>(ar);                a0[1] = *reinterpret_cast<const unsigned *>(ar + 8 * zs);
            a0[2] = *reinterpret_cast<const unsigned *>(ar + 16);           a0[3] = *reinterpret_cast<const unsigned *>(ar + 8 * zs + 16);
            a1[0] = *reinterpret_cast<const unsigned *>(ar + 16 * zs);      a1[1] = *reinterpret_cast<const unsigned *>(ar + 24 * zs);
            a1[2] = *reinterpret_cast<const unsigned *>(ar + 16 * zs + 16); a1[3] = *reinterpret_cast<const unsigned *>(ar + 24 * zs + 16);
#pragma unroll
            for (int d = 0; d < FPT_LARGE_DIGITS; d++) {
                const unsigned char *bp = brow + (size_t)d * dstride + 32 * ks;
                const unsigned b0 = *reinterpret_cast<const unsigned *>(bp), b1 = *reinterpret_cast<const unsigned *>(bp + 16);
                fpt_mma_u8(c0[d], a0, b0, b1);
                fpt_mma_u8(c1[d], a1, b0, b1);
            }
        }
        /* columns 8nt + 2t, +1 held by this lane: keep those outside the group (padded columns are zero anyway) */
        const unsigned char *mcol = myrows + 8 * nt + 2 * t;
        const unsigned m00 = *reinterpret_cast<const unsigned short *>(mcol);
        const unsigned m01 = *reinterpret_cast<const unsigned short *>(mcol + 8 * zs);
        const unsigned m10 = *reinterpret_cast<const unsigned short *>(mcol + 16 * zs);
        const unsigned m11 = *reinterpret_cast<const unsigned short *>(mcol + 24 * zs);
#pragma unroll
        for (int i = 0; i < 4; i++) {
            long long r0 = 0, r1 = 0;
#pragma unroll
            for (int d = FPT_LARGE_DIGITS; d--;) { r0 = (r0 << 8) + c0[d][i]; r1 = (r1 << 8) + c1[d][i]; }
            const unsigned bytemask = (i & 1) ? 0xff00u : 0x00ffu;
            const unsigned mk0 = (i < 2) ? m00 : m01, mk1 = (i < 2) ? m10 : m11;
            if (!(mk0 & bytemask)) sum[i >> 1] += r0;
            if (!(mk1 & bytemask)) sum[2 + (i >> 1)] += r1;
        }
    }
#pragma unroll
    for (int r = 0; r < 4; r++) { sum[r] += __shfl_xor_sync(FPT_FULL_MASK, sum[r], 1); sum[r] += __shfl_xor_sync(FPT_FULL_MASK, sum[r], 2); }
    const int src = 4 * (lane & 7);
    const long long v0 = __shfl_sync(FPT_FULL_MASK, sum[0], src), v1 = __shfl_sync(FPT_FULL_MASK, sum[1], src);
    const long long v2 = __shfl_sync(FPT_FULL_MASK, sum[2], src), v3 = __shfl_sync(FPT_FULL_MASK, sum[3], src);
    const int sel = ((lane >> 4) << 1) | ((lane >> 3) & 1);
    return sel == 0 ? v0 : (sel == 1 ? v1 : (sel == 2 ? v2 : v3));
}

/* observed score (identity labels), all threads of the CTA: bit-identical to fpt_css_score on labels 0..m-1.
   The between-group sum runs over rows asize-1..0 and, inside a row, columns m-1..asize: contiguous memory. */
FPT_D double fpt_css_score_identity(const double *dist, int m, int asize, int bsize, double *buf, int bufdoubles) {
    __shared__ double s_val;
    const int T = blockDim.x, tid = threadIdx.x;
    const int rows_per_tile = bufdoubles / bsize;
    double bet = 0.0;
    if (rows_per_tile < 1) {
        if (tid == 0) {
            for (int i = asize; i--;) { const double *row = dist + (size_t)i * m + asize; for (int j = bsize; j--;) bet = __dadd_rn(bet, row[j]); }
        }
    } else {
        for (int ihi = asize - 1; ihi >= 0; ihi -= rows_per_tile) {
            const int ilo = ihi - rows_per_tile + 1 > 0 ? ihi - rows_per_tile + 1 : 0, nr = ihi - ilo + 1;
            for (int e = tid; e < nr * bsize; e += T) {         /* one flat sweep: many loads in flight per thread */
                const int r = e / bsize, j = e - r * bsize;
                buf[e] = dist[(size_t)(ilo + r) * m + asize + j];
            }
            __syncthreads();
            if (tid == 0) {                                   /* loads run ahead, the additions stay one chain in order */
                for (int r = nr; r--;) {
                    const double *row = buf + (size_t)r * bsize;
                    int j = bsize;
                    for (; j >= 8; j -= 8) {
                        const double x7 = row[j - 1], x6 = row[j - 2], x5 = row[j - 3], x4 = row[j - 4];
                        const double x3 = row[j - 5], x2 = row[j - 6], x1 = row[j - 7], x0 = row[j - 8];
                        bet = __dadd_rn(bet, x7); bet = __dadd_rn(bet, x6); bet = __dadd_rn(bet, x5); bet = __dadd_rn(bet, x4);
                        bet = __dadd_rn(bet, x3); bet = __dadd_rn(bet, x2); bet = __dadd_rn(bet, x1); bet = __dadd_rn(bet, x0);
                    }
                    for (; j--;) bet = __dadd_rn(bet, row[j]);
                }
            }
            __syncthreads();
        }
    }
    if (tid == 0) {
        bet = __ddiv_rn(bet, (double)((long long)asize * bsize));
        double wa = 0.0, wb = 0.0;
        if (asize > 1) {
            for (int i = asize - 1; i--;) wa = __dadd_rn(wa, dist[(size_t)i * m + i + 1]);
            wa = __ddiv_rn(wa, (double)((long long)asize * asize * (asize - 1)));
        }
        if (bsize > 1) {
            for (int i = bsize - 1; i--;) wb = __dadd_rn(wb, dist[(size_t)(asize + i) * m + asize + i + 1]);
            wb = __ddiv_rn(wb, (double)((long long)bsize * bsize * (bsize - 1)));
        }
        s_val = __dsub_rn(bet, __dmul_rn((double)(asize + bsize), __dadd_rn(wa, wb)));
    }
    __syncthreads();
    const double v = s_val;
    __syncthreads();
    return v;
}

template <typename TrackT>
FPT_D double fpt_css_score_one_thread(const double *dist, int m, const TrackT *labels, int asize, int bsize) {
    __shared__ double s_val;
    __syncthreads();
    if (threadIdx.x == 0) s_val = fpt_css_score<TrackT>(dist, m, labels, labels + asize, asize, bsize);
    __syncthreads();
    const double v = s_val;
    __syncthreads();
    return v;
}

FPT_HD size_t fpt_css_perm_smem_bytes(int m, int nthreads, int track_bytes, int dist_in_smem, int tracks_in_smem, int surrogate) {
    size_t off = dist_in_smem ? (size_t)m * m * 8 : 0;
    off += (size_t)2 * m * 8;                                   /* X */
    off += (size_t)nthreads * 4 * 2;                            /* offs, cons */
    off += 33 * 4 + 16;
    off = (off + 15) & ~(size_t)15;
    off += (size_t)m * track_bytes;                             /* carry */
    off = (off + 15) & ~(size_t)15;
    if (tracks_in_smem) off += (size_t)2 * nthreads * m * track_bytes;
    if (surrogate) { off = (off + 15) & ~(size_t)15; off += (size_t)nthreads * fpt_perm_large_stride(m); }   /* membership rows */
    off = (off + 15) & ~(size_t)15;
    off += (size_t)(m + 1) * 8;                                 /* per-n (limit, magic) of the shuffle draws */
    return off;
}

/* Fisher-Yates of fresh identity labels into `row` (shared or global memory), css.c:700-706. Optimistic first: a draw is
   rejected with probability < n / 2^31, so the loop only remembers whether one was and the permutation is replayed exactly
   (counting the draws) when it happened. `st` is the stream state at the permutation's first draw; returns the draws used. */
template <typename TrackT>
FPT_D int fpt_generate_labels(TrackT *row, int m, const uint2 *rtab, uint64_t st) {
    for (int e = 0; e < m; e++) row[e] = (TrackT)e;
    uint64_t s2 = st;
    uint32_t over = 0u;
    for (int i = m - 1; i > 0; i--) {
        const uint2 lm = rtab[i + 1];
        const uint32_t n = (uint32_t)(i + 1);
        const uint32_t r = (uint32_t)(fpt_lcg_next(s2) >> 17);
        over |= lm.x - r;
        uint32_t rem = r - __umulhi(r, lm.y) * n;
        if (rem >= n) rem -= n;
        const TrackT t = row[i]; row[i] = row[rem]; row[rem] = t;
    }
    int used = m - 1;
    if (over >> 31) {
        used = 0;
        for (int e = 0; e < m; e++) row[e] = (TrackT)e;
        for (int i = m - 1; i > 0; i--) {
            const uint2 lm = rtab[i + 1];
            const int rr = (int)fpt_randint_fast((uint32_t)(i + 1), lm.x, lm.y, st, used);
            const TrackT t = row[i]; row[i] = row[rr]; row[rr] = t;
        }
    }
    return used;
}

template <typename TrackT>
__global__ void __launch_bounds__(256, 1)
fpt_css_perm_kernel(const double *__restrict__ Xall, int m, int asize, int bsize, long long wbase, long long nwin,
                    const unsigned char *__restrict__ status, int treshold, int runs, uint64_t seed,
                    const uint64_t *__restrict__ state_override, int chain, int dist_in_smem, int tracks_in_smem,
                    double *__restrict__ gscratch, size_t gscratch_per_cta, int qbits, double *__restrict__ out_score,
                    double *__restrict__ out_p, int *__restrict__ out_hits, int *__restrict__ out_n,
                    unsigned long long *__restrict__ recheck_counter) {
    FPT_DYN_SMEM(smem);
    const int T = blockDim.x, tid = threadIdx.x;
    size_t off = 0;
    unsigned char *gs = gscratch ? (unsigned char *)gscratch + (size_t)blockIdx.x * gscratch_per_cta : 0;
    double *dist;
    if (dist_in_smem) { dist = (double *)(smem + off); off += (size_t)m * m * 8; }
    else { dist = (double *)gs; gs += (size_t)m * m * 8; }
    double *X = (double *)(smem + off); off += (size_t)2 * m * 8;
    int *offs = (int *)(smem + off); off += (size_t)T * 4;
    int *cons = (int *)(smem + off); off += (size_t)T * 4;
    int *scan = (int *)(smem + off); off += 33 * 4 + 16;
    off = (off + 15) & ~(size_t)15;
    TrackT *carry = (TrackT *)(smem + off); off += (size_t)m * sizeof(TrackT);
    off = (off + 15) & ~(size_t)15;
    TrackT *buf0, *buf1;
    if (tracks_in_smem) { buf0 = (TrackT *)(smem + off); buf1 = buf0 + (size_t)T * m; off += (size_t)2 * T * m * sizeof(TrackT); }
    else { buf0 = (TrackT *)gs; buf1 = buf0 + (size_t)T * m; gs += (((size_t)2 * T * m * sizeof(TrackT)) + 15) & ~(size_t)15; }
    /* integer surrogate (qbits > 0): q and its digit matrices in the global scratch, membership rows in shared memory */
    const int zs = fpt_perm_large_stride(m), qd_rows = ((m + 7) >> 3) << 3;
    unsigned *q = 0; unsigned char *qd = 0, *ind = 0;
    if (qbits > 0) {
        q = (unsigned *)gs; qd = (unsigned char *)(q + (size_t)m * m);
        off = (off + 15) & ~(size_t)15;
        ind = smem + off;
    }
    off = (off + 15) & ~(size_t)15;
    if (qbits > 0) off += (size_t)T * zs;
    off = (off + 15) & ~(size_t)15;
    uint2 *rtab = (uint2 *)(smem + off);
    for (int n = tid; n <= m; n += T) {
        uint2 lm;
        lm.x = n > 0 ? fpt_randint_limit((uint32_t)n) : 0u; lm.y = n > 0 ? fpt_randint_magic((uint32_t)n) : 0u;
        rtab[n] = lm;
    }
    __syncthreads();
    unsigned long long rechecks = 0;
    __shared__ double s_dmax;
    __shared__ int s_flag;

    for (long long w = blockIdx.x; w < nwin; w += gridDim.x) {
        if (status[w] != FPT_WIN_SCORED) continue;
        for (int e = tid; e < 2 * m; e += T) X[e] = Xall[(size_t)w * 2 * m + e];
        __syncthreads();
        for (int e = tid; e < m; e += T) carry[e] = (TrackT)e;
        /* Surrogate scale from the bounding box of the embedding: dub >= every distance, at most sqrt(2) above the largest,
           and it is known BEFORE the distances are, so one pass writes dist, q and the three digit matrices together. */
        double xlo = 1e308, xhi = -1e308, ylo = 1e308, yhi = -1e308;
        for (int e = tid; e < m; e += T) {
            const double x = X[2 * e], y = X[2 * e + 1];
            xlo = fmin(xlo, x); xhi = fmax(xhi, x); ylo = fmin(ylo, y); yhi = fmax(yhi, y);
        }
        for (int o = 16; o > 0; o >>= 1) {
            xlo = fmin(xlo, __shfl_xor_sync(FPT_FULL_MASK, xlo, o)); xhi = fmax(xhi, __shfl_xor_sync(FPT_FULL_MASK, xhi, o));
            ylo = fmin(ylo, __shfl_xor_sync(FPT_FULL_MASK, ylo, o)); yhi = fmax(yhi, __shfl_xor_sync(FPT_FULL_MASK, yhi, o));
        }
        double *wbox = reinterpret_cast<double *>(offs);        /* 2T ints (offs, cons) = T doubles >= 4 per warp */
        if ((tid & 31) == 0) { double *b4 = wbox + 4 * (tid >> 5); b4[0] = xlo; b4[1] = xhi; b4[2] = ylo; b4[3] = yhi; }
        __syncthreads();
        if (tid == 0) {
            for (int k = 1; k < (T >> 5); k++) {
                xlo = fmin(xlo, wbox[4 * k]); xhi = fmax(xhi, wbox[4 * k + 1]); ylo = fmin(ylo, wbox[4 * k + 2]); yhi = fmax(yhi, wbox[4 * k + 3]);
            }
            const double ex = xhi - xlo, ey = yhi - ylo;
            s_dmax = sqrt(ex * ex + ey * ey) * 1.000000000001;
        }
        __syncthreads();
        const double dmax = s_dmax;
        const bool scale_ok = qbits > 0 && (dmax > 0.0) && (dmax < 1e300);
        const double S = scale_ok ? (double)(1u << qbits) / dmax : 0.0;
        int bad = 0;
        if (qbits > 0) {
            const int words = zs >> 2;
            for (int e = tid; e < qd_rows * words; e += T) {    /* row n, columns k4 .. k4+3 */
                const int n = e / words, k4 = (e - n * words) << 2;
                unsigned w0 = 0, w1 = 0, w2 = 0, w3 = 0;
                if (n < m && k4 < m) {
                    const double xn = X[2 * n], yn = X[2 * n + 1];
                    for (int b = 0; b < 4; b++) {
                        const int k = k4 + b;
                        if (k >= m) break;
                        double d = 0.0;
                        if (k != n) {                           /* calc_dist, css.c:573-587 (symmetric bit for bit) */
                            const double dx = __dsub_rn(xn, X[2 * k]), dy = __dsub_rn(yn, X[2 * k + 1]);
                            d = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                        }
                        dist[(size_t)n * m + k] = d;
                        if (!(d == d)) bad = 1;
                        const unsigned qv = (scale_ok && d == d) ? (unsigned)__double2ll_rn(d * S) : 0u;
                        q[(size_t)n * m + k] = qv;
                        w0 |= (qv & 0xffu) << (8 * b); w1 |= ((qv >> 8) & 0xffu) << (8 * b);
                        w2 |= ((qv >> 16) & 0xffu) << (8 * b); w3 |= (qv >> 24) << (8 * b);
                    }
                }
                {
                    unsigned char *p0 = qd + (size_t)n * zs + k4;
                    const size_t ds = (size_t)qd_rows * zs;
                    *reinterpret_cast<unsigned *>(p0) = w0;
                    *reinterpret_cast<unsigned *>(p0 + ds) = w1;
                    *reinterpret_cast<unsigned *>(p0 + 2 * ds) = w2;
                    *reinterpret_cast<unsigned *>(p0 + 3 * ds) = w3;
                }
            }
        } else {
            for (int e = tid; e < m * m; e += T) {              /* calc_dist, css.c:573-587 */
                const int i = e / m, j = e - i * m;
                if (j < i) {
                    const double dx = __dsub_rn(X[2 * i], X[2 * j]), dy = __dsub_rn(X[2 * i + 1], X[2 * j + 1]);
                    const double d = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                    dist[e] = d; dist[j * m + i] = d;
                } else if (j == i) dist[e] = 0.0;
            }
        }
        bad = __syncthreads_or(bad);
        /* observed score in the reference's summation order: with identity labels the rows it walks are contiguous, so the
           CTA stages them through shared memory (the membership rows' space) and one thread adds them up in order */
        const double score = (qbits > 0) ? fpt_css_score_identity(dist, m, asize, bsize, reinterpret_cast<double *>(ind), (T * zs) >> 3)
                                         : fpt_css_score_one_thread<TrackT>(dist, m, carry, asize, bsize);
        const bool use_surrogate = scale_ok && !bad && (score == score) && (fabs(score) < 1e300);
        const double a_ = (double)asize, b_ = (double)bsize;
        const double wterm = (asize > 1 ? 1.0 / (a_ * a_) : 0.0) + (bsize > 1 ? 1.0 / (b_ * b_) : 0.0);
        /* |surrogate - reference score| <= E: quantisation (0.5/S)(1 + (a+b)(1/a^2 + 1/b^2)) on the three means, plus the
           rounding of the two fp64 evaluations: a recursive sum of n terms <= dmax is off by at most (n-1) u n dmax, i.e.
           (n-1) u dmax on the mean (u = 2^-53); the two adjacent-pair means contribute less than (a+b) u dmax each. Taken
           eight times over. */
        const double E = use_surrogate ? (0.5 / S) * (1.0 + (a_ + b_) * wterm) * 1.0000001 +
                                         8.0 * 1.2e-16 * dmax * (a_ * b_ + 2.0 * (a_ + b_)) : 0.0;
        const double invS = use_surrogate ? 1.0 / S : 0.0;
        const double c_bet = invS / (a_ * b_);
        const double c_wa = asize > 1 ? invS / (a_ * a_ * (a_ - 1.0)) : 0.0;
        const double c_wb = bsize > 1 ? invS / (b_ * b_ * (b_ - 1.0)) : 0.0;
        const int use_a = asize <= bsize;
        const uint64_t st_win = state_override ? state_override[w] : fpt_stream_state(seed, wbase + w, FPT_STREAM_RESAMPLE);
        const int draws = m - 1;
        long long stream_pos = 0;                               /* draws consumed by finished chunks */
        int hits = 0, ndone = 0;
        bool stopped = false;
        while (!stopped && hits < treshold && ndone < runs) {
            const int nvalid = min(T, runs - ndone);
            offs[tid] = tid * draws;
            __syncthreads();
            TrackT *mine = buf0 + (size_t)tid * m;
            if (!chain && ind) {
                /* independent shuffles: permutation k starts k*(m-1) draws in, nothing to repair. The swaps are random
                   accesses, so the labels are generated in SHARED memory — the membership rows' space, free until the
                   scoring below — as many rows at a time as fit, and then copied out to the global rows in one sweep. */
                const int rbytes = (int)((((size_t)m * sizeof(TrackT) + 3) >> 2) | 1) << 2;      /* odd word count: conflict-free rows */
                const int rows_fit = (int)(((size_t)T * zs) / rbytes);
                for (int base = 0; base < nvalid; base += rows_fit) {
                    const int nb = min(rows_fit, nvalid - base);
                    if (tid >= base && tid < base + nb)
                        fpt_generate_labels<TrackT>(reinterpret_cast<TrackT *>(ind + (size_t)(tid - base) * rbytes), m, rtab,
                                                    fpt_lcg_skip(st_win, (uint64_t)(ndone + tid) * (uint64_t)draws));
                    __syncthreads();
                    for (int e = tid; e < nb * m; e += T) {
                        const int rr = e / m, col = e - rr * m;
                        buf0[(size_t)(base + rr) * m + col] = reinterpret_cast<const TrackT *>(ind + (size_t)rr * rbytes)[col];
                    }
                    __syncthreads();
                }
                cons[tid] = 0;
                __syncthreads();
            } else {
                for (;;) {                                      /* generate; repair offsets after rejections */
                    int used = 0;
                    if (tid < nvalid) {
                        /* chain: stream consumed sequentially; independent: permutation k starts k*(m-1) draws in */
                        used = fpt_generate_labels<TrackT>(mine, m, rtab, fpt_lcg_skip(st_win, chain ? (uint64_t)(stream_pos + offs[tid])
                                                                                                    : (uint64_t)(ndone + tid) * (uint64_t)draws));
                    }
                    int total = 0;
                    const int incl = fpt_block_scan_incl(used, scan, &total);
                    const int want = incl - used;
                    const int bad = (chain && tid < nvalid && want != offs[tid]) ? 1 : 0;
                    if (bad) offs[tid] = want;
                    cons[tid] = total;
                    if (!__syncthreads_or(bad)) break;
                }
            }
            const int chunk_draws = cons[0];
            /* inclusive scan under composition: G_k = s_1 o ... o s_k, (f o g)[pos] = f[g[pos]] */
            TrackT *src = buf0, *dst = buf1;
            for (int d = 1; chain && d < nvalid; d <<= 1) {
                __syncthreads();
                if (tid < nvalid) {
                    const TrackT *g = src + (size_t)tid * m;
                    TrackT *o = dst + (size_t)tid * m;
                    if (tid >= d) {
                        const TrackT *f = src + (size_t)(tid - d) * m;
                        for (int e = 0; e < m; e++) o[e] = f[g[e]];
                    } else {
                        for (int e = 0; e < m; e++) o[e] = g[e];
                    }
                }
                TrackT *tmp = src; src = dst; dst = tmp;
            }
            __syncthreads();
            int hit = 0;
            {
                const bool valid = tid < nvalid;
                const TrackT *g = src + (size_t)tid * m;
                TrackT *oc = dst + (size_t)tid * m;
                if (valid && chain) for (int e = 0; e < m; e++) oc[e] = carry[g[e]];   /* labels after permutation ndone+tid+1 */
                const TrackT *o = chain ? oc : g;               /* independent: fresh identity labels every time */
                bool exact = valid && !use_surrogate;
                const bool warp_active = (tid & ~31) < nvalid;  /* the tensor-core sum is a warp-wide operation */
                if (use_surrogate && warp_active) {
                    /* one walk over the labels: membership row of the smaller group + both adjacent-pair sums */
                    unsigned char *myind = ind + (size_t)tid * zs;
                    for (int e = 0; e < zs; e += 16) *reinterpret_cast<uint4 *>(myind + e) = make_uint4(0u, 0u, 0u, 0u);
                    long long wa = 0, wb = 0;
                    if (valid) {
                        int prev = o[0];
                        if (use_a) myind[prev] = 1;
                        for (int i = 1; i < asize; i++) {
                            const int c = o[i];
                            if (use_a) myind[c] = 1;
                            wa += (long long)q[(size_t)prev * m + c]; prev = c;
                        }
                        prev = o[asize];
                        if (!use_a) myind[prev] = 1;
                        for (int i = 1; i < bsize; i++) {
                            const int c = o[asize + i];
                            if (!use_a) myind[c] = 1;
                            wb += (long long)q[(size_t)prev * m + c]; prev = c;
                        }
                    }
                    __syncwarp();
                    const long long bet = fpt_bet_mma_large(ind + (size_t)(tid & ~31) * zs, zs, qd, zs, m);
                    __syncwarp();
                    if (valid) {
                        const double approx = (double)bet * c_bet - (a_ + b_) * ((double)wa * c_wa + (double)wb * c_wb);
                        const double diff = approx - score;
                        hit = diff > 0.0;
                        exact = !(fabs(diff) > E);
                    }
                }
                if (exact) {
                    hit = fpt_css_score<TrackT>(dist, m, o, o + asize, asize, bsize) >= score ? 1 : 0;
                    if (use_surrogate) rechecks++;
                }
            }
            int chunk_hits = 0;
            const int hincl = fpt_block_scan_incl(hit, scan, &chunk_hits);
            if (tid == 0) s_flag = -1;
            __syncthreads();
            if (hit && hits + hincl == treshold) s_flag = tid;  /* the permutation at which the loop exits */
            __syncthreads();
            if (s_flag >= 0) {
                ndone += s_flag + 1; hits = treshold; stopped = true;
            } else {
                hits += chunk_hits; ndone += nvalid;
                if (chain) {
                    const TrackT *last = dst + (size_t)(nvalid - 1) * m;
                    for (int e = tid; e < m; e += T) carry[e] = last[e];
                }
                stream_pos += chunk_draws;
            }
            __syncthreads();
        }
        if (tid == 0) {
            out_score[w] = score;
            out_p[w] = __ddiv_rn(__dmul_rn((double)(hits + 1), 1.0), (double)(ndone + 1));
            if (out_hits) out_hits[w] = hits;
            if (out_n) out_n[w] = ndone;
        }
        __syncthreads();
    }
    if (recheck_counter && rechecks) atomicAdd(recheck_counter, rechecks);
}


#endif
