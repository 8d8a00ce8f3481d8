/*
 * fpt_css_perm3.cuh — score + Monte-Carlo permutation test for the cohorts the genome scans are made of (8 <= m <= 64,
 * independent shuffles): the headline kernel of BASELINE configs[2]. Same decisions as fpt_css_perm2_kernel (fpt_css_perm.cuh):
 * reference Fisher-Yates from the window's nrand48 stream (css/css.c:700-706), every permutation scored by the exact integer
 * surrogate on the u8 tensor cores, the rare permutation within the proven bound of the observed score re-scored in the
 * reference's summation order (css.c:608-647), early stop and p as css.c:727-752. What changed is how the work meets the SM
 * (ncu of the round-1 kernel: 38 % of the shared-memory wavefronts were bank-conflict replays of the random label swaps, warps
 * stalled ~2 cycles per issue each on fixed-latency chains, shared-memory loads and CTA barriers):
 *
 *   - two permutations per thread IN FLIGHT: their LCG chains, draws and swaps are independent, so one hides the other's latency;
 *   - labels live in a [position][column] layout, one 256-byte line per position: thread t owns byte (t >> 6) of word (t & 63),
 *     so the address of position idx is ONE multiply-add (base_t + 256 idx) and, whatever index a draw picks, the 32 lanes of a
 *     warp touch 32 different banks — the random swaps are conflict-free by construction;
 *   - the 48-bit LCG runs on a (32 high bits, 16 low bits) split: five integer instructions per step and the draw is a shift;
 *   - no fp64 distance matrix: distances are quantised as they are computed (scale from the bounding box of the embedding), the
 *     observed score is summed in the reference's order by one warp straight from the embedding WHILE the other warps build the
 *     digit matrices, and the rare exact re-score recomputes its distances (same expression as calc_dist, css.c:573-587, so
 *     the same bits);
 *   - one block scan per 512 permutations instead of barriers around every stage.
 */
#ifndef FPT_CSS_PERM3_CUH
#define FPT_CSS_PERM3_CUH

#include "fpt_css_perm.cuh"

#define FPT_P3_T 256                      /* threads per CTA */
#define FPT_P3_ROUND (2 * FPT_P3_T)       /* permutations per round: two per thread */

FPT_HD int fpt_css_perm3_ok(int m, int chain) { return !chain && m >= 8 && m <= 64; }
#define FPT_P3_LINE 256                   /* bytes per label position: 64 words x 4 threads per word */

FPT_HD size_t fpt_css_perm3_smem_bytes(int m) {
    const int qd_rows = ((m + 7) >> 3) << 3;
    size_t off = (size_t)2 * m * 8;                                   /* X */
    off += (size_t)m * m * 4;                                         /* q */
    off += (size_t)3 * qd_rows * FPT_QD_STRIDE;                       /* digit matrices */
    off = (off + 15) & ~(size_t)15;
    off += (size_t)2 * m * FPT_P3_LINE;                               /* labels: two sets of m lines */
    off += (size_t)FPT_P3_T * FPT_IND_STRIDE;                         /* membership rows */
    off += (size_t)(m + 1) * 8;                                       /* (limit, magic) per n */
    off = (off + 15) & ~(size_t)15;
    off += (size_t)(FPT_P3_T + 2) * 16;                               /* affine skip maps: per thread, per round, per permutation */
    off += 40 * 8 + 33 * 4 + 16;                                      /* reductions, scan */
    return off;
}

/* X <- A X + C mod 2^48 on the split state (hi = bits 16..47, lo = bits 0..15); returns nrand48's 31-bit draw (bits 17..47) */
FPT_D uint32_t fpt_p3_lcg(uint32_t &hi, uint32_t &lo) {
    const uint32_t t0 = lo * 0xE66Du + 0xBu;                          /* < 2^32 */
    hi = hi * 0xDEECE66Du + (lo * 0x5DEECu + (t0 >> 16));
    lo = t0 & 0xffffu;
    return hi >> 1;
}

FPT_D uint64_t fpt_p3_join(uint32_t hi, uint32_t lo) { return ((uint64_t)hi << 16) | (uint64_t)lo; }

/* label at position `idx` of the row whose position 0 lives at `row` (row = set base + the thread's column byte) */
FPT_D unsigned char *fpt_p3_label(unsigned char *row, int idx) { return row + (unsigned)idx * FPT_P3_LINE; }

/* reference-order score (css.c:608-647) of the labelling `lab` (byte idx -> individual, read through fpt_p3_label) with the
   distances recomputed from the embedding on the fly; one thread */
FPT_D double fpt_p3_exact_score(const double *X, unsigned char *row, int asize, int bsize) {
    double bet = 0.0;
    for (int i = asize; i--;) {
        const int a = *fpt_p3_label(row, i);
        const double xa = X[2 * a], ya = X[2 * a + 1];
        for (int j = bsize; j--;) {
            const int b = *fpt_p3_label(row, asize + j);
            const double dx = __dsub_rn(xa, X[2 * b]), dy = __dsub_rn(ya, X[2 * b + 1]);
            bet = __dadd_rn(bet, a == b ? 0.0 : __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))));
        }
    }
    bet = __ddiv_rn(bet, (double)((long long)asize * bsize));
    double wa = 0.0, wb = 0.0;
    if (asize > 1) {
        for (int i = asize - 1; i--;) {
            const int a = *fpt_p3_label(row, i), b = *fpt_p3_label(row, i + 1);
            const double dx = __dsub_rn(X[2 * a], X[2 * b]), dy = __dsub_rn(X[2 * a + 1], X[2 * b + 1]);
            wa = __dadd_rn(wa, __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))));
        }
        wa = __ddiv_rn(wa, (double)((long long)asize * asize * (asize - 1)));
    }
    if (bsize > 1) {
        for (int i = bsize - 1; i--;) {
            const int a = *fpt_p3_label(row, asize + i), b = *fpt_p3_label(row, asize + i + 1);
            const double dx = __dsub_rn(X[2 * a], X[2 * b]), dy = __dsub_rn(X[2 * a + 1], X[2 * b + 1]);
            wb = __dadd_rn(wb, __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy))));
        }
        wb = __ddiv_rn(wb, (double)((long long)bsize * bsize * (bsize - 1)));
    }
    return __dsub_rn(bet, __dmul_rn((double)(asize + bsize), __dadd_rn(wa, wb)));
}

/* the observed score (identity labels) by ONE WARP: the lanes compute the next 32 terms, every lane's accumulator adds them in
   the reference's order (all lanes hold the same running sum) */
FPT_D double fpt_p3_observed_score(const double *X, int asize, int bsize) {
    const int lane = threadIdx.x & 31;
    const int nbet = asize * bsize;
    double bet = 0.0;
    for (int t0 = 0; t0 < nbet; t0 += 32) {                           /* term t: i = asize-1 - t / bsize, j = bsize-1 - t % bsize */
        const int t = t0 + lane;
        double term = 0.0;
        if (t < nbet) {
            const int a = asize - 1 - t / bsize, b = asize + (bsize - 1 - t % bsize);
            const double dx = __dsub_rn(X[2 * a], X[2 * b]), dy = __dsub_rn(X[2 * a + 1], X[2 * b + 1]);
            term = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        }
        const int cnt = min(32, nbet - t0);
        for (int l = 0; l < cnt; l++) bet = __dadd_rn(bet, __shfl_sync(FPT_FULL_MASK, term, l));
    }
    bet = __ddiv_rn(bet, (double)((long long)asize * bsize));
    double w2[2] = { 0.0, 0.0 };
    for (int grp = 0; grp < 2; grp++) {
        const int n = grp ? bsize : asize, base = grp ? asize : 0;
        if (n <= 1) continue;
        double acc = 0.0;
        for (int t0 = 0; t0 < n - 1; t0 += 32) {                      /* term t: i = n-2 - t, pair (base+i, base+i+1) */
            const int t = t0 + lane;
            double term = 0.0;
            if (t < n - 1) {
                const int a = base + (n - 2 - t), b = a + 1;
                const double dx = __dsub_rn(X[2 * a], X[2 * b]), dy = __dsub_rn(X[2 * a + 1], X[2 * b + 1]);
                term = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
            }
            const int cnt = min(32, n - 1 - t0);
            for (int l = 0; l < cnt; l++) acc = __dadd_rn(acc, __shfl_sync(FPT_FULL_MASK, term, l));
        }
        w2[grp] = __ddiv_rn(acc, (double)((long long)n * n * (n - 1)));
    }
    return __dsub_rn(bet, __dmul_rn((double)(asize + bsize), __dadd_rn(w2[0], w2[1])));
}

/* both permutations of a thread, one Fisher-Yates step each (position i is final afterwards). MEMBER: position i belongs to the
   smaller group, so the label landing there joins the membership mask; PAIR: the adjacent pair (i, i + 1) counts, its quantised
   distance is added to the running within-group sum. */
struct FptP3Pair {
    uint32_t h0, l0, h1, l1;            /* split LCG states */
    uint32_t over0, over1;              /* bit 31 set once a draw exceeded its limit */
    uint32_t mlo0, mhi0, mlo1, mhi1;    /* membership masks (individuals 0..31, 32..63) */
    int acc0, acc1, prev0, prev1;
};

template <bool MEMBER, bool PAIR>
FPT_D void fpt_p3_step(FptP3Pair &p, int i, const uint2 lm, unsigned char *pos0, unsigned char *pos1, unsigned char *row0,
                       unsigned char *row1, const unsigned *q, int m) {
    const uint32_t n = (uint32_t)(i + 1);
    const uint32_t r0 = fpt_p3_lcg(p.h0, p.l0), r1 = fpt_p3_lcg(p.h1, p.l1);
    p.over0 |= lm.x - r0; p.over1 |= lm.x - r1;
    uint32_t rem0 = r0 - __umulhi(r0, lm.y) * n, rem1 = r1 - __umulhi(r1, lm.y) * n;
    rem0 = min(rem0, rem0 - n); rem1 = min(rem1, rem1 - n);          /* unsigned: rem - n wraps unless rem >= n */
    unsigned char *pr0 = row0 + rem0 * FPT_P3_LINE, *pr1 = row1 + rem1 * FPT_P3_LINE;
    const int a0 = *pos0, c0 = *pr0, a1 = *pos1, c1 = *pr1;
    *pos0 = (unsigned char)c0; *pr0 = (unsigned char)a0;
    *pos1 = (unsigned char)c1; *pr1 = (unsigned char)a1;
    if (MEMBER) {
        const uint32_t b0 = 1u << (c0 & 31), b1 = 1u << (c1 & 31);
        if (c0 & 32) p.mhi0 |= b0; else p.mlo0 |= b0;
        if (c1 & 32) p.mhi1 |= b1; else p.mlo1 |= b1;
    }
    if (PAIR) { p.acc0 += (int)q[c0 * m + p.prev0]; p.acc1 += (int)q[c1 * m + p.prev1]; }
    p.prev0 = c0; p.prev1 = c1;
}

/* positions [hi, lo] (descending) of both permutations */
template <bool MEMBER, bool PAIR>
FPT_D void fpt_p3_run(FptP3Pair &p, int hi, int lo, const uint2 *rtab, unsigned char *row0, unsigned char *row1, const unsigned *q, int m) {
    unsigned char *pos0 = row0 + (unsigned)hi * FPT_P3_LINE, *pos1 = row1 + (unsigned)hi * FPT_P3_LINE;
    for (int i = hi; i >= lo; i--) {
        fpt_p3_step<MEMBER, PAIR>(p, i, rtab[i + 1], pos0, pos1, row0, row1, q, m);
        pos0 -= FPT_P3_LINE; pos1 -= FPT_P3_LINE;
    }
}

/* identity labels for both sets, written cooperatively: the word of (position, column) holds the four threads of that column,
   and all four hold `position` there; the caller's next barrier publishes them */
FPT_D void fpt_p3_identity(unsigned char *labels, int m) {
    unsigned *lw = reinterpret_cast<unsigned *>(labels);
    for (int e = threadIdx.x; e < 2 * m * (FPT_P3_LINE / 4); e += blockDim.x) {
        int pos = e >> 6;
        if (pos >= m) pos -= m;
        lw[e] = (unsigned)pos * 0x01010101u;
    }
}

__global__ void __launch_bounds__(FPT_P3_T, 3)
fpt_css_perm3_kernel(const double *__restrict__ Xall, int m, int asize, int bsize, long long wbase, long long nwin,
                     const unsigned char *__restrict__ status, int treshold, int runs, uint64_t seed,
                     const uint64_t *__restrict__ state_override, int qbits, double *__restrict__ out_score,
                     double *__restrict__ out_p, int *__restrict__ out_hits, int *__restrict__ out_n,
                     unsigned long long *__restrict__ recheck_counter) {
    FPT_DYN_SMEM(smem);
    const int T = FPT_P3_T, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ndigits = (qbits + 8) >> 3;
    const int qd_rows = ((m + 7) >> 3) << 3;
    size_t off = 0;
    double *X = (double *)(smem + off); off += (size_t)2 * m * 8;
    unsigned *q = (unsigned *)(smem + off); off += (size_t)m * m * 4;
    unsigned char *qd = smem + off; off += (size_t)3 * qd_rows * FPT_QD_STRIDE;
    off = (off + 15) & ~(size_t)15;
    unsigned char *labels = smem + off; off += (size_t)2 * m * FPT_P3_LINE;
    unsigned char *ind = smem + off; off += (size_t)T * FPT_IND_STRIDE;
    uint2 *rtab = (uint2 *)(smem + off); off += (size_t)(m + 1) * 8;
    off = (off + 15) & ~(size_t)15;
    ulonglong2 *skipmap = (ulonglong2 *)(smem + off); off += (size_t)(T + 2) * 16;
    double *red = (double *)(smem + off); off += 40 * 8;
    int *scan = (int *)(smem + off);
    unsigned char *myind = ind + (size_t)tid * FPT_IND_STRIDE, *warpind = ind + (size_t)(tid & ~31) * FPT_IND_STRIDE;
    __shared__ double s_score;
    __shared__ int s_flag;
    const int use_a = asize <= bsize;
    const int draws = m - 1;
    unsigned char *row0 = labels + ((tid & 63) << 2) + (tid >> 6);  /* my two label rows (position 0): sets 0 and 1 */
    unsigned char *row1 = row0 + (size_t)m * FPT_P3_LINE;
    unsigned long long rechecks = 0;
    {
        /* affine maps x -> a x + b (mod 2^48) of the stream: [tid] to thread tid's first permutation of a round, [T] over a whole
           round, [T + 1] over one permutation */
        const uint64_t n_t = (uint64_t)tid * 2 * (uint64_t)draws;
        const uint64_t b_t = fpt_lcg_skip(0ULL, n_t);
        skipmap[tid] = make_ulonglong2((fpt_lcg_skip(1ULL, n_t) - b_t) & FPT_MASK48, b_t);
        if (tid < 2) {
            const uint64_t n_c = tid == 0 ? (uint64_t)FPT_P3_ROUND * (uint64_t)draws : (uint64_t)draws;
            const uint64_t b_c = fpt_lcg_skip(0ULL, n_c);
            skipmap[T + tid] = make_ulonglong2((fpt_lcg_skip(1ULL, n_c) - b_c) & FPT_MASK48, b_c);
        }
    }
    for (int n = tid; n <= m; n += T) {
        uint2 lm;
        lm.x = n > 0 ? fpt_randint_limit((uint32_t)n) : 0u; lm.y = n > 0 ? fpt_randint_magic((uint32_t)n) : 0u;
        rtab[n] = lm;
    }
    __syncthreads();

    for (long long w = blockIdx.x; w < nwin; w += gridDim.x) {
        if (status[w] != FPT_WIN_SCORED) continue;
        for (int e = tid; e < 2 * m; e += T) X[e] = Xall[(size_t)w * 2 * m + e];
        __syncthreads();
        /* surrogate scale from the bounding box of the embedding: dmax >= every distance, known before the distances are */
        double dmax;
        {
            double xlo = 1e308, xhi = -1e308, ylo = 1e308, yhi = -1e308;
            int nan_ = 0;
            for (int e = lane; e < m; e += 32) {
                const double x = X[2 * e], y = X[2 * e + 1];
                if (!(x == x) || !(y == y)) nan_ = 1;
                xlo = fmin(xlo, x); xhi = fmax(xhi, x); ylo = fmin(ylo, y); yhi = fmax(yhi, y);
            }
            for (int o = 16; o > 0; o >>= 1) {
                xlo = fmin(xlo, __shfl_xor_sync(FPT_FULL_MASK, xlo, o)); xhi = fmax(xhi, __shfl_xor_sync(FPT_FULL_MASK, xhi, o));
                ylo = fmin(ylo, __shfl_xor_sync(FPT_FULL_MASK, ylo, o)); yhi = fmax(yhi, __shfl_xor_sync(FPT_FULL_MASK, yhi, o));
                nan_ |= __shfl_xor_sync(FPT_FULL_MASK, nan_, o);
            }
            const double ex = xhi - xlo, ey = yhi - ylo;
            dmax = nan_ ? 0.0 : sqrt(ex * ex + ey * ey) * 1.000000000001;      /* every warp computes the same value */
        }
        const bool scale_ok = (dmax > 0.0) && (dmax < 1e300);
        const double S = scale_ok ? (double)(1u << qbits) / dmax : 0.0;
        if (warp == 0) {
            /* the observed score in the reference's order, while the other warps quantise */
            const double sc = fpt_p3_observed_score(X, asize, bsize);
            if (lane == 0) s_score = sc;
        } else {
            /* quantised distances q = rint(d S) <= 2^qbits (|q/S - d| <= 0.5/S), symmetric, zero diagonal */
            for (int e = tid - 32; e < m * m; e += T - 32) {
                const int i = e / m, j = e - i * m;
                if (j < i) {
                    const double dx = __dsub_rn(X[2 * i], X[2 * j]), dy = __dsub_rn(X[2 * i + 1], X[2 * j + 1]);
                    const double d = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                    const unsigned qv = (scale_ok && d == d) ? (unsigned)__double2ll_rn(d * S) : 0u;
                    q[e] = qv; q[j * m + i] = qv;
                } else if (j == i) q[e] = 0u;
            }
        }
        __syncthreads();
        const double score = s_score;
        const bool use_surrogate = scale_ok && (score == score) && (fabs(score) < 1e300);
        /* base-256 digits of q, zero padded to 8-row / 64-column tiles */
        for (int e = tid; e < ndigits * qd_rows * 16; e += T) {
            const int d = e / (qd_rows * 16), rem = e - d * qd_rows * 16, n = rem >> 4, k4 = (rem & 15) << 2;
            unsigned wv = 0;
            if (n < m) {
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const int k = k4 + b;
                    const unsigned v = k < m ? ((q[n * m + k] >> (8 * d)) & 0xffu) : 0u;
                    wv |= v << (8 * b);
                }
            }
            *reinterpret_cast<unsigned *>(qd + ((size_t)d * qd_rows + n) * FPT_QD_STRIDE + k4) = wv;
        }
        /* |surrogate - reference score| <= E: quantisation (0.5/S)(1 + (a+b)(1/a^2 + 1/b^2)) on the three means plus a generous
           bound on the fp64 rounding of both evaluations (same bound as fpt_css_perm2_kernel) */
        const double a_ = (double)asize, b_ = (double)bsize;
        const double wterm = (asize > 1 ? 1.0 / (a_ * a_) : 0.0) + (bsize > 1 ? 1.0 / (b_ * b_) : 0.0);
        const double E = use_surrogate ? (0.5 / S) * (1.0 + (a_ + b_) * wterm) * 1.0000001 + 1e-11 * dmax * (1.0 + (a_ + b_)) : 0.0;
        const double invS = use_surrogate ? 1.0 / S : 0.0;
        const double c_bet = invS / (a_ * b_);
        const double c_wa = asize > 1 ? invS / (a_ * a_ * (a_ - 1.0)) : 0.0;
        const double c_wb = bsize > 1 ? invS / (b_ * b_ * (b_ - 1.0)) : 0.0;
        const uint64_t st_win = state_override ? state_override[w] : fpt_stream_state(seed, wbase + w, FPT_STREAM_RESAMPLE);
        uint64_t st_round = st_win;                         /* window state advanced by the finished rounds */
        int hits = 0, ndone = 0;
        bool stopped = false;
        fpt_p3_identity(labels, m);
        __syncthreads();
        while (!stopped && hits < treshold && ndone < runs) {
            const int nvalid = min(FPT_P3_ROUND, runs - ndone);
            const int first = 2 * tid;                      /* my permutations of this round: first, first + 1 */
            const int mycount = max(0, min(2, nvalid - first));
            const bool warp_active = (tid & ~31) * 2 < nvalid;
            int hit0 = 0, hit1 = 0;
            if (warp_active) {
                /* ---- both shuffles at once */
                uint32_t h0, l0, h1, l1;
                {
                    const ulonglong2 sk = skipmap[tid], s1 = skipmap[T + 1];
                    const uint64_t a = (sk.x * st_round + sk.y) & FPT_MASK48, b = (s1.x * a + s1.y) & FPT_MASK48;
                    h0 = (uint32_t)(a >> 16); l0 = (uint32_t)a & 0xffffu; h1 = (uint32_t)(b >> 16); l1 = (uint32_t)b & 0xffffu;
                }
                /* Fisher-Yates fixes position i at step i (i = m-1 .. 1), so everything the score needs is read off the swap as it
                   happens: the label that lands at i joins the membership mask of its group (m <= 64: two words per permutation)
                   and closes the adjacent pair (i, i+1), whose quantised distance goes to the within-B sum while i >= asize and
                   to the within-A sum below asize - 1 (the pair across the group boundary counts for neither, css.c:627-643).
                   The position ranges are split so that each loop has a fixed body. */
                FptP3Pair pp;
                pp.h0 = h0; pp.l0 = l0; pp.h1 = h1; pp.l1 = l1;
                pp.over0 = pp.over1 = 0u; pp.mlo0 = pp.mhi0 = pp.mlo1 = pp.mhi1 = 0u;
                pp.acc0 = pp.acc1 = 0; pp.prev0 = pp.prev1 = 0;
                int wb0 = 0, wb1 = 0;
                /* group B positions m-1 .. asize (the first one closes no pair) */
                if (use_a) {
                    if (m - 1 >= asize && m - 1 >= 1) fpt_p3_run<false, false>(pp, m - 1, m - 1, rtab, row0, row1, q, m);
                    fpt_p3_run<false, true>(pp, m - 2, max(asize, 1), rtab, row0, row1, q, m);
                } else {
                    if (m - 1 >= asize && m - 1 >= 1) fpt_p3_run<true, false>(pp, m - 1, m - 1, rtab, row0, row1, q, m);
                    fpt_p3_run<true, true>(pp, m - 2, max(asize, 1), rtab, row0, row1, q, m);
                }
                wb0 = pp.acc0; wb1 = pp.acc1; pp.acc0 = 0; pp.acc1 = 0;
                /* position asize-1: first of group A from the top, its pair with position asize crosses the boundary */
                if (asize - 1 >= 1) {
                    if (use_a) fpt_p3_run<true, false>(pp, asize - 1, asize - 1, rtab, row0, row1, q, m);
                    else fpt_p3_run<false, false>(pp, asize - 1, asize - 1, rtab, row0, row1, q, m);
                    if (use_a) fpt_p3_run<true, true>(pp, asize - 2, 1, rtab, row0, row1, q, m);
                    else fpt_p3_run<false, true>(pp, asize - 2, 1, rtab, row0, row1, q, m);
                }
                {   /* position 0 keeps what is left there */
                    const int c0 = *row0, c1 = *row1;
                    if (use_a) {
                        const uint32_t b0 = 1u << (c0 & 31), b1 = 1u << (c1 & 31);
                        if (c0 & 32) pp.mhi0 |= b0; else pp.mlo0 |= b0;
                        if (c1 & 32) pp.mhi1 |= b1; else pp.mlo1 |= b1;
                    }
                    if (asize >= 2) { pp.acc0 += (int)q[c0 * m + pp.prev0]; pp.acc1 += (int)q[c1 * m + pp.prev1]; }
                }
                const uint32_t over0 = pp.over0, over1 = pp.over1;
                uint32_t mlo0 = pp.mlo0, mhi0 = pp.mhi0, mlo1 = pp.mlo1, mhi1 = pp.mhi1;
                const int acc0 = pp.acc0, acc1 = pp.acc1;
                int wa0 = acc0, wa1 = acc1;
                /* a rejected draw (probability < n / 2^31 each): replay that permutation on the exact path and walk its labels */
                if ((over0 | over1) >> 31) {
                    for (int k = 0; k < 2; k++) {
                        if (!(((k ? over1 : over0) >> 31) & 1u)) continue;
                        uint64_t st = fpt_lcg_skip(st_win, (uint64_t)(ndone + first + k) * (uint64_t)draws);
                        int used = 0;
                        unsigned char *row = k ? row1 : row0;
                        for (int e = 0; e < m; e++) *fpt_p3_label(row, e) = (unsigned char)e;
                        for (int i = m - 1; i > 0; i--) {
                            const uint2 lm = rtab[i + 1];
                            const int rr = (int)fpt_randint_fast((uint32_t)(i + 1), lm.x, lm.y, st, used);
                            unsigned char *pi = fpt_p3_label(row, i), *pr = fpt_p3_label(row, rr);
                            const unsigned char t = *pi; *pi = *pr; *pr = t;
                        }
                        uint32_t lo = 0u, hi = 0u;
                        int wa = 0, wb = 0, prev = 0;
                        for (int i = 0; i < m; i++) {
                            const int c = *fpt_p3_label(row, i);
                            const bool in_a = i < asize;
                            if (in_a == (use_a != 0)) { if (c & 32) hi |= 1u << (c & 31); else lo |= 1u << (c & 31); }
                            if (i != 0 && i != asize) { const int qv = (int)q[prev * m + c]; if (in_a) wa += qv; else wb += qv; }
                            prev = c;
                        }
                        if (k) { mlo1 = lo; mhi1 = hi; wa1 = wa; wb1 = wb; } else { mlo0 = lo; mhi0 = hi; wa0 = wa; wb0 = wb; }
                    }
                }
                /* ---- score both, one after the other (the tensor-core product is a warp-wide operation) */
#pragma unroll 1
                for (int k = 0; k < 2; k++) {
                    const bool valid = k < mycount;
                    unsigned char *row = k ? row1 : row0;
                    int hit = 0;
                    bool exact = valid && !use_surrogate;
                    if (use_surrogate) {
                        /* membership row of the smaller group (the A operand of the u8 MMA): four mask bits -> four 0/1 bytes */
                        const uint32_t lo = valid ? (k ? mlo1 : mlo0) : 0u, hi = valid ? (k ? mhi1 : mhi0) : 0u;
                        uint4 *row4 = reinterpret_cast<uint4 *>(myind);
#pragma unroll
                        for (int g = 0; g < 4; g++) {
                            const uint32_t src = g < 2 ? lo : hi, sh = (uint32_t)(g & 1) * 16u;
                            uint4 v;
                            v.x = (((src >> sh) & 0xfu) * 0x00204081u) & 0x01010101u;
                            v.y = (((src >> (sh + 4u)) & 0xfu) * 0x00204081u) & 0x01010101u;
                            v.z = (((src >> (sh + 8u)) & 0xfu) * 0x00204081u) & 0x01010101u;
                            v.w = (((src >> (sh + 12u)) & 0xfu) * 0x00204081u) & 0x01010101u;
                            row4[g] = v;
                        }
                        __syncwarp();
                        const long long bet = (long long)fpt_bet_mma(warpind, qd, m, ndigits);
                        __syncwarp();
                        if (valid) {
                            const int wa = k ? wa1 : wa0, wb = k ? wb1 : wb0;
                            const double approx = (double)bet * c_bet - (a_ + b_) * ((double)wa * c_wa + (double)wb * c_wb);
                            const double diff = approx - score;
                            hit = diff > 0.0;
                            exact = !(fabs(diff) > E);
                        }
                    }
                    if (exact) {
                        hit = fpt_p3_exact_score(X, row, asize, bsize) >= score ? 1 : 0;
                        rechecks++;
                    }
                    if (k == 0) hit0 = hit; else hit1 = hit;
                }
            }
            const int myhits = hit0 + hit1;
            int round_hits = 0;
            const int hincl = fpt_block_scan_incl(myhits, scan, &round_hits);
            if (tid == 0) s_flag = -1;
            __syncthreads();
            if (myhits > 0 && hits + hincl >= treshold && hits + hincl - myhits < treshold) {
                /* the treshold-th hit is one of mine: which permutation */
                const int need = treshold - (hits + hincl - myhits);
                s_flag = first + ((need == 1 && hit0) ? 0 : 1);
            }
            __syncthreads();
            if (s_flag >= 0) {
                ndone += s_flag + 1; hits = treshold; stopped = true;
            } else {
                hits += round_hits; ndone += nvalid;
                const ulonglong2 sk = skipmap[T];
                st_round = (sk.x * st_round + sk.y) & FPT_MASK48;
                fpt_p3_identity(labels, m);                 /* every warp is past its last use of the labels (barriers above) */
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
    (void)red;
    if (recheck_counter && rechecks) atomicAdd(recheck_counter, rechecks);
}

#endif
