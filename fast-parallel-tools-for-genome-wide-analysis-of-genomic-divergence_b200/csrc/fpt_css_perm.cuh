/*
 * fpt_css_perm.cuh — score + Monte-Carlo permutation test for cohorts that fit shared memory (m <= 250)
 * (reference: calc_dist css/css.c:573-587, css css.c:608-647, random_shuffle css.c:700-706,
 *  significance_treshold css.c:727-752).
 *
 * One CTA per window, embedding distances in shared memory, PP consecutive permutations per thread.
 *
 * Exactness without paying for it. The reference decides `permuted score >= observed score` on sums
 * accumulated in one fixed order; reproducing that order costs asize*bsize + m - 2 dependent fp64
 * adds fed by 8-byte shared-memory gathers per permutation (shared-memory-wavefront bound: 30 ms per chromosome of
 * BASELINE configs[2] against 5.6 ms now). Here every permutation is first scored with a cheap, order-independent
 * surrogate:
 *     distances quantised to integers q = round(d * S) (4-byte gathers, integer adds, no rounding at all),
 *     between-group sum through the identity  sum_{A'xB'} q = sum_{i in G} rowsum_q(i) - 2 sum_{i<j in G} q_ij
 *     (G = the smaller group; exact in integers), i.e. |G| + |G|(|G|-1)/2 gathers instead of asize*bsize — or, for
 *     8 <= m <= 64, as a u8 tensor-core product of the 0/1 membership rows with the base-256 digits of q (below).
 * The surrogate differs from the true score by at most E (quantisation bound + fp bound, derived below); only when
 * |surrogate - observed| <= E — a handful of permutations per thousand windows — is the score recomputed in the
 * reference's order. Decisions, hence hits, early-stop index and p, are those of the reference arithmetic.
 *
 * Shuffles. mode 0 (default): permutation k is the reference's Fisher-Yates shuffle applied to FRESH identity
 * labels with the window's nrand48 stream positioned at k*(m-1) draws (counter-based: any permutation can be
 * regenerated on its own). mode 1 ("chain"): the reference's exact semantics — one persistent label array
 * shuffled again and again, stream consumed sequentially with rejections: thread blocks of PP permutations are
 * shuffled from identity to get the block's net permutation, a prefix scan under composition gives every block
 * its starting labels, and the block is replayed from there (bit-identical to significance_treshold run from
 * the same 48-bit state on identity labels).
 */
#ifndef FPT_CSS_PERM_CUH
#define FPT_CSS_PERM_CUH

#include "fpt_css.cuh"

#define FPT_PERM_PP 4                     /* permutations per thread per chunk */

FPT_HD int fpt_perm_row_stride(int m) {   /* bytes; a multiple of 4 whose word count is odd: conflict-free rows */
    int w = (m + 3) >> 2;
    if ((w & 1) == 0) w++;
    return w << 2;
}

FPT_HD int fpt_css_perm2_uses_mma(int m) { return m >= 8 && m <= 64; }

FPT_HD size_t fpt_css_perm2_smem_bytes(int m, int nthreads, int chain) {
    size_t off = (size_t)m * m * 8;                       /* dist */
    if (fpt_css_perm2_uses_mma(m)) off += (size_t)3 * (((m + 7) >> 3) << 3) * 80;   /* digit matrices, FPT_QD_STRIDE */
    off += (size_t)m * m * 4;                             /* q */
    off += (size_t)m * 4;                                 /* rowsum_q */
    off = (off + 7) & ~(size_t)7;
    off += (size_t)(m + 1) * 8;                           /* per-n (limit, magic) of the shuffle draws */
    off += (size_t)2 * m * 8;                             /* X */
    off += (size_t)nthreads * 4 * 2;                      /* offs, used */
    off += 33 * 4 + 12;
    off = (off + 15) & ~(size_t)15;
    off += (size_t)fpt_perm_row_stride(m);                /* carry */
    off = (off + 15) & ~(size_t)15;
    off += (size_t)nthreads * fpt_perm_row_stride(m) * (chain ? 2 : 1);
    if (fpt_css_perm2_uses_mma(m)) { off = (off + 15) & ~(size_t)15; off += (size_t)nthreads * 80; }   /* membership rows, FPT_IND_STRIDE */
    off = (off + 15) & ~(size_t)15;
    off += (size_t)(nthreads + 1) * 16;                   /* affine skip-ahead maps: one per thread, one per chunk */
    return off;
}

/* reference-order score on byte labels (same as fpt_css_score<unsigned char>) is reused for the rare recheck */

/* Fisher-Yates on a byte row, css.c:700-706 */
FPT_D void fpt_shuffle_row(unsigned char *row, int m, const uint2 *rtab, uint64_t &st, int &used) {
    for (int i = m - 1; i > 0; i--) {
        const uint2 lm = rtab[i + 1];                       /* (limit, magic) of n = i + 1, same for every thread */
        const int rr = (int)fpt_randint_fast((uint32_t)(i + 1), lm.x, lm.y, st, used);
        const unsigned char t = row[i]; row[i] = row[rr]; row[rr] = t;
    }
}

/* The same Fisher-Yates, optimistic: a draw is rejected with probability < n / 2^31, so the loop only REMEMBERS whether one
   was (no branch, no draw counter). Returns true and advances `st` by m - 1 draws when none was; otherwise `st` is left
   alone and the caller replays the permutation on the exact path above. */
FPT_D bool fpt_shuffle_row_optimistic(unsigned char *row, int m, const uint2 *rtab, uint64_t &st) {
    uint64_t s = st;
    uint32_t over = 0u;                                     /* bit 31 set once some r > limit (both are < 2^31) */
    for (int i = m - 1; i > 0; i--) {
        const uint2 lm = rtab[i + 1];
        const uint32_t n = (uint32_t)(i + 1);
        const uint32_t r = (uint32_t)(fpt_lcg_next(s) >> 17);
        over |= lm.x - r;
        uint32_t rem = r - __umulhi(r, lm.y) * n;
        if (rem >= n) rem -= n;
        const unsigned char t = row[i]; row[i] = row[rem]; row[rem] = t;
    }
    const bool rejected = (over >> 31) != 0u;
    if (!rejected) st = s;
    return !rejected;
}

FPT_D void fpt_identity_row(unsigned char *row, int m) {
    /* 4 labels per store; the row is 4-byte aligned and padded to a multiple of 4 */
    unsigned *r4 = reinterpret_cast<unsigned *>(row);
    for (int e = 0; e < m; e += 4) r4[e >> 2] = (unsigned)e * 0x01010101u + 0x03020100u;
}

/* integer surrogate of the score for labels `row`: returns sum over A'xB' of q and the two adjacent-pair sums */
FPT_D void fpt_surrogate(const unsigned *q, const int *rowsum, int m, const unsigned char *row, int asize, int bsize,
                         int use_a, long long &bet, int &wa, int &wb) {
    const unsigned char *g = use_a ? row : row + asize;
    const int ng = use_a ? asize : bsize;
    int rs = 0, pr = 0;
    for (int i = 0; i < ng; i++) {
        const int gi = g[i];
        rs += rowsum[gi];
        const unsigned *qi = q + gi * m;
        for (int j = i + 1; j < ng; j++) pr += (int)qi[g[j]];
    }
    bet = (long long)rs - 2LL * (long long)pr;
    int a = 0, b = 0;
    for (int i = 0; i + 1 < asize; i++) a += (int)q[row[i] * m + row[i + 1]];
    for (int i = 0; i + 1 < bsize; i++) b += (int)q[row[asize + i] * m + row[asize + i + 1]];
    wa = a; wb = b;
}

/* ---------------------------------------------------------------------------------------------------------------
 * Tensor-core form of the between-group sum (cohorts of up to 64 individuals).
 *
 * For 16 permutations at a time,  R = Z Q  with Z (16 x m) the 0/1 membership of the smaller group G and Q the m x m
 * matrix of quantised distances, then  sum_{A'xB'} q = sum_j (1 - z_j) R_j.  Q is split into three base-256 digits
 * so that the product runs on the integer tensor cores (mma.sync m16n8k32, u8 x u8 -> s32) and stays EXACT — the
 * surrogate keeps its "no rounding at all" property. This replaces |G| + |G|(|G|-1)/2 scattered shared-memory gathers
 * per permutation by ~4 conflict-free fragment loads and 2 MMAs.
 *
 * Fragment layout (PTX ISA, m16n8k32 with 8-bit operands; g = lane/4, t = lane%4):
 *   A (16x32, row): a0 = row g,   cols 4t..4t+3;  a1 = row g+8, cols 4t..4t+3;  a2/a3 = same rows, cols 16+4t..
 *   B (32x8,  col): b0 = rows 4t..4t+3 of column g;  b1 = rows 16+4t.. of column g
 *   C (16x8):       c0,c1 = row g, cols 2t,2t+1;     c2,c3 = row g+8, cols 2t,2t+1
 */
#define FPT_QD_STRIDE 80          /* bytes per digit row: 64 + 16 padding -> the 8 column groups hit disjoint banks */

#ifndef FPT_EMU
FPT_D void fpt_mma_u8(int (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
#else
/* CPU emulation of the warp-wide MMA with the layout documented above (tests/emu) */
static inline void fpt_mma_u8(int (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    const uint32_t regs[6] = { a[0], a[1], a[2], a[3], b0, b1 };
    const uint32_t (*all)[8] = emu_warp_publish(regs, 6);
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int i = 0; i < 4; i++) {
        const int row = g + 8 * (i >> 1), col = 2 * t + (i & 1);
        int acc = c[i];
        for (int k = 0; k < 32; k++) {
            /* A(row, k): register (row>>3) + 2*(k>>4) of lane (row&7)*4 + ((k&15)>>2), byte k&3 */
            const int av = (int)((all[(row & 7) * 4 + ((k & 15) >> 2)][(row >> 3) + 2 * (k >> 4)] >> (8 * (k & 3))) & 0xffu);
            /* B(k, col): register 4 + (k>>4) of lane col*4 + ((k&15)>>2), byte k&3 */
            const int bv = (int)((all[col * 4 + ((k & 15) >> 2)][4 + (k >> 4)] >> (8 * (k & 3))) & 0xffu);
            acc += av * bv;
        }
        c[i] = acc;
    }
    emu_warp_release();
}
#endif

/* Membership rows: ind[lane][individual] = 1 when the individual belongs to the smaller group of the permutation owned by
   that lane, else 0; 64 bytes used per row, rows FPT_IND_STRIDE = 80 bytes (20 words) apart so that the eight rows times
   four words of one fragment load fall into 32 different banks. The rows ARE the A operand of the u8 MMA. */
#define FPT_IND_STRIDE 80

FPT_D void fpt_ind_row_clear(unsigned char *myrow) {
    uint4 *z = reinterpret_cast<uint4 *>(myrow);
    z[0] = z[1] = z[2] = z[3] = make_uint4(0u, 0u, 0u, 0u);
}

/* sum over A'xB' of q for the permutation owned by THIS lane, computed cooperatively by the warp.
   ind = the warp's 32 membership rows (written by their lanes, __syncwarp()ed by the caller);
   qd = digit matrices [ndigits][nrows][FPT_QD_STRIDE], nrows = m rounded up to 8.
   The warp's 32 permutations form two 16-row tiles; every B fragment is loaded once and used by both. */
FPT_D int fpt_bet_mma(const unsigned char *ind, const unsigned char *qd, int m, int ndigits) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int ntiles = (m + 7) >> 3, nrows = ntiles << 3, ksteps = (m + 31) >> 5;
    /* my four rows: tile 0 rows g, g+8 (lanes g, g+8) and tile 1 rows g, g+8 (lanes 16+g, 24+g) */
    const unsigned char *myrows = ind + (size_t)g * FPT_IND_STRIDE;
    /* A fragments of m16n8k32: reg 0 = (row g, k 4t..4t+3), 1 = (row g+8, same k), 2 / 3 = the same rows at k + 16;
       a0 = k-step 0 (individuals 0..31), a1 = k-step 1 (32..63) */
    unsigned a0[2][4], a1[2][4] = { { 0u, 0u, 0u, 0u }, { 0u, 0u, 0u, 0u } };
#pragma unroll
    for (int tile = 0; tile < 2; tile++) {
        const unsigned char *lo_row = myrows + (size_t)(16 * tile) * FPT_IND_STRIDE + 4 * t;
        const unsigned char *hi_row = lo_row + 8 * FPT_IND_STRIDE;
        a0[tile][0] = *reinterpret_cast<const unsigned *>(lo_row);      a0[tile][1] = *reinterpret_cast<const unsigned *>(hi_row);
        a0[tile][2] = *reinterpret_cast<const unsigned *>(lo_row + 16); a0[tile][3] = *reinterpret_cast<const unsigned *>(hi_row + 16);
        if (ksteps > 1) {
            a1[tile][0] = *reinterpret_cast<const unsigned *>(lo_row + 32); a1[tile][1] = *reinterpret_cast<const unsigned *>(hi_row + 32);
            a1[tile][2] = *reinterpret_cast<const unsigned *>(lo_row + 48); a1[tile][3] = *reinterpret_cast<const unsigned *>(hi_row + 48);
        }
    }
    int sum[4] = { 0, 0, 0, 0 };                        /* masked row sums: tile 0 rows g, g+8; tile 1 rows g, g+8 */
    const unsigned char *colbase = qd + (size_t)g * FPT_QD_STRIDE + 4 * t;
    const size_t dstride = (size_t)nrows * FPT_QD_STRIDE;
    for (int nt = 0; nt < ntiles; nt++) {
        int r0[4] = { 0, 0, 0, 0 }, r1[4] = { 0, 0, 0, 0 };
        for (int d = ndigits - 1; d >= 0; d--) {
            const unsigned char *row = colbase + (size_t)d * dstride + (size_t)(8 * nt) * FPT_QD_STRIDE;
            int c0[4] = { 0, 0, 0, 0 }, c1[4] = { 0, 0, 0, 0 };
            const unsigned b0 = *reinterpret_cast<const unsigned *>(row), b1 = *reinterpret_cast<const unsigned *>(row + 16);
            fpt_mma_u8(c0, a0[0], b0, b1);
            fpt_mma_u8(c1, a0[1], b0, b1);
            if (ksteps > 1) {
                const unsigned b2 = *reinterpret_cast<const unsigned *>(row + 32), b3 = *reinterpret_cast<const unsigned *>(row + 48);
                fpt_mma_u8(c0, a1[0], b2, b3);
                fpt_mma_u8(c1, a1[1], b2, b3);
            }
#pragma unroll
            for (int i = 0; i < 4; i++) { r0[i] = (r0[i] << 8) + c0[i]; r1[i] = (r1[i] << 8) + c1[i]; }
        }
        /* columns 8nt + 2t, +1 held by this lane: keep those outside the group (padded columns are zero anyway) */
        const unsigned char *mcol = myrows + 8 * nt + 2 * t;
        const unsigned m00 = *reinterpret_cast<const unsigned short *>(mcol);
        const unsigned m01 = *reinterpret_cast<const unsigned short *>(mcol + 8 * FPT_IND_STRIDE);
        const unsigned m10 = *reinterpret_cast<const unsigned short *>(mcol + 16 * FPT_IND_STRIDE);
        const unsigned m11 = *reinterpret_cast<const unsigned short *>(mcol + 24 * FPT_IND_STRIDE);
        if (!(m00 & 0x00ffu)) sum[0] += r0[0];
        if (!(m00 & 0xff00u)) sum[0] += r0[1];
        if (!(m01 & 0x00ffu)) sum[1] += r0[2];
        if (!(m01 & 0xff00u)) sum[1] += r0[3];
        if (!(m10 & 0x00ffu)) sum[2] += r1[0];
        if (!(m10 & 0xff00u)) sum[2] += r1[1];
        if (!(m11 & 0x00ffu)) sum[3] += r1[2];
        if (!(m11 & 0xff00u)) sum[3] += r1[3];
    }
#pragma unroll
    for (int r = 0; r < 4; r++) { sum[r] += __shfl_xor_sync(FPT_FULL_MASK, sum[r], 1); sum[r] += __shfl_xor_sync(FPT_FULL_MASK, sum[r], 2); }
    /* lane L owns row L%16 of tile L/16: that row sits in the quad of lanes 4*(L%8).. as sum[2*tile + (L%16 >= 8)] */
    const int src = 4 * (lane & 7);
    const int v0 = __shfl_sync(FPT_FULL_MASK, sum[0], src), v1 = __shfl_sync(FPT_FULL_MASK, sum[1], src);
    const int v2 = __shfl_sync(FPT_FULL_MASK, sum[2], src), v3 = __shfl_sync(FPT_FULL_MASK, sum[3], src);
    const int sel = ((lane >> 4) << 1) | ((lane >> 3) & 1);
    return sel == 0 ? v0 : (sel == 1 ? v1 : (sel == 2 ? v2 : v3));
}

__global__ void __launch_bounds__(256, 3)
fpt_css_perm2_kernel(const double *__restrict__ Xall, int m, int asize, int bsize, long long wbase, long long nwin,
                     const unsigned char *__restrict__ status, int treshold, int runs, uint64_t seed,
                     const uint64_t *__restrict__ state_override, int chain, int qbits,
                     double *__restrict__ out_score, double *__restrict__ out_p, int *__restrict__ out_hits,
                     int *__restrict__ out_n, unsigned long long *__restrict__ recheck_counter) {
    FPT_DYN_SMEM(smem);
    const int T = blockDim.x, tid = threadIdx.x, RS = fpt_perm_row_stride(m);
    size_t off = 0;
    double *dist = (double *)(smem + off); off += (size_t)m * m * 8;
    const int use_mma = fpt_css_perm2_uses_mma(m);
    const int ndigits = (qbits + 8) >> 3;                  /* q <= 2^qbits */
    const int qd_rows = ((m + 7) >> 3) << 3;
    unsigned char *qd = smem + off; if (use_mma) off += (size_t)3 * qd_rows * FPT_QD_STRIDE;
    unsigned *q = (unsigned *)(smem + off); off += (size_t)m * m * 4;
    int *rowsum = (int *)(smem + off); off += (size_t)m * 4;
    off = (off + 7) & ~(size_t)7;
    uint2 *rtab = (uint2 *)(smem + off); off += (size_t)(m + 1) * 8;
    double *X = (double *)(smem + off); off += (size_t)2 * m * 8;
    int *offs = (int *)(smem + off); off += (size_t)T * 4;
    int *usedv = (int *)(smem + off); off += (size_t)T * 4;
    int *scan = (int *)(smem + off); off += 33 * 4 + 12;
    off = (off + 15) & ~(size_t)15;
    unsigned char *carry = smem + off; off += (size_t)RS;
    off = (off + 15) & ~(size_t)15;
    unsigned char *rows = smem + off;
    unsigned char *rows2 = rows + (size_t)T * RS;          /* chain mode only */
    unsigned char *ind = smem + ((off + (size_t)T * RS * (chain ? 2 : 1) + 15) & ~(size_t)15);   /* MMA path only */
    unsigned char *myind = ind + (size_t)tid * FPT_IND_STRIDE, *warpind = ind + (size_t)(tid & ~31) * FPT_IND_STRIDE;
    /* independent shuffles: permutation k of a window starts k (m-1) draws into the window's stream. Thread t's first
       permutation of chunk c is (c T + t) PP, so its start is  skip_t( skip_chunk^c (window state) )  with two affine maps
       x -> a x + b (mod 2^48) that do not depend on the window: computed once here, applied with two multiplies per chunk */
    ulonglong2 *skipmap = (ulonglong2 *)(smem + (((size_t)(ind - smem) + (use_mma ? (size_t)T * FPT_IND_STRIDE : 0) + 15) & ~(size_t)15));
    __shared__ double s_score, s_dmax;
    __shared__ int s_flag;
    unsigned char *mine = rows + (size_t)tid * RS;
    const int use_a = asize <= bsize;
    const int draws = m - 1;
    unsigned long long rechecks = 0;
    {
        /* a = A^n, b = skip(0, n): skip(x, n) = a x + b */
        const uint64_t n_t = (uint64_t)tid * FPT_PERM_PP * (uint64_t)(m - 1);
        const uint64_t b_t = fpt_lcg_skip(0ULL, n_t);
        skipmap[tid] = make_ulonglong2((fpt_lcg_skip(1ULL, n_t) - b_t) & FPT_MASK48, b_t);
        if (tid == 0) {
            const uint64_t n_c = (uint64_t)T * FPT_PERM_PP * (uint64_t)(m - 1);
            const uint64_t b_c = fpt_lcg_skip(0ULL, n_c);
            skipmap[T] = make_ulonglong2((fpt_lcg_skip(1ULL, n_c) - b_c) & FPT_MASK48, b_c);
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
        double dmax = 0.0;
        for (int e = tid; e < m * m; e += T) {              /* calc_dist, css.c:573-587 */
            const int i = e / m, j = e - i * m;
            if (j < i) {
                const double dx = __dsub_rn(X[2 * i], X[2 * j]), dy = __dsub_rn(X[2 * i + 1], X[2 * j + 1]);
                const double d = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                dist[e] = d; dist[j * m + i] = d;
                dmax = fmax(dmax, d);                       /* NaN distances: fmax ignores them, see `usable` */
            } else if (j == i) dist[e] = 0.0;
        }
        for (int e = tid; e < RS; e += T) carry[e] = (unsigned char)e;
        for (int o = 16; o > 0; o >>= 1) dmax = fmax(dmax, __shfl_xor_sync(FPT_FULL_MASK, dmax, o));
        double *wmax = reinterpret_cast<double *>(offs);    /* T ints = T/2 doubles >= T/32 warp maxima; 8-byte aligned */
        if ((tid & 31) == 0) wmax[tid >> 5] = dmax;
        __syncthreads();
        if (tid == 0) {
            double mx = 0.0;
            for (int k = 0; k < (T >> 5); k++) mx = fmax(mx, wmax[k]);
            s_dmax = mx;
            s_score = fpt_css_score<unsigned char>(dist, m, carry, carry + asize, asize, bsize);
        }
        __syncthreads();
        const double score = s_score;
        dmax = s_dmax;
        /* quantisation: q = rint(d * S), S = 2^qbits / dmax  =>  q <= 2^qbits, |q/S - d| <= 0.5/S.
           The surrogate is only usable for finite, positive distances and a finite observed score. */
        const bool usable = (dmax > 0.0) && (dmax < 1e300) && (score == score) && (fabs(score) < 1e300);
        const double S = usable ? (double)(1u << qbits) / dmax : 0.0;
        int bad = 0;
        for (int e = tid; e < m * m; e += T) {
            const double d = dist[e];
            if (!(d == d)) bad = 1;
            q[e] = usable && d == d ? (unsigned)__double2ll_rn(d * S) : 0u;
        }
        const bool use_surrogate = usable && !__syncthreads_or(bad);
        if (use_mma) {                                      /* base-256 digits of q, zero padded to 8-row / 64-column tiles */
            for (int e = tid; e < ndigits * qd_rows * 16; e += T) {
                const int d = e / (qd_rows * 16), rem = e - d * qd_rows * 16, n = rem >> 4, k4 = (rem & 15) << 2;
                unsigned wv = 0;
                for (int b = 0; b < 4; b++) {
                    const int k = k4 + b;
                    const unsigned v = (n < m && k < m) ? ((q[n * m + k] >> (8 * d)) & 0xffu) : 0u;
                    wv |= v << (8 * b);
                }
                *reinterpret_cast<unsigned *>(qd + ((size_t)d * qd_rows + n) * FPT_QD_STRIDE + k4) = wv;
            }
        }
        for (int i = tid; i < m; i += T) {
            int s = 0;
            for (int j = 0; j < m; j++) s += (int)q[i * m + j];
            rowsum[i] = s;
        }
        __syncthreads();
        /* |surrogate - reference score| <= E:  quantisation  (0.5/S)(1 + (a+b)(1/a^2 + 1/b^2))  on the three
           means, plus a generous bound on the fp64 rounding of both evaluations */
        const double a_ = (double)asize, b_ = (double)bsize;
        const double wterm = (asize > 1 ? 1.0 / (a_ * a_) : 0.0) + (bsize > 1 ? 1.0 / (b_ * b_) : 0.0);
        const double E = use_surrogate ? (0.5 / S) * (1.0 + (a_ + b_) * wterm) * 1.0000001 + 1e-11 * dmax * (1.0 + (a_ + b_)) : 0.0;
        const double invS = use_surrogate ? 1.0 / S : 0.0;
        const double c_bet = invS / (a_ * b_);
        const double c_wa = asize > 1 ? invS / (a_ * a_ * (a_ - 1.0)) : 0.0;
        const double c_wb = bsize > 1 ? invS / (b_ * b_ * (b_ - 1.0)) : 0.0;

        const uint64_t st_win = state_override ? state_override[w] : fpt_stream_state(seed, wbase + w, FPT_STREAM_RESAMPLE);
        long long stream_pos = 0;                           /* chain mode: draws consumed by finished chunks */
        uint64_t st_chunk = st_win;                         /* independent mode: window state advanced by the finished chunks */
        int hits = 0, ndone = 0;
        bool stopped = false;
        while (!stopped && hits < treshold && ndone < runs) {
            const int nvalid = min(T * FPT_PERM_PP, runs - ndone);
            const int first = tid * FPT_PERM_PP;            /* my permutations: first .. first+PP-1 of this chunk */
            const int mycount = max(0, min(FPT_PERM_PP, nvalid - first));
            uint64_t st = 0;
            if (chain) {
                /* pass 1: net permutation of my block from identity; repair stream offsets after rejections */
                offs[tid] = first * draws;
                __syncthreads();
                for (;;) {
                    int used = 0;
                    if (mycount > 0) {
                        fpt_identity_row(mine, m);
                        uint64_t s1 = fpt_lcg_skip(st_win, (uint64_t)(stream_pos + offs[tid]));
                        for (int j = 0; j < mycount; j++) fpt_shuffle_row(mine, m, rtab, s1, used);
                    }
                    int total = 0;
                    const int incl = fpt_block_scan_incl(used, scan, &total);
                    const int want = incl - used;
                    const int wrong = (mycount > 0 && want != offs[tid]) ? 1 : 0;
                    if (wrong) offs[tid] = want;
                    usedv[tid] = total;
                    if (!__syncthreads_or(wrong)) break;
                }
                /* exclusive scan under composition over the thread blocks: (f o g)[pos] = f[g[pos]] */
                const int nblk = (nvalid + FPT_PERM_PP - 1) / FPT_PERM_PP;
                unsigned char *src = rows, *dst = rows2;
                for (int d = 1; d < nblk; d <<= 1) {
                    __syncthreads();
                    if (tid < nblk) {
                        const unsigned char *g = src + (size_t)tid * RS;
                        unsigned char *o = dst + (size_t)tid * RS;
                        if (tid >= d) {
                            const unsigned char *f = src + (size_t)(tid - d) * RS;
                            for (int e = 0; e < m; e++) o[e] = f[g[e]];
                        } else {
                            for (int e = 0; e < m; e++) o[e] = g[e];
                        }
                    }
                    unsigned char *tmp = src; src = dst; dst = tmp;
                }
                __syncthreads();
                /* my starting labels: carry o (inclusive scan up to block tid-1); my block is replayed from there */
                unsigned char *start = dst + (size_t)tid * RS;
                if (mycount > 0) {
                    if (tid == 0) { for (int e = 0; e < m; e++) start[e] = carry[e]; }
                    else { const unsigned char *g = src + (size_t)(tid - 1) * RS; for (int e = 0; e < m; e++) start[e] = carry[g[e]]; }
                }
                __syncthreads();
                mine = start;
                st = fpt_lcg_skip(st_win, (uint64_t)(stream_pos + offs[tid]));
            }
            /* pass 2 (the only pass without the chain): shuffle, score, decide */
            int myhits = 0, hitmask = 0;
            bool resync = false;
            /* warps whose 32 x PP permutations all lie beyond the chunk skip the loop; inside, every lane runs every
               iteration because the tensor-core path is a warp-wide operation (idle lanes contribute empty masks) */
            const bool warp_active = (tid & ~31) * FPT_PERM_PP < nvalid;
            for (int j = 0; warp_active && j < FPT_PERM_PP; j++) {
                const bool valid = j < mycount;
                int used = 0;
                if (valid) {
                    if (!chain) {
                        fpt_identity_row(mine, m);
                        /* the stream of permutation k starts k*(m-1) draws in; after a shuffle without a rejected draw the
                           state already sits at the next permutation's start */
                        if (j == 0) { const ulonglong2 sk = skipmap[tid]; st = (sk.x * st_chunk + sk.y) & FPT_MASK48; }
                        else if (resync) st = fpt_lcg_skip(st_win, (uint64_t)(ndone + first + j) * (uint64_t)draws);
                        resync = !fpt_shuffle_row_optimistic(mine, m, rtab, st);
                        if (resync) {                       /* a rejected draw: replay this permutation exactly */
                            fpt_identity_row(mine, m);
                            fpt_shuffle_row(mine, m, rtab, st, used);
                        }
                    } else {
                        fpt_shuffle_row(mine, m, rtab, st, used);
                    }
                }
                int hit = 0;
                bool exact = valid && !use_surrogate;
                if (use_surrogate) {
                    long long bet = 0; int wa = 0, wb = 0;
                    if (use_mma) {
                        /* one walk over the labels: membership row of the smaller group + both adjacent-pair sums */
                        fpt_ind_row_clear(myind);
                        if (valid) {
                            int prev = mine[0];
                            if (use_a) myind[prev] = 1;
                            for (int i = 1; i < asize; i++) {
                                const int c = mine[i];
                                if (use_a) myind[c] = 1;
                                wa += (int)q[prev * m + c]; prev = c;
                            }
                            prev = mine[asize];
                            if (!use_a) myind[prev] = 1;
                            for (int i = 1; i < bsize; i++) {
                                const int c = mine[asize + i];
                                if (!use_a) myind[c] = 1;
                                wb += (int)q[prev * m + c]; prev = c;
                            }
                        }
                        __syncwarp();
                        bet = (long long)fpt_bet_mma(warpind, qd, m, ndigits);
                        __syncwarp();
                    } else if (valid) {
                        fpt_surrogate(q, rowsum, m, mine, asize, bsize, use_a, bet, wa, wb);
                    }
                    if (valid) {
                        const double approx = (double)bet * c_bet - (a_ + b_) * ((double)wa * c_wa + (double)wb * c_wb);
                        const double diff = approx - score;
                        hit = diff > 0.0;
                        exact = !(fabs(diff) > E);
                    }
                }
                if (exact) {
                    hit = fpt_css_score<unsigned char>(dist, m, mine, mine + asize, asize, bsize) >= score ? 1 : 0;
                    rechecks++;
                }
                myhits += hit; hitmask |= hit << j;
            }
            int chunk_hits = 0;
            const int hincl = fpt_block_scan_incl(myhits, scan, &chunk_hits);
            if (tid == 0) s_flag = -1;
            __syncthreads();
            if (myhits > 0 && hits + hincl >= treshold && hits + hincl - myhits < treshold) {
                /* the treshold-th hit is one of mine: find which permutation */
                int need = treshold - (hits + hincl - myhits);
                for (int j = 0; j < FPT_PERM_PP; j++) if ((hitmask >> j) & 1) { if (--need == 0) { s_flag = first + j; break; } }
            }
            __syncthreads();
            if (s_flag >= 0) {
                ndone += s_flag + 1; hits = treshold; stopped = true;
            } else {
                hits += chunk_hits; ndone += nvalid;
                { const ulonglong2 sk = skipmap[T]; st_chunk = (sk.x * st_chunk + sk.y) & FPT_MASK48; }
                if (chain) {
                    const int lastblk = (nvalid - 1) / FPT_PERM_PP;
                    if (tid == lastblk) for (int e = 0; e < m; e++) carry[e] = mine[e];
                    stream_pos += usedv[0];
                }
            }
            __syncthreads();
            mine = rows + (size_t)tid * RS;
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
