/*
 * fpt_css_perm3.cuh — Monte-Carlo permutation test for the cohorts the genome scans are made of (8 <= m <= 64, independent
 * shuffles): the headline kernel of BASELINE configs[2]. Same decisions as fpt_css_perm2_kernel (fpt_css_perm.cuh): reference
 * Fisher-Yates from the window's nrand48 stream (css/css.c:700-706), every permutation scored by the exact integer surrogate on
 * the u8 tensor cores, the rare permutation within the proven bound of the observed score re-scored in the reference's summation
 * order (css.c:608-647), early stop and p as css.c:727-752. The observed score itself comes from fpt_css_observed_kernel (one
 * warp per window at full occupancy: it is a chain of asize*bsize dependent fp64 additions, and inside this kernel seven warps
 * of eight waited for it — 20 % of all warp samples in the ncu capture of the previous form, profiles/r2_perm3_before.md).
 *
 * How the work meets the SM (ncu of the round-1 kernel: 90 warp instructions per permutation, 38 % of the shared-memory
 * wavefronts bank-conflict replays of the random label swaps, a third of the warp samples waiting at CTA barriers):
 *
 *   - 128 threads per CTA, FOUR permutations per thread in flight: their LCG chains, draws and swaps are independent, so the
 *     fixed-latency chains hide each other; the per-step table load and loop overhead are paid once for four;
 *   - labels live in a [position][thread] layout, one 32-bit word per thread and position, byte k = permutation k of the thread:
 *     the address of position idx is one multiply-add and, whatever index a draw picks, the 32 lanes of a warp touch 32
 *     different banks — the random swaps are conflict-free by construction. The words are private to their thread: the reset to
 *     identity needs no barrier, and the four labels at the position a step finalises come with ONE load;
 *   - a step stores only the label that moves down (position i is final and never read again by the shuffle): membership mask
 *     and adjacent-pair sums are read off the swaps as they happen. The rare permutation that needs its labels (exact re-score,
 *     rejected draw) is regenerated alone from its stream position;
 *   - draws: the 48-bit LCG on a (32 high, 16 low) split, r mod n by an exact 32-bit magic quotient (no fix-up), the rejection
 *     test as one running maximum (r <= 2^31 - 64 is accepted for every n <= 64);
 *   - between-group sums: 32 permutations x m individuals x 3 base-256 digits per warp on mma.sync m16n8k32 (u8 x u8 -> s32)
 *     with the contraction index PERMUTED so that each lane's A and B fragments of both k-steps are 16 contiguous bytes: one
 *     128-bit load per membership row and one per (column tile, digit); the column mask comes from the membership bit masks by
 *     four shuffles, not from shared memory; the per-row sums are reduced by a 3-shuffle reduce-scatter;
 *   - one CTA barrier per round, and only in rounds that can stop early or end the window.
 */
#ifndef FPT_CSS_PERM3_CUH
#define FPT_CSS_PERM3_CUH

#include "fpt_css_perm.cuh"

#define FPT_P3_T 128                      /* threads per CTA */
#define FPT_P3_PP 4                       /* permutations per thread in flight */
#define FPT_P3_ROUND (FPT_P3_PP * FPT_P3_T)
#define FPT_P3_LINE (4 * FPT_P3_T)        /* bytes per label position: one word per thread */
#define FPT_P3_WARPS (FPT_P3_T / 32)

FPT_HD int fpt_css_perm3_ok(int m, int chain) { return !chain && m >= 8 && m <= 64; }
FPT_HD int fpt_css_perm3_ksteps(int m) { return m <= 32 ? 1 : 2; }

FPT_HD size_t fpt_css_perm3_smem_bytes(int m) {
    const int qd_rows = ((m + 7) >> 3) << 3, kb = 32 * fpt_css_perm3_ksteps(m);
    size_t off = (size_t)2 * m * 8;                                   /* X */
    off += ((size_t)m * m * 4 + 7) & ~(size_t)7;                      /* q (odd m: the 8-byte table behind it stays aligned) */
    off += (size_t)(m + 1) * 8;                                       /* (magic, shift) per n */
    off = (off + 15) & ~(size_t)15;
    off += (size_t)3 * qd_rows * kb;                                  /* digit matrices */
    off += (size_t)m * FPT_P3_LINE;                                   /* labels */
    off += (size_t)FPT_P3_WARPS * 32 * 8;                             /* exact re-scoring: 32 staged distances per warp */
    off += 2 * 2 * FPT_P3_WARPS * 4 + 16;                             /* per-warp hit counts (double-buffered), stop flag */
    return off;
}

/* exact quotient of a 31-bit draw by n (2 <= n <= 64): floor(r / n) = umulhi(r, M) >> sh with M = ceil(2^(31+s) / n),
   s = ceil(log2 n), sh = s - 1. M fits 32 bits (n > 2^(s-1), or n = 2^s and M = 2^31); the error term r e / (n 2^(31+s)),
   e = M n - 2^(31+s) < n <= 2^s, stays below 1/n for r < 2^31, so the floor is exact (tests/test_emu_kernels.py sweeps it). */
FPT_HD uint2 fpt_p3_magic(uint32_t n) {
    uint2 r; r.x = 0u; r.y = 0u;
    if (n < 2) return r;
    uint32_t s = 0;
    while ((1u << s) < n) s++;
    const unsigned long long p = 1ULL << (31 + s);
    r.x = (uint32_t)((p + n - 1) / n);
    r.y = s - 1;
    return r;
}

/* X <- A X + C mod 2^48 on the split state (hi = bits 16..47, lo = bits 0..15); returns nrand48's 31-bit draw (bits 17..47) */
FPT_D uint32_t fpt_p3_lcg(uint32_t &hi, uint32_t &lo) {
    const uint32_t t0 = lo * 0xE66Du + 0xBu;                          /* < 2^32 */
    hi = hi * 0xDEECE66Du + (lo * 0x5DEECu + (t0 >> 16));
    lo = t0 & 0xffffu;
    return hi >> 1;
}

/* the four permutations of a thread */
template <bool WIDE> struct FptP3Mask { typedef uint32_t type; };
template <> struct FptP3Mask<true> { typedef unsigned long long type; };

template <bool WIDE>
struct FptP3State {
    uint32_t h[FPT_P3_PP], l[FPT_P3_PP];             /* split LCG states */
    typename FptP3Mask<WIDE>::type mask[FPT_P3_PP];  /* membership of the smaller group, bit = individual */
    int acc[FPT_P3_PP];                              /* running adjacent-pair sum of quantised distances */
    uint32_t prev[FPT_P3_PP];                        /* label at the position above */
    uint32_t maxr;                                   /* largest draw so far (rejection test) */
};

/* one Fisher-Yates step of all four permutations: position i (n = i + 1) becomes final. `word_i` = my label word at position i,
   `col0` = my word at position 0. MEMBER: position i belongs to the smaller group, so the label landing there joins the
   membership mask; PAIR: the adjacent pair (i, i + 1) counts, its quantised distance is added to the running sum. */
template <bool MEMBER, bool PAIR, bool WIDE>
FPT_D void fpt_p3_step(FptP3State<WIDE> &p, uint32_t n, const uint2 mg, const unsigned char *word_i, unsigned char *col0,
                       const unsigned *q, int m) {
    const uint32_t aw = *reinterpret_cast<const uint32_t *>(word_i);
#pragma unroll
    for (int k = 0; k < FPT_P3_PP; k++) {
        const uint32_t r = fpt_p3_lcg(p.h[k], p.l[k]);
        p.maxr = max(p.maxr, r);
        const uint32_t rem = r - (__umulhi(r, mg.x) >> mg.y) * n;
        unsigned char *pr = col0 + rem * FPT_P3_LINE + k;
        const uint32_t a = (aw >> (8 * k)) & 0xffu;
        const uint32_t c = *pr;                                      /* rem == i: reads a itself and stores it back */
        *pr = (unsigned char)a;
        if (MEMBER) p.mask[k] |= (typename FptP3Mask<WIDE>::type)1 << c;
        if (PAIR) p.acc[k] += (int)q[c * m + p.prev[k]];
        p.prev[k] = c;
    }
}

/* positions [hi, lo] (descending) */
template <bool MEMBER, bool PAIR, bool WIDE>
FPT_D void fpt_p3_run(FptP3State<WIDE> &p, int hi, int lo, const uint2 *rtab, unsigned char *col0, const unsigned *q, int m) {
    const unsigned char *w = col0 + (unsigned)hi * FPT_P3_LINE;
#pragma unroll 2
    for (int i = hi; i >= lo; i--) {
        fpt_p3_step<MEMBER, PAIR, WIDE>(p, (uint32_t)(i + 1), rtab[i + 1], w, col0, q, m);
        w -= FPT_P3_LINE;
    }
}

/* one permutation regenerated on its own, on the exact path (rejections honoured), with every store: byte column `col`
   (position idx at col + idx * FPT_P3_LINE) holds the final labels afterwards */
FPT_D void fpt_p3_regen(unsigned char *col, int m, uint64_t st) {
    for (int e = 0; e < m; e++) col[(unsigned)e * FPT_P3_LINE] = (unsigned char)e;
    int used = 0;
    for (int i = m - 1; i > 0; i--) {
        const int rr = (int)fpt_randint((uint32_t)(i + 1), st, used);
        unsigned char *pi = col + (unsigned)i * FPT_P3_LINE, *pr = col + (unsigned)rr * FPT_P3_LINE;
        const unsigned char t = *pi; *pi = *pr; *pr = t;
    }
}

/* reference-order sums (css.c:608-647) over the labels in byte column `col` by a whole warp: the lanes compute the next 32
   distances (calc_dist, css.c:573-587, from the embedding: the same expression, hence the same bits) while lane 0 adds the
   previous 32 in order. kind 0: between-group pairs, 1 / 2: adjacent pairs of the first / second group. Lane 0 holds the sum. */
FPT_D double fpt_p3_chain(const double *X, double *stage, int kind, int asize, int bsize, int lane, const unsigned char *col) {
    const int total = kind == 0 ? asize * bsize : (kind == 1 ? asize - 1 : bsize - 1);
    double acc = 0.0;
    auto dist_of = [&](int e) -> double {
        int i, j;
        if (kind == 0) { const int r = e / bsize; i = asize - 1 - r; j = asize + bsize - 1 - (e - r * bsize); }
        else if (kind == 1) { i = asize - 2 - e; j = i + 1; }
        else { i = asize + bsize - 2 - e; j = i + 1; }
        i = col[(unsigned)i * FPT_P3_LINE]; j = col[(unsigned)j * FPT_P3_LINE];
        const double dx = __dsub_rn(X[2 * i], X[2 * j]), dy = __dsub_rn(X[2 * i + 1], X[2 * j + 1]);
        return __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
    };
    double dcur = lane < total ? dist_of(lane) : 0.0;
    for (int base = 0; base < total; base += 32) {
        stage[lane] = dcur;
        __syncwarp();
        const int nxt = base + 32 + lane;
        const double dnext = nxt < total ? dist_of(nxt) : 0.0;
        if (lane == 0) {
            const int cnt = total - base < 32 ? total - base : 32;
            if (cnt == 32) {
#pragma unroll
                for (int t = 0; t < 32; t++) acc = __dadd_rn(acc, stage[t]);
            } else {
                for (int t = 0; t < cnt; t++) acc = __dadd_rn(acc, stage[t]);
            }
        }
        __syncwarp();
        dcur = dnext;
    }
    return acc;
}

/* css() of css.c:608-647 for the labels in byte column `col`, by the whole warp; every lane returns the score */
FPT_D double fpt_p3_warp_exact_score(const double *X, double *stage, const unsigned char *col, int asize, int bsize, int lane) {
    double bet = fpt_p3_chain(X, stage, 0, asize, bsize, lane, col);
    const double wa0 = asize > 1 ? fpt_p3_chain(X, stage, 1, asize, bsize, lane, col) : 0.0;
    const double wb0 = bsize > 1 ? fpt_p3_chain(X, stage, 2, asize, bsize, lane, col) : 0.0;
    bet = __ddiv_rn(bet, (double)((long long)asize * bsize));
    const double wa = asize > 1 ? __ddiv_rn(wa0, (double)((long long)asize * asize * (asize - 1))) : 0.0;
    const double wb = bsize > 1 ? __ddiv_rn(wb0, (double)((long long)bsize * bsize * (bsize - 1))) : 0.0;
    return __shfl_sync(FPT_FULL_MASK, __dsub_rn(bet, __dmul_rn((double)(asize + bsize), __dadd_rn(wa, wb))), 0);
}

/* ------------------------------------------------------------------------------------------------------------------------
 * Between-group sums of the warp's 32 permutations (one per lane) on the u8 tensor cores.
 *
 * R = Z Q_d per base-256 digit d, Z (32 x K) the 0/1 membership rows of the smaller group, then sum_j (1 - z_j) R_j.
 * mma.sync.m16n8k32 (g = lane / 4, t = lane % 4): A regs a0/a2 = row g, k-slots 4t..4t+3 / 16+4t..; a1/a3 = row g + 8;
 * B regs b0/b1 = the same k-slots of column g; C c0,c1 = row g, columns 2t, 2t+1; c2,c3 = row g + 8.
 * The hardware only requires that A and B agree on which individual sits in which k-slot, so slot (k-step s, register half h,
 * byte b) of lane t is individual 8 KS t + 8 s + 4 h + b: the 8 KS bytes a lane needs from a digit row are contiguous (one
 * 128-bit load per column tile and digit at KS = 2, one 64-bit load at KS = 1), and the 8 KS membership bytes of a row are 8 KS
 * consecutive BITS of that permutation's mask: the A fragments are expanded from the masks in registers, straight into the
 * register quads the MMA reads — no shared-memory round trip, no fragment shuffling. (A first form staged the rows through
 * shared memory with one 128-bit load per row: the loaded registers were in (row, row) order, the MMA wants (row, row + 8)
 * interleaved, and the 240 register moves per call that followed were a third of the kernel's instructions.)
 * `mk` = this lane's membership mask (bit j set: individual j belongs to the smaller group of this lane's permutation).
 */
FPT_D unsigned fpt_p3_nibble_bytes(unsigned x) { return ((x & 0xfu) * 0x00204081u) & 0x01010101u; }   /* bit i -> byte i */

template <int KS>
FPT_D long long fpt_p3_bet_mma(const unsigned char *qd, int ntiles, int qd_rows, typename FptP3Mask<KS == 2>::type mk) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    constexpr int KB = 32 * KS;
    /* masks of my four rows (permutations g + 8 s: rows g, g + 8 of tile 0, then of tile 1) */
    typename FptP3Mask<KS == 2>::type mr[4];
#pragma unroll
    for (int s = 0; s < 4; s++) mr[s] = __shfl_sync(FPT_FULL_MASK, mk, g + 8 * s);
    /* A fragments: my 8 KS individuals of each row */
    unsigned a[2][KS][4];
#pragma unroll
    for (int s = 0; s < 4; s++) {
        const unsigned bits = (unsigned)(mr[s] >> (8 * KS * t));     /* 8 KS valid bits */
        const int tile = s >> 1, j = s & 1;
#pragma unroll
        for (int ks = 0; ks < KS; ks++) {
            a[tile][ks][j] = fpt_p3_nibble_bytes(bits >> (8 * ks));
            a[tile][ks][2 + j] = fpt_p3_nibble_bytes(bits >> (8 * ks + 4));
        }
    }
    /* column masks: bit 8 nt (+1) of nm[s] is column 8 nt + 2 t (+1) of row s; set = outside the group */
    typename FptP3Mask<KS == 2>::type nm[4];
#pragma unroll
    for (int s = 0; s < 4; s++) nm[s] = ~mr[s] >> (2 * t);
    int sum[4] = { 0, 0, 0, 0 };                        /* every partial sum is a part of sum_{G x not G} q < 2^31 */
    const unsigned char *colbase = qd + (size_t)g * KB + 8 * KS * t;
    const size_t dstride = (size_t)qd_rows * KB;
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
        if (nt < ntiles) {
            int c[3][2][4];
#pragma unroll
            for (int d = 0; d < 3; d++) {
                const unsigned char *src = colbase + (size_t)d * dstride + (size_t)(8 * nt) * KB;
                unsigned b[2 * KS];
                if (KS == 2) { const uint4 v = *reinterpret_cast<const uint4 *>(src); b[0] = v.x; b[1] = v.y; b[2 * KS - 2] = v.z; b[2 * KS - 1] = v.w; }
                else { const uint2 v = *reinterpret_cast<const uint2 *>(src); b[0] = v.x; b[1] = v.y; }
#pragma unroll
                for (int tile = 0; tile < 2; tile++) {
#pragma unroll
                    for (int e = 0; e < 4; e++) c[d][tile][e] = 0;
#pragma unroll
                    for (int ks = 0; ks < KS; ks++) fpt_mma_u8(c[d][tile], a[tile][ks], b[2 * ks], b[2 * ks + 1]);
                }
            }
            /* the three digits of R, then the columns outside the group */
#pragma unroll
            for (int tile = 0; tile < 2; tile++) {
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int s = 2 * tile + (e >> 1);
                    const int v = (c[2][tile][e] * 256 + c[1][tile][e]) * 256 + c[0][tile][e];   /* < min(a, b) * 2^(qbits + 1) < 2^31 */
                    if ((nm[s] >> (8 * nt + (e & 1))) & 1) sum[s] += v;
                }
            }
        }
    }
    /* reduce over the four lanes of a quad so that lane t keeps row s = t (reduce-scatter), then hand every lane its own row */
    const bool b0 = (t & 1) != 0, b1 = (t & 2) != 0;
    const int u0 = (b0 ? sum[1] : sum[0]) + __shfl_xor_sync(FPT_FULL_MASK, b0 ? sum[0] : sum[1], 1);
    const int u1 = (b0 ? sum[3] : sum[2]) + __shfl_xor_sync(FPT_FULL_MASK, b0 ? sum[2] : sum[3], 1);
    const int mine = (b1 ? u1 : u0) + __shfl_xor_sync(FPT_FULL_MASK, b1 ? u0 : u1, 2);
    return (long long)__shfl_sync(FPT_FULL_MASK, mine, 4 * (lane & 7) + (lane >> 3));
}

template <int KS>
__global__ void __launch_bounds__(FPT_P3_T, 4)
fpt_css_perm3_kernel(const double *__restrict__ Xall, int m, int asize, int bsize, long long wbase, long long nwin,
                     const unsigned char *__restrict__ status, int treshold, int runs, uint64_t seed,
                     const uint64_t *__restrict__ state_override, int qbits, const double *__restrict__ obs_score,
                     double *__restrict__ out_p, int *__restrict__ out_hits, int *__restrict__ out_n,
                     unsigned long long *__restrict__ recheck_counter) {
    FPT_DYN_SMEM(smem);
    constexpr bool WIDE = KS == 2;
    constexpr int KB = 32 * KS;
    typedef typename FptP3Mask<WIDE>::type mask_t;
    const int T = FPT_P3_T, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int qd_rows = ((m + 7) >> 3) << 3, ntiles = qd_rows >> 3;
    size_t off = 0;
    double *X = (double *)(smem + off); off += (size_t)2 * m * 8;
    unsigned *q = (unsigned *)(smem + off); off += ((size_t)m * m * 4 + 7) & ~(size_t)7;
    uint2 *rtab = (uint2 *)(smem + off); off += (size_t)(m + 1) * 8;
    off = (off + 15) & ~(size_t)15;
    unsigned char *qd = smem + off; off += (size_t)3 * qd_rows * KB;
    unsigned char *labels = smem + off; off += (size_t)m * FPT_P3_LINE;
    double *stage = reinterpret_cast<double *>(smem + off) + warp * 32; off += (size_t)FPT_P3_WARPS * 32 * 8;
    int *whits = (int *)(smem + off); off += 2 * 2 * FPT_P3_WARPS * 4;         /* [parity][0: round hits, 1: pending hits][warp] */
    int *s_flag = (int *)(smem + off);
    unsigned char *col0 = labels + 4 * tid;                                    /* my label word at position 0 */
    const int use_a = asize <= bsize;
    const int draws = m - 1;
    unsigned long long rechecks = 0;
    /* affine maps x -> a x + b (mod 2^48) of the stream: to my first permutation of a round, over one permutation, over a round */
    uint64_t map_t_a, map_t_b, map_1_a, map_1_b, map_r_a, map_r_b;
    {
        const uint64_t n_t = (uint64_t)tid * FPT_P3_PP * (uint64_t)draws;
        map_t_b = fpt_lcg_skip(0ULL, n_t); map_t_a = (fpt_lcg_skip(1ULL, n_t) - map_t_b) & FPT_MASK48;
        map_1_b = fpt_lcg_skip(0ULL, (uint64_t)draws); map_1_a = (fpt_lcg_skip(1ULL, (uint64_t)draws) - map_1_b) & FPT_MASK48;
        const uint64_t n_r = (uint64_t)FPT_P3_ROUND * (uint64_t)draws;
        map_r_b = fpt_lcg_skip(0ULL, n_r); map_r_a = (fpt_lcg_skip(1ULL, n_r) - map_r_b) & FPT_MASK48;
    }
    for (int n = tid; n <= m; n += T) rtab[n] = fpt_p3_magic((uint32_t)n);
    int parity = 0;
    __syncthreads();

    for (long long w = blockIdx.x; w < nwin; w += gridDim.x) {
        if (status[w] != FPT_WIN_SCORED) continue;
        for (int e = tid; e < 2 * m; e += T) X[e] = Xall[(size_t)w * 2 * m + e];
        const double score = obs_score[w];
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
        /* quantised distances q = rint(d S) <= 2^qbits (|q/S - d| <= 0.5/S), symmetric, zero diagonal */
        for (int i = warp; i < m; i += FPT_P3_WARPS) {
            const double xi = X[2 * i], yi = X[2 * i + 1];
            for (int j = lane; j <= i; j += 32) {
                unsigned qv = 0u;
                if (j < i) {
                    const double dx = __dsub_rn(xi, X[2 * j]), dy = __dsub_rn(yi, X[2 * j + 1]);
                    const double d = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
                    qv = (scale_ok && d == d) ? (unsigned)__double2ll_rn(d * S) : 0u;
                }
                q[i * m + j] = qv; q[j * m + i] = qv;
            }
        }
        __syncthreads();
        const bool use_surrogate = scale_ok && (score == score) && (fabs(score) < 1e300);
        /* base-256 digits of q: row n = individual n, byte k = individual k, zero padded to 8-row tiles and KB columns */
        for (int e = tid; e < 3 * qd_rows * (KB / 4); e += T) {
            const int d = e / (qd_rows * (KB / 4)), rem = e - d * qd_rows * (KB / 4), n = rem / (KB / 4), k4 = (rem - n * (KB / 4)) << 2;
            unsigned wv = 0;
            if (n < m) {
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const int k = k4 + b;
                    const unsigned v = k < m ? ((q[n * m + k] >> (8 * d)) & 0xffu) : 0u;
                    wv |= v << (8 * b);
                }
            }
            *reinterpret_cast<unsigned *>(qd + ((size_t)d * qd_rows + n) * KB + k4) = wv;
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
        int hits = 0, ndone = 0, pending = 0;               /* pending: my hits of rounds that ended without a barrier */
        bool stopped = false;
        __syncthreads();
        while (!stopped && hits < treshold && ndone < runs) {
            const int nvalid = min(FPT_P3_ROUND, runs - ndone);
            const int first = FPT_P3_PP * tid;              /* my permutations of this round: first .. first + 3 */
            const int mycount = max(0, min(FPT_P3_PP, nvalid - first));
            const bool warp_active = (tid & ~31) * FPT_P3_PP < nvalid;
            unsigned hitmask = 0u;
            if (warp_active) {
                /* ---- identity labels (my own words), then the four shuffles at once */
                for (int e = 0; e < m; e++) *reinterpret_cast<uint32_t *>(col0 + (unsigned)e * FPT_P3_LINE) = (uint32_t)e * 0x01010101u;
                FptP3State<WIDE> pp;
                {
                    uint64_t s = (map_t_a * st_round + map_t_b) & FPT_MASK48;
#pragma unroll
                    for (int k = 0; k < FPT_P3_PP; k++) {
                        pp.h[k] = (uint32_t)(s >> 16); pp.l[k] = (uint32_t)s & 0xffffu;
                        pp.mask[k] = 0; pp.acc[k] = 0; pp.prev[k] = 0u;
                        s = (map_1_a * s + map_1_b) & FPT_MASK48;
                    }
                    pp.maxr = 0u;
                }
                /* Fisher-Yates fixes position i at step i (i = m-1 .. 1), so everything the score needs is read off the swap as it
                   happens: the label that lands at i joins the membership mask of its group and closes the adjacent pair
                   (i, i+1), whose quantised distance goes to the within-B sum while i >= asize and to the within-A sum below
                   asize - 1 (the pair across the group boundary counts for neither, css.c:627-643). The position ranges are
                   split so that each loop has a fixed body. */
                int wa[FPT_P3_PP], wb[FPT_P3_PP];
                if (use_a) {
                    if (m - 1 >= asize && m - 1 >= 1) fpt_p3_run<false, false, WIDE>(pp, m - 1, m - 1, rtab, col0, q, m);
                    fpt_p3_run<false, true, WIDE>(pp, m - 2, max(asize, 1), rtab, col0, q, m);
                } else {
                    if (m - 1 >= asize && m - 1 >= 1) fpt_p3_run<true, false, WIDE>(pp, m - 1, m - 1, rtab, col0, q, m);
                    fpt_p3_run<true, true, WIDE>(pp, m - 2, max(asize, 1), rtab, col0, q, m);
                }
#pragma unroll
                for (int k = 0; k < FPT_P3_PP; k++) { wb[k] = pp.acc[k]; pp.acc[k] = 0; }
                /* position asize-1: first of group A from the top, its pair with position asize crosses the boundary */
                if (asize - 1 >= 1) {
                    if (use_a) {
                        fpt_p3_run<true, false, WIDE>(pp, asize - 1, asize - 1, rtab, col0, q, m);
                        fpt_p3_run<true, true, WIDE>(pp, asize - 2, 1, rtab, col0, q, m);
                    } else {
                        fpt_p3_run<false, false, WIDE>(pp, asize - 1, asize - 1, rtab, col0, q, m);
                        fpt_p3_run<false, true, WIDE>(pp, asize - 2, 1, rtab, col0, q, m);
                    }
                }
                {   /* position 0 keeps what is left there */
                    const uint32_t w0 = *reinterpret_cast<const uint32_t *>(col0);
#pragma unroll
                    for (int k = 0; k < FPT_P3_PP; k++) {
                        const uint32_t c = (w0 >> (8 * k)) & 0xffu;
                        if (use_a) pp.mask[k] |= (mask_t)1 << c;
                        if (asize >= 2) pp.acc[k] += (int)q[c * m + pp.prev[k]];
                        wa[k] = pp.acc[k];
                    }
                }
                /* a draw above 2^31 - 64 may have been a rejected one (probability < 2^-25 per permutation): regenerate that
                   permutation alone on the exact path and walk its labels */
                if (pp.maxr > 2147483648u - 64u) {
                    for (int k = 0; k < FPT_P3_PP; k++) {
                        unsigned char *col = col0 + k;
                        fpt_p3_regen(col, m, fpt_lcg_skip(st_win, (uint64_t)(ndone + first + k) * (uint64_t)draws));
                        mask_t mk = 0;
                        int a = 0, b = 0, prev = 0;
                        for (int i = 0; i < m; i++) {
                            const int c = col[(unsigned)i * FPT_P3_LINE];
                            const bool in_a = i < asize;
                            if (in_a == (use_a != 0)) mk |= (mask_t)1 << c;
                            if (i != 0 && i != asize) { const int qv = (int)q[prev * m + c]; if (in_a) a += qv; else b += qv; }
                            prev = c;
                        }
                        pp.mask[k] = mk; wa[k] = a; wb[k] = b;
                    }
                }
                /* ---- score: the tensor-core product is a warp-wide operation, one batch of 32 permutations (permutation k of
                   every lane) after the other */
#pragma unroll 1
                for (int k = 0; k < FPT_P3_PP; k++) {
                    const bool valid = k < mycount;
                    int hit = 0;
                    bool exact = valid && !use_surrogate;
                    if (use_surrogate) {
                        const long long bet = fpt_p3_bet_mma<KS>(qd, ntiles, qd_rows, pp.mask[0]);
                        if (valid) {
                            const double approx = (double)bet * c_bet - (a_ + b_) * ((double)wa[0] * c_wa + (double)wb[0] * c_wb);
                            const double diff = approx - score;
                            hit = diff > 0.0;
                            exact = !(fabs(diff) > E);
                        }
                    }
                    /* the rare permutation too close to the observed score: its labels regenerated by their lane, re-scored in the
                       reference's order by the whole warp */
                    unsigned ex = __ballot_sync(FPT_FULL_MASK, exact);
                    while (ex) {
                        const int src = 31 - __clz((int)(ex & (0u - ex)));
                        ex &= ex - 1u;
                        unsigned char *col = labels + 4 * ((tid & ~31) + src) + k;
                        if (lane == src) fpt_p3_regen(col, m, fpt_lcg_skip(st_win, (uint64_t)(ndone + first + k) * (uint64_t)draws));
                        __syncwarp();
                        const double sc = fpt_p3_warp_exact_score(X, stage, col, asize, bsize, lane);
                        if (lane == src) { hit = sc >= score ? 1 : 0; rechecks++; }
                        __syncwarp();
                    }
                    hitmask |= (unsigned)hit << k;
                    /* rotate: permutation k + 1 moves to slot 0 (keeps the loop body free of dynamic register indexing) */
#pragma unroll
                    for (int j = 0; j + 1 < FPT_P3_PP; j++) { pp.mask[j] = pp.mask[j + 1]; wa[j] = wa[j + 1]; wb[j] = wb[j + 1]; }
                }
            }
            const int myhits = __popc(hitmask);
            const bool can_stop = ndone + nvalid >= treshold;          /* hits <= permutations drawn: no stop before that */
            const bool last = ndone + nvalid >= runs;
            if (!can_stop && !last) {
                pending += myhits; ndone += nvalid;
                st_round = (map_r_a * st_round + map_r_b) & FPT_MASK48;
                continue;
            }
            /* hits of this round in permutation order (warp scan + per-warp totals), pending hits summed */
            int incl = myhits, pend = pending;
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(FPT_FULL_MASK, incl, o);
                if (lane >= o) incl += y;
            }
            for (int o = 16; o > 0; o >>= 1) pend += __shfl_xor_sync(FPT_FULL_MASK, pend, o);
            int *wh = whits + parity * 2 * FPT_P3_WARPS;
            if (lane == 31) { wh[warp] = incl; wh[FPT_P3_WARPS + warp] = pend; }
            if (tid == 0) *s_flag = -1;
            __syncthreads();
            int before = 0, round_hits = 0, pend_all = 0;
#pragma unroll
            for (int v = 0; v < FPT_P3_WARPS; v++) {
                const int h = wh[v];
                if (v < warp) before += h;
                round_hits += h; pend_all += wh[FPT_P3_WARPS + v];
            }
            parity ^= 1;
            hits += pend_all; pending = 0;
            incl += before;
            if (hits + round_hits >= treshold) {
                /* the treshold-th hit falls into this round: its owner says which permutation it is */
                if (myhits > 0 && hits + incl >= treshold && hits + incl - myhits < treshold) {
                    int need = treshold - (hits + incl - myhits), which = 0;
                    for (int j = 0; j < FPT_P3_PP; j++) if ((hitmask >> j) & 1u) { if (--need == 0) { which = j; break; } }
                    *s_flag = first + which;
                }
                __syncthreads();
                ndone += *s_flag + 1; hits = treshold; stopped = true;
            } else {
                hits += round_hits; ndone += nvalid;
                st_round = (map_r_a * st_round + map_r_b) & FPT_MASK48;
            }
        }
        if (tid == 0) {
            out_p[w] = __ddiv_rn(__dmul_rn((double)(hits + 1), 1.0), (double)(ndone + 1));
            if (out_hits) out_hits[w] = hits;
            if (out_n) out_n[w] = ndone;
        }
        __syncthreads();
    }
    if (recheck_counter && rechecks) atomicAdd(recheck_counter, rechecks);
}

#endif
