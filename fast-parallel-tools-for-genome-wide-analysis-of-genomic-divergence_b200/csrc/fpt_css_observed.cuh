/*
 * fpt_css_observed.cuh — the SIMT building blocks of the large-cohort permutation path (fpt_css_perm_umma.cuh): the
 * surrogate distance, the pipelined Fisher-Yates shuffle, the warp-wide css() in the reference's summation order and the
 * observed-score kernel built on it. No tensor-core instructions here, so this header also compiles under the CPU emulator
 * (tests/emu), which pins it against the oracle.
 *
 * Reference: calc_dist css/css.c:573-587, css css.c:608-647, random_shuffle css.c:700-706 (paths relative to
 * /root/reference/statistics/).
 */
#ifndef FPT_CSS_OBSERVED_CUH
#define FPT_CSS_OBSERVED_CUH

#include "fpt_css.cuh"
#include "fpt_css_perm.cuh"
#include "fpt_css_perm_large.cuh"

/* Distance for the SURROGATE only (the exact paths use the IEEE square root): single-precision reciprocal square root,
   one Newton step on it and one Heron step on the root — eight fp64 operations instead of the ~30 of __dsqrt_rn. Relative
   error against the reference's distance below 2^-50 (the squared length by one fma: 1.5 ulp; the root after the Heron step:
   the square of 1.5 * 2^-42 plus two roundings), which the error bound E accounts for. Symmetric in its two points bit for
   bit (only squares of the differences enter), so every stage that needs q(i, j) computes the same integer. */
FPT_D double fpt_umma_dist(double xi, double yi, double xj, double yj) {
    const double dx = xi - xj, dy = yi - yj;
    const double x = fma(dx, dx, dy * dy);
    if (!(x > 1e-30 && x < 1e30)) return __dsqrt_rn(x);         /* zero, denormal-ish, huge, NaN: the IEEE path */
    double y = (double)rsqrtf((float)x);
    const double e = fma(-(x * y), y, 1.0);
    y = fma(0.5 * y, e, y);
    const double g = x * y;
    return fma(0.5 * y, fma(-g, g, x), g);
}
/* quantised distance between individuals i and j of the embedding X: the value the distance pass stores for (i, j) */
FPT_D unsigned fpt_umma_q(const double *X, int i, int j, double S) {
    return (unsigned)__double2ll_rn(fpt_umma_dist(X[2 * i], X[2 * i + 1], X[2 * j], X[2 * j + 1]) * S);
}

/* exact quotient of a 31-bit draw by n (2 <= n < 2^31): floor(r / n) = umulhi(r, M) >> sh with M = ceil(2^(31+s) / n),
   s = ceil(log2 n), sh = s - 1 = 31 - clz(n - 1). M fits 32 bits (n > 2^(s-1), or n = 2^s and M = 2^31); the error term
   r e / (n 2^(31+s)), e = M n - 2^(31+s) < n <= 2^s, stays below 1/n for r < 2^31, so the floor is exact — no fix-up step. */
FPT_HD uint32_t fpt_magic31(uint32_t n) {
    if (n < 2) return 0u;
    uint32_t s = 0;
    while ((1u << s) < n) s++;
    const unsigned long long p = 1ULL << (31 + s);
    return (uint32_t)((p + n - 1) / n);
}
/* the shuffle's per-n table entry: x = largest accepted draw (fpt_randint_limit), y = fpt_magic31 */
FPT_HD uint2 fpt_umma_rtab_entry(int n) {
    uint2 lm;
    lm.x = n > 0 ? fpt_randint_limit((uint32_t)n) : 0u;
    lm.y = n > 0 ? fpt_magic31((uint32_t)n) : 0u;
    return lm;
}

/* Fisher-Yates of fresh identity labels (css.c:700-706): the reference's random_shuffle on its nrand48 stream, restructured for
   latency. A round of shuffles is ONE dependent chain per permutation (m - 1 swaps), and at a row per thread only a few warps
   run, so what counts is the length of that chain in cycles, not the instruction count:
     - the generator runs as FOUR interleaved sub-streams (states 1..4 steps into the stream, each advanced by the 4-step affine
       map x -> A4 x + C4 mod 2^48), on the (32 high, 16 low) split of the state: three 32-bit multiply-adds per draw, and the
       four draws of a group are independent of each other;
     - r mod n by the exact magic quotient above (no correction step); a draw above the acceptance limit is only remembered,
       and that permutation is replayed on the exact path afterwards (probability < n / 2^31 per draw);
     - the draws of the NEXT group are issued before the swaps of the current one, so the only serial chain left is
       load-load-store-store of the swaps themselves. */
#define FPT_LCG4_A32 0x772C5F11u          /* A^4 mod 2^32 */
#define FPT_LCG4_AH 0x32EB772Cu           /* (A^4 mod 2^48) >> 16 */
#define FPT_LCG4_AL 0x5F11u               /* A^4 mod 2^16 */
#define FPT_LCG4_CH 0x2D3873C4u           /* C4 >> 16, C4 = C (A^3 + A^2 + A + 1) mod 2^48 */
#define FPT_LCG4_CL 0xCD04u               /* C4 mod 2^16 */
/* `pairs`: also return the two adjacent-pair sums of the surrogate (css.c:627-643 on quantised distances: sa over the pairs inside
   the first group, sb inside the second) — position i is final once step i has run, so the pair (i, i+1) is closed right there,
   from the embedding (fpt_umma_q: the same integers the distance pass stores). That arithmetic fills issue slots the swap chain
   leaves empty; a separate sweep over the finished rows cost as much as the shuffles themselves.
   Swaps go four steps at a time: the eight labels are loaded together from the state before the group, the effect of the earlier
   swaps of the group on the later ones is applied in registers (step u reads positions i-u and r_u; of the positions written by an
   earlier step v only r_v can coincide with them), and the eight stores follow in program order — one shared-memory round trip
   per four steps instead of one per step. */
FPT_D void fpt_umma_shuffle(unsigned short *__restrict__ row, int m, const uint2 *__restrict__ rtab, uint64_t st, bool pairs = false,
                            const double *__restrict__ X = nullptr, double S = 0.0, int asize = 0, long long *sa_out = nullptr,
                            long long *sb_out = nullptr) {
    if ((((size_t)row) & 3) == 0) {
        for (int e = 0; e + 1 < m; e += 2) *reinterpret_cast<uint32_t *>(row + e) = (uint32_t)e | ((uint32_t)(e + 1) << 16);
        if (m & 1) row[m - 1] = (unsigned short)(m - 1);
    } else {
        for (int e = 0; e < m; e++) row[e] = (unsigned short)e;
    }
    uint32_t hi[4], lo[4];
    {
        uint64_t s2 = st;
#pragma unroll
        for (int u = 0; u < 4; u++) { fpt_lcg_next(s2); hi[u] = (uint32_t)(s2 >> 16); lo[u] = (uint32_t)s2 & 0xffffu; }
    }
    uint32_t over = 0u;
    /* the swap positions of steps i, i-1, i-2, i-3 (n = i+1 .. i-2); a step below 1 draws with n = 2 and is never used */
    auto draw4 = [&](int i, uint32_t *rem) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint32_t n = (uint32_t)max(i - u + 1, 2);
            const uint2 lm = rtab[n];
            const uint32_t r = hi[u] >> 1;                      /* nrand48: bits 17..47 */
            const uint32_t t0 = lo[u] * FPT_LCG4_AL + FPT_LCG4_CL;
            hi[u] = hi[u] * FPT_LCG4_A32 + (lo[u] * FPT_LCG4_AH + ((t0 >> 16) + FPT_LCG4_CH));
            lo[u] = t0 & 0xffffu;
            over |= lm.x - r;
            rem[u] = r - (__umulhi(r, lm.y) >> (31 - __clz((int)(n - 1)))) * n;
        }
    };
    long long sa = 0, sb = 0;
    double px = 0.0, py = 0.0;                                  /* coordinates of the label one position up */
    auto close_pair = [&](int pos, unsigned c) {                /* label c is final at position pos: the pair (pos, pos + 1) */
        const double cx = X[2 * c], cy = X[2 * c + 1];
        const int col = pos + 1;
        if (col < m && col != asize) {
            const long long qv = (long long)(unsigned)__double2ll_rn(fpt_umma_dist(cx, cy, px, py) * S);
            if (col > asize) sb += qv; else sa += qv;
        }
        px = cx; py = cy;
    };
    uint32_t cur[4], nxt[4];
    draw4(m - 1, cur);
    int i = m - 1;
    for (; i >= 4; i -= 4) {
        draw4(i - 4, nxt);                                      /* the next group's draws go out ahead of this group's swaps */
        unsigned t[4], x[4];
#pragma unroll
        for (int u = 0; u < 4; u++) { t[u] = row[i - u]; x[u] = row[cur[u]]; }
#pragma unroll
        for (int u = 1; u < 4; u++)
#pragma unroll
            for (int v = 0; v < u; v++) {
                if (cur[v] == (uint32_t)(i - u)) t[u] = t[v];
                if (cur[u] == cur[v]) x[u] = t[v];
            }
#pragma unroll
        for (int u = 0; u < 4; u++) { row[i - u] = (unsigned short)x[u]; row[cur[u]] = (unsigned short)t[u]; }
        if (pairs) {
            /* the four pairs of the group, stage by stage (four independent dependency chains side by side): fpt_umma_dist's
               fast path for all four, its IEEE path for the rare operand outside the fast range */
            double cx[4], cy[4], d2[4], dd[4];
#pragma unroll
            for (int u = 0; u < 4; u++) { const double2 c2 = *reinterpret_cast<const double2 *>(X + 2 * x[u]); cx[u] = c2.x; cy[u] = c2.y; }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const double dx = cx[u] - (u ? cx[u - 1] : px), dy = cy[u] - (u ? cy[u - 1] : py);
                d2[u] = fma(dx, dx, dy * dy);
            }
            bool slow = false;
#pragma unroll
            for (int u = 0; u < 4; u++) slow |= !(d2[u] > 1e-30 && d2[u] < 1e30);
#pragma unroll
            for (int u = 0; u < 4; u++) dd[u] = (double)rsqrtf((float)d2[u]);
#pragma unroll
            for (int u = 0; u < 4; u++) { const double e = fma(-(d2[u] * dd[u]), dd[u], 1.0); dd[u] = fma(0.5 * dd[u], e, dd[u]); }
#pragma unroll
            for (int u = 0; u < 4; u++) { const double g = d2[u] * dd[u]; dd[u] = fma(0.5 * dd[u], fma(-g, g, d2[u]), g); }
            if (slow) {
#pragma unroll
                for (int u = 0; u < 4; u++) if (!(d2[u] > 1e-30 && d2[u] < 1e30)) dd[u] = __dsqrt_rn(d2[u]);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int col = i - u + 1;
                const long long qv = (long long)(unsigned)__double2ll_rn(dd[u] * S);
                if (col < m && col != asize) { if (col > asize) sb += qv; else sa += qv; }
            }
            px = cx[3]; py = cy[3];
        }
#pragma unroll
        for (int u = 0; u < 4; u++) cur[u] = nxt[u];
    }
#pragma unroll
    for (int u = 0; u < 3; u++) {                               /* the last one to three steps, one at a time */
        if (i - u > 0) {
            const unsigned short tt = row[i - u], xx = row[cur[u]];
            row[i - u] = xx; row[cur[u]] = tt;
            if (pairs) close_pair(i - u, xx);
        }
    }
    if (pairs) close_pair(0, row[0]);
    if (over >> 31) {                                           /* a rejected draw: exact replay of this permutation */
        for (int e = 0; e < m; e++) row[e] = (unsigned short)e;
        uint64_t s2 = st;
        int used = 0;
        for (int j = m - 1; j > 0; j--) {
            const int rr = (int)fpt_randint((uint32_t)(j + 1), s2, used);
            const unsigned short tt = row[j]; row[j] = row[rr]; row[rr] = tt;
        }
        if (pairs) {
            sa = 0; sb = 0;
            for (int pos = m - 1; pos >= 0; pos--) close_pair(pos, row[pos]);
        }
    }
    if (pairs) { *sa_out = sa; *sb_out = sb; }
}

/* Observed score of every window (identity labels) in the reference's summation order, css.c:608-647: a quarter of a million
   dependent fp64 additions per window at m = 1000, i.e. latency, so ONE WARP per window and many windows per SM. The lanes
   compute the next 32 distances (calc_dist, css.c:573-587) while lane 0 adds the previous 32 in order. Bit-identical to
   fpt_css_score_identity on the stored distance matrix. */
#define FPT_OBS_WARPS 4
FPT_HD size_t fpt_css_observed_smem_bytes(int m) { return (size_t)FPT_OBS_WARPS * ((size_t)2 * m * 8 + 32 * 8); }

/* kind 0: between-group pairs (i from asize-1 down, j from m-1 down to asize); 1: adjacent pairs of the first group;
   2: of the second group — each in the order the reference adds them */
FPT_D double fpt_observed_chain(const double *X, double *stage, int kind, int asize, int bsize, int lane,
                                const unsigned short *lab = nullptr) {
    const int total = kind == 0 ? asize * bsize : (kind == 1 ? asize - 1 : bsize - 1);     /* m <= 1024: fits an int */
    double acc = 0.0;
    auto dist_of = [&](int e) -> double {
        int i, j;
        if (kind == 0) { const int r = e / bsize; i = asize - 1 - r; j = asize + bsize - 1 - (e - r * bsize); }
        else if (kind == 1) { i = asize - 2 - e; j = i + 1; }
        else { i = asize + bsize - 2 - e; j = i + 1; }
        if (lab) { i = lab[i]; j = lab[j]; }                    /* a permutation's labels instead of the identity */
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
    return acc;                                                 /* lane 0 holds the sum */
}

/* css() of css.c:608-647 for one label row by a whole warp, reference summation order; every lane returns the score */
FPT_D double fpt_warp_css_score(const double *X, double *stage, const unsigned short *lab, int asize, int bsize, int lane) {
    double bet = fpt_observed_chain(X, stage, 0, asize, bsize, lane, lab);
    const double wa0 = asize > 1 ? fpt_observed_chain(X, stage, 1, asize, bsize, lane, lab) : 0.0;
    const double wb0 = bsize > 1 ? fpt_observed_chain(X, stage, 2, asize, bsize, lane, lab) : 0.0;
    bet = __ddiv_rn(bet, (double)((long long)asize * bsize));
    const double wa = asize > 1 ? __ddiv_rn(wa0, (double)((long long)asize * asize * (asize - 1))) : 0.0;
    const double wb = bsize > 1 ? __ddiv_rn(wb0, (double)((long long)bsize * bsize * (bsize - 1))) : 0.0;
    return __shfl_sync(FPT_FULL_MASK, __dsub_rn(bet, __dmul_rn((double)(asize + bsize), __dadd_rn(wa, wb))), 0);
}

__global__ void __launch_bounds__(FPT_OBS_WARPS * 32)
fpt_css_observed_kernel(const double *__restrict__ Xall, int m, int asize, int bsize, long long nwin,
                        const unsigned char *__restrict__ status, double *__restrict__ out_score) {
    FPT_DYN_SMEM(smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *X = reinterpret_cast<double *>(smem) + (size_t)warp * (2 * m + 32);
    double *stage = X + 2 * m;
    for (long long w = (long long)blockIdx.x * FPT_OBS_WARPS + warp; w < nwin; w += (long long)gridDim.x * FPT_OBS_WARPS) {
        if (status[w] != FPT_WIN_SCORED) continue;
        __syncwarp();
        for (int e = lane; e < 2 * m; e += 32) X[e] = Xall[(size_t)w * 2 * m + e];
        __syncwarp();
        const double sc = fpt_warp_css_score(X, stage, nullptr, asize, bsize, lane);
        if (lane == 0) out_score[w] = sc;
    }
}

#endif
