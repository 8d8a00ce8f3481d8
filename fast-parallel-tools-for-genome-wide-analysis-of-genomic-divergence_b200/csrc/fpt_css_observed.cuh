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

/* Fisher-Yates of fresh identity labels (css.c:700-706) like fpt_generate_labels, software-pipelined: the draws do not depend on
   the labels, so the four draws of the NEXT group (table loads, generator steps, modulo) are issued before the four swaps of the
   current one and the two dependency chains run side by side */
FPT_D void fpt_umma_shuffle(unsigned short *row, int m, const uint2 *rtab, uint64_t st) {
    for (int e = 0; e < m; e++) row[e] = (unsigned short)e;
    uint64_t s2 = st;
    uint32_t over = 0u;
    /* the swap positions of steps i, i-1, i-2, i-3 (n = i+1 .. i-2); positions of a step below 1 are never used */
    auto draw4 = [&](int i, uint32_t *rem) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int iu = i - u;
            const uint32_t n = (uint32_t)(iu > 0 ? iu + 1 : 2);
            const uint2 lm = rtab[n];
            if (iu > 0) {
                const uint32_t r = (uint32_t)(fpt_lcg_next(s2) >> 17);
                over |= lm.x - r;
                uint32_t rm = r - __umulhi(r, lm.y) * n;
                if (rm >= n) rm -= n;
                rem[u] = rm;
            } else {
                rem[u] = 0u;
            }
        }
    };
    uint32_t cur[4], nxt[4];
    draw4(m - 1, cur);
    for (int i = m - 1; i > 0; i -= 4) {
        draw4(i - 4, nxt);                                      /* the next group's draws go out ahead of this group's swaps */
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (i - u > 0) {
                const unsigned short t = row[i - u], x = row[cur[u]];
                row[i - u] = x; row[cur[u]] = t;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) cur[u] = nxt[u];
    }
    if (over >> 31) fpt_generate_labels<unsigned short>(row, m, rtab, st);     /* a rejected draw: exact replay */
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
