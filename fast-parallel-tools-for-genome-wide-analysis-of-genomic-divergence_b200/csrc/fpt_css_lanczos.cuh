/*
 * fpt_css_lanczos.cuh — classical MDS of one window for LARGE cohorts (m beyond the one-warp path of fpt_css_eig.cuh),
 * one CTA per window (reference: cmds, css/css.c:505-560 — GSL's dense symmetric eigensolver, two largest eigenpairs kept).
 *
 * At m = 1000 (BASELINE configs[4]) a full reduction of the 8 MB matrix re-reads it ~m times; the reference only wants the
 * two largest eigenpairs, and those are the ones a Krylov method finds first. So:
 *
 *   1. B = -1/2 (S - r 1' - 1 r' + g), S = D.D, formed once in the CTA's global scratch (read-only afterwards)
 *   2. Lanczos with FULL re-orthogonalisation (classical Gram-Schmidt applied twice) against every previous vector:
 *      w = B q_j (one coalesced pass over B, a warp per row), h = Q'w, w -= Q h, twice; alpha_j = h_j, beta_j = |w|
 *   3. every FPT_LANCZOS_CHECK steps warp 0 solves the j x j tridiagonal (fpt_warp_tri_eig: Sturm multisection + inverse
 *      iteration) and the residual bounds beta_j |y_c[j-1]| of the two leading Ritz pairs decide: stop once both are
 *      below 1e-9 of their spectral gap (eigenvector error ~ residual / gap), or at 1e-14 of the norm, or at the cap
 *   4. X = [Q y_1, Q y_2] diag(sqrt l1, sqrt l2)
 *
 * With full re-orthogonalisation the recurrence is numerically a Householder reduction stopped early: no ghost
 * eigenvalues, and at j = m it is complete. Traffic per window: (steps) x 8 m^2 bytes for the products (HBM/L2, coalesced)
 * plus ~16 m j bytes per step for the basis — against ~4 m^3 bytes for a reduction in place.
 */
#ifndef FPT_CSS_LANCZOS_CUH
#define FPT_CSS_LANCZOS_CUH

#include "fpt_css.cuh"
#include "fpt_css_eig.cuh"

/* diagnostic: SM cycles thread 0 spent per phase of the large-cohort MDS kernel, summed over CTAs and windows (0 dissimilarity,
   1 row means + code conversion, 2 products, 3 Gram-Schmidt, 4 tridiagonal solves, 5 norms / next vector, 6 coordinates; slot 7
   counts Lanczos steps over all windows); read and reset by fpt_debug_lanczos_phases() */
#ifndef FPT_EMU
__device__ unsigned long long fpt_lanczos_phase_cycles[8];
#define FPT_LZ_MARK(slot) do { if (threadIdx.x == 0) { const long long now_ = clock64(); atomicAdd(&fpt_lanczos_phase_cycles[slot], (unsigned long long)(now_ - lz_mark)); lz_mark = now_; } } while (0)
#define FPT_LZ_START() long long lz_mark = clock64()
#else
#define FPT_LZ_MARK(slot) do { } while (0)
#define FPT_LZ_START() do { } while (0)
#endif

#define FPT_LANCZOS_CHECK 8
#define FPT_LANCZOS_FIRST_CHECK 16

FPT_HD int fpt_lanczos_cap(int m) { return m < 384 ? m : 384; }

struct FptLanczosSmem {
    double *q;        /* m: current Lanczos vector (read by every row of the product) */
    double *w;        /* m: B q and its orthogonalised remainder */
    double *alpha;    /* cap */
    double *beta;     /* cap */
    double *h;        /* cap: projections on the basis */
    double *rmean;    /* m */
    double *lut;      /* 256: S(code) of the 8-bit code form */
    int *zoff;        /* m + 1: row offsets into the list of blank (code 0) entries, arithmetic form of the product */
    FptEigWork ew;    /* tridiagonal eigen-solve of order <= cap, warp 0 */
    FptCssScratch sc; /* reductions + bit-plane words of the dissimilarity stage */
};

FPT_HD size_t fpt_lanczos_smem_bytes(int m, int wch) {
    const int cap = fpt_lanczos_cap(m);
    size_t off = (size_t)(3 * m + 16) * 8 + 256 * 8 + (size_t)3 * cap * 8 + (size_t)13 * cap * 8 + 66 * 8;
    off = (off + 15) & ~(size_t)15;
    off += ((size_t)(m + 1) * 4 + 15) & ~(size_t)15;
    return off + (size_t)wch * 2 * m * 4;
}

/* capacity of the per-CTA list of blank entries (column indices, row by row): windows with more blanks than that stream their
   squares through the table form of the product instead */
FPT_HD size_t fpt_lanczos_zcap(int m) { return (size_t)16 * m; }
/* global scratch per CTA of the code route: the Lanczos basis, then the blank list */
FPT_HD size_t fpt_lanczos_cta_scratch_bytes(int m) {
    return (((size_t)fpt_lanczos_cap(m) * m * 8 + fpt_lanczos_zcap(m) * 4) + 255) & ~(size_t)255;
}

FPT_D FptLanczosSmem fpt_lanczos_carve(unsigned char *smem, int m, int wch) {
    FptLanczosSmem s;
    const int cap = fpt_lanczos_cap(m);
    double *p = (double *)smem;
    s.q = p; p += m + 16; s.w = p; p += m; s.rmean = p; p += m;      /* q: 16 spare entries, kept zero (padded code columns) */
    s.lut = p; p += 256;
    s.alpha = p; p += cap; s.beta = p; p += cap; s.h = p; p += cap;
    s.ew.A = 0;
    s.ew.d = p; p += cap; s.ew.e = p; p += cap; s.ew.tau = p; p += cap;
    s.ew.pv = p; p += cap; s.ew.wv = p; p += cap;
    s.ew.y = p; p += 2 * cap;
    s.ew.lu = p; p += 6 * cap;
    s.ew.wbuf = 0; s.ew.wch = 0;
    s.sc.red = p; p += 33;
    s.sc.redi = (long long *)p; p += 33;
    size_t off = (size_t)((unsigned char *)p - smem);
    off = (off + 15) & ~(size_t)15;
    s.zoff = (int *)(smem + off);
    off += ((size_t)(m + 1) * 4 + 15) & ~(size_t)15;
    s.sc.wbuf = (unsigned *)(smem + off);
    s.sc.wch = wch;
    s.sc.pairs = 0;
    return s;
}

/* y = B x for a symmetric m x m matrix in global memory: a warp per row, lanes across the row (coalesced), four rows in
   flight per warp and two column strides per trip so that enough loads are outstanding to cover the HBM latency;
   x and y in shared memory */
FPT_D void fpt_cta_symv(const double *__restrict__ B, int m, const double *x, double *y) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int i = 4 * warp; i < m; i += 4 * nwarp) {
        const int nr = m - i < 4 ? m - i : 4;
        const double *r0 = B + (size_t)i * m;
        const double *r1 = nr > 1 ? r0 + m : r0, *r2 = nr > 2 ? r0 + 2 * (size_t)m : r0, *r3 = nr > 3 ? r0 + 3 * (size_t)m : r0;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
        int j = lane;
        for (; j + 32 < m; j += 64) {
            const double xa = x[j], xb = x[j + 32];
            const double a0 = r0[j], a1 = r1[j], a2 = r2[j], a3 = r3[j];
            const double b0 = r0[j + 32], b1 = r1[j + 32], b2 = r2[j + 32], b3 = r3[j + 32];
            s0 += a0 * xa; s1 += a1 * xa; s2 += a2 * xa; s3 += a3 * xa;
            s0 += b0 * xb; s1 += b1 * xb; s2 += b2 * xb; s3 += b3 * xb;
        }
        if (j < m) { const double xa = x[j]; s0 += r0[j] * xa; s1 += r1[j] * xa; s2 += r2[j] * xa; s3 += r3[j] * xa; }
        s0 = fpt_warp_sum(s0); s1 = fpt_warp_sum(s1); s2 = fpt_warp_sum(s2); s3 = fpt_warp_sum(s3);
        if (lane == 0) { y[i] = s0; if (nr > 1) y[i + 1] = s1; if (nr > 2) y[i + 2] = s2; if (nr > 3) y[i + 3] = s3; }
    }
}

/* The same product from the COMPACT form of B (a quarter of the bytes): B = -1/2 (S - r 1' - 1 r' + g) and S = D.D holds only
   squares of small integers (counts of opposite homozygotes) plus one repeated real (the fill_averages value, css.c:337-366),
   so the matrix streamed per Lanczos step is 16-bit codes — c >= 1: S = c^2, c = 0: S = v2 — and
       y = -1/2 ( S x - r (1'x) - 1 (r'x) + g (1'x) ).
   The integer c^2 becomes a double by the 2^52 trick (one fp64 add instead of a conversion instruction). A warp per row,
   ROWS rows in flight (eight with 8-bit codes: 2 KB of loads outstanding per warp), each lane one 8-byte load of consecutive codes: four 16-bit ones (m % 4 == 0) or, when no count exceeds
   255, eight 8-bit ones (m % 8 == 0). sx = 1'x and rx = r'x are given. */
template <typename CodeT, int ROWS, bool LUT, bool ZLIST>
FPT_D void fpt_cta_symv_codes(const CodeT *__restrict__ C, int m, int ld, const double *x, double *y, const double *rmean,
                              double g, double v2, double sx, double rx, const double *lut, const int *zoff, const int *zlist) {
    constexpr int PER = 8 / (int)sizeof(CodeT);                 /* codes per 8-byte load: 4 or 8 */
    constexpr int BITS = 8 * (int)sizeof(CodeT);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    /* The (row group, column chunk) pairs of this warp as ONE sequence, software-pipelined: the ROWS loads of the next pair are
       issued before the multiply-adds of the current one, so a warp's own arithmetic covers its own load latency (ncu of the
       nested-loop form: 42 % of the warp cycles waiting on the loads just issued, issue slots 39 % busy). */
    const int nchunk = (m + 32 * PER - 1) / (32 * PER);
    const int first = ROWS * warp, gstep = ROWS * nwarp;
    const int ngrp = first < m ? (m - first + gstep - 1) / gstep : 0;
    const int total = ngrp * nchunk;
    const int j0 = PER * lane;
    uint2 nxt[ROWS];
    int gi = first, ch = 0;
    auto load = [&](int i, int chunk, uint2 (&cc)[ROWS]) {
        const int j = j0 + chunk * 32 * PER;
        const int nr = m - i < ROWS ? m - i : ROWS;
        const CodeT *r0 = C + (size_t)i * ld + j;              /* columns m .. ld-1 (if any) meet x = 0 */
        if (j < m) {
#pragma unroll
            for (int r = 0; r < ROWS; r++) cc[r] = *reinterpret_cast<const uint2 *>(r0 + (size_t)(r < nr ? r : 0) * ld);
        } else {
#pragma unroll
            for (int r = 0; r < ROWS; r++) cc[r] = make_uint2(0u, 0u);
        }
    };
    if (total > 0) load(gi, ch, nxt);
    double acc[ROWS];
#pragma unroll
    for (int r = 0; r < ROWS; r++) acc[r] = 0.0;
#pragma unroll 1
    for (int t = 0; t < total; t++) {
        uint2 cc[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; r++) cc[r] = nxt[r];
        const int i = gi, j = j0 + ch * 32 * PER;
        const bool last_chunk = ch == nchunk - 1;
        if (last_chunk) { ch = 0; gi += gstep; } else ch++;
        if (t + 1 < total) load(gi, ch, nxt);
        if (j < m) {
#pragma unroll
            for (int k = 0; k < PER; k += 2) {
                const double2 xv = *reinterpret_cast<const double2 *>(x + j + k);
#pragma unroll
                for (int kk = 0; kk < 2; kk++) {
                    const double xk = kk ? xv.y : xv.x;
#pragma unroll
                    for (int r = 0; r < ROWS; r++) {
                        const unsigned word = ((k + kk) * BITS) < 32 ? cc[r].x : cc[r].y;
                        const unsigned c = (word >> (((k + kk) * BITS) & 31)) & ((1u << BITS) - 1u);
                        if (ZLIST) {
                            /* arithmetic form: c^2 lands in the mantissa of 2^52 by ONE wide multiply-add (IMAD.WIDE with the
                               exponent pattern as the addend), one fp64 add removes the 2^52; blanks contribute nothing here —
                               their fill value is added per row from the blank list below. No shared-memory traffic per element
                               (ncu of the table form: the table look-ups keep the shared-memory pipe 56 % busy). */
                            const unsigned long long t52 = (unsigned long long)c * c + 0x4330000000000000ULL;
                            acc[r] = fma(__longlong_as_double((long long)t52) - 4503599627370496.0, xk, acc[r]);
                        } else if (LUT) {
                            /* 8-bit codes: S(c) from a 256-entry shared-memory table (c^2, or v2 for c = 0) — one load instead
                               of a multiply, the 2^52 conversion add, a compare and two selects per element */
                            acc[r] = fma(lut[c], xk, acc[r]);
                        } else {
                            const double sq = __hiloint2double(0x43300000, (int)(c * c)) - 4503599627370496.0;
                            acc[r] = fma(c ? sq : v2, xk, acc[r]);
                        }
                    }
                }
            }
        }
        if (last_chunk) {
            const int nr = m - i < ROWS ? m - i : ROWS;
#pragma unroll
            for (int r = 0; r < ROWS; r++) acc[r] = fpt_warp_sum(acc[r]);
            if (ZLIST) {
                /* lane r finishes row i + r: the sum of x over the row's blank columns (the diagonal at least) times the fill value */
                const int zb = lane < nr ? zoff[i + lane] : 0, zl = lane < nr ? zoff[i + lane + 1] - zb : 0;
                double zc = 0.0;
                if (__all_sync(FPT_FULL_MASK, zl <= 4)) {
                    for (int t = 0; t < zl; t++) zc += x[zlist[zb + t]];
                } else {
                    for (int r = 0; r < nr; r++) {
                        const int b = __shfl_sync(FPT_FULL_MASK, zb, r), l = __shfl_sync(FPT_FULL_MASK, zl, r);
                        double part = 0.0;
                        for (int t = lane; t < l; t += 32) part += x[zlist[b + t]];
                        part = fpt_warp_sum(part);
                        if (lane == r) zc = part;
                    }
                }
                double mine = 0.0;
#pragma unroll
                for (int r = 0; r < ROWS; r++) if (lane == r) mine = acc[r];
                if (lane < nr) y[i + lane] = -0.5 * ((((mine + v2 * zc) - rmean[i + lane] * sx) - rx) + g * sx);
            } else if (lane == 0) {
#pragma unroll
                for (int r = 0; r < ROWS; r++)
                    if (r < nr) y[i + r] = -0.5 * (((acc[r] - rmean[i + r] * sx) - rx) + g * sx);
            }
#pragma unroll
            for (int r = 0; r < ROWS; r++) acc[r] = 0.0;
        }
    }
}

/* one Gram-Schmidt pass of w against the nq basis vectors Q[0..nq) (rows of length m in global memory), BLOCK by block of one
   vector per warp: the projections of w on the block (a warp per vector), then w minus the block's share — so the second read of
   a basis vector follows its first within microseconds and comes from L1 / L2, and the basis crosses the HBM bus once per step
   instead of twice (the basis of 296 resident windows is ~190 MB: it does not live in L2). Between blocks this is modified
   Gram-Schmidt, inside a block classical; the projections are ADDED to h so that two passes accumulate the exact coefficients. */
FPT_D void fpt_cta_cgs_pass(const double *__restrict__ Q, int nq, int m, double *w, double *h, double *hpass) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    for (int b0 = 0; b0 < nq; b0 += nwarp) {
        const int nb = min(nwarp, nq - b0);
        if (warp < nb) {
            const int i = b0 + warp;
            const double *qi = Q + (size_t)i * m;
            /* a basis vector is read once and comes from L2 / HBM: the pass is a chain of load latencies unless a lane has its
               share of the vector in flight in a few large batches — 16 loads per lane, 512 elements per trip (the four-loads form
               spent 8 latencies per 1000-element vector; Gram-Schmidt was 29 % of the kernel) */
            double s = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            for (int base = 0; base < m; base += 512) {
                double v[16];
#pragma unroll
                for (int u = 0; u < 16; u++) { const int e = base + lane + 32 * u; v[u] = e < m ? qi[e] : 0.0; }
#pragma unroll
                for (int u = 0; u < 16; u += 4) {
                    const int e = base + lane + 32 * u;
                    if (e < m) s = fma(v[u], w[e], s);
                    if (e + 32 < m) s1 = fma(v[u + 1], w[e + 32], s1);
                    if (e + 64 < m) s2 = fma(v[u + 2], w[e + 64], s2);
                    if (e + 96 < m) s3 = fma(v[u + 3], w[e + 96], s3);
                }
            }
            s = fpt_warp_sum((s + s1) + (s2 + s3));
            if (lane == 0) { hpass[i] = s; h[i] += s; }
        }
        __syncthreads();
        for (int e = threadIdx.x; e < m; e += blockDim.x) {
            double acc = w[e];
            const double *qe = Q + (size_t)b0 * m + e;
            int u0 = 0;
            for (; u0 + 8 <= nb; u0 += 8) {                            /* eight independent loads, then the eight updates in order */
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) v[u] = qe[(size_t)(u0 + u) * m];
#pragma unroll
                for (int u = 0; u < 8; u++) acc = fma(-hpass[b0 + u0 + u], v[u], acc);
            }
            for (; u0 < nb; u0++) acc = fma(-hpass[b0 + u0], qe[(size_t)u0 * m], acc);
            w[e] = acc;
        }
        __syncthreads();
    }
}

/* The matrix B = -1/2 J (D.D) J as the Lanczos product streams it: 8-bit count codes (form 2), 16-bit count codes (form 1) —
   code c >= 1: S = c^2, code 0: S = v2, and y = -1/2 (S x - r (1'x) - 1 (r'x) + g (1'x)) — or B itself in fp64 (form 0). `ld` is the
   row stride in elements (codes: columns m .. ld-1 hold code 0 and meet the zero tail of the Lanczos vector). */
struct FptLzMatrix {
    int form, ld;
    const void *data;
    double g, v2;
    const int *zlist;     /* form 2 only: non-null = arithmetic squares + the blank list (row offsets in FptLanczosSmem::zoff) */
};

/* Lanczos with full re-orthogonalisation on B (given as FptLzMatrix; s.rmean holds the row means of S for the code forms);
   Q (>= cap x m doubles, global) receives the basis. X: 2m doubles written by the CTA; evals3 optional. All threads take part.
   PRODUCT 3: only the arithmetic 8-bit product is compiled in (M.form == 2 with a blank list) — the kernel that carries it holds
   one hot loop and keeps it in registers; PRODUCT 0: table, 16-bit and fp64 products, chosen per window from M.form. */
template <int PRODUCT>
FPT_D void fpt_lanczos_iterate(const FptLzMatrix &M, double *Q, int m, double *X, double *evals3, const FptLanczosSmem &s, int *steps_out) {
    const int T = blockDim.x, tid = threadIdx.x;
    const int cap = fpt_lanczos_cap(m);
    __shared__ double sh_val[4];
    __shared__ int sh_flag;
    FPT_LZ_START();
    const int compact = M.form > 0, narrow = M.form == 2;
    const unsigned short *codes = reinterpret_cast<const unsigned short *>(M.data);
    const unsigned char *codes8 = reinterpret_cast<const unsigned char *>(M.data);
    const double *A = reinterpret_cast<const double *>(M.data);
    const double g = M.g, v2 = M.v2;
    for (int e = tid; e < 16; e += T) s.q[m + e] = 0.0;
    if (PRODUCT != 3 && narrow) for (int e = tid; e < 256; e += T) s.lut[e] = e ? (double)(e * e) : v2;
    /* ---- 2. Lanczos. Start vector: fixed pseudo-random signs and magnitudes (any vector with a component along the
       leading eigenvectors works; a fixed one keeps runs reproducible) */
    double nn = 0.0;
    for (int e = tid; e < m; e += T) {
        const unsigned hsh = ((unsigned)e * 2654435761u + 40503u) >> 7;
        const double v = ((double)(hsh & 0xffffu) / 65536.0) - 0.5;
        s.w[e] = v; nn += v * v;
    }
    nn = fpt_block_sum(nn, s.sc.red);
    const double rn0 = 1.0 / sqrt(nn);
    for (int e = tid; e < m; e += T) { const double v = s.w[e] * rn0; s.q[e] = v; Q[e] = v; }
    __syncthreads();

    int nvec = 0;                       /* order of T when the iteration stops */
    int have = 0;                       /* 1: a converged (or final) tridiagonal solve sits in s.ew.y */
    double lam1 = 0.0, lam2 = 0.0, lam3 = 0.0, bnorm = 0.0;
    int degenerate = 0;                 /* 1: B q0 = 0 or non-finite */
    for (int j = 0; j < cap; j++) {
        if (compact) {
            double sx = 0.0, rx = 0.0;
            for (int e = tid; e < m; e += T) { const double v = s.q[e]; sx += v; rx += s.rmean[e] * v; }
            sx = fpt_block_sum(sx, s.sc.red);
            rx = fpt_block_sum(rx, s.sc.red);
            if (PRODUCT == 3) fpt_cta_symv_codes<unsigned char, 8, false, true>(codes8, m, M.ld, s.q, s.w, s.rmean, g, v2, sx, rx, s.lut, s.zoff, M.zlist);
            else if (narrow) fpt_cta_symv_codes<unsigned char, 8, true, false>(codes8, m, M.ld, s.q, s.w, s.rmean, g, v2, sx, rx, s.lut, nullptr, nullptr);
            else fpt_cta_symv_codes<unsigned short, 4, false, false>(codes, m, M.ld, s.q, s.w, s.rmean, g, v2, sx, rx, s.lut, nullptr, nullptr);
        } else if (PRODUCT != 3) {
            fpt_cta_symv(A, m, s.q, s.w);
        }
        FPT_LZ_MARK(2);
        /* Orthogonalisation = the three-term recurrence (the two large components, along q_j and q_{j-1}, removed locally)
           followed by ONE classical Gram-Schmidt pass against the whole basis, which then only has to remove what rounding
           re-introduced; a second pass runs when the first one took away more than half of the vector (the "twice is enough"
           criterion) — with the recurrence in front that is the exception, so the basis is read once per step, not twice. */
        for (int i = tid; i <= j; i += T) s.h[i] = 0.0;
        __syncthreads();                                            /* the product's rows were written by other warps */
        double aj = 0.0;
        for (int e = tid; e < m; e += T) aj += s.q[e] * s.w[e];
        aj = fpt_block_sum(aj, s.sc.red);
        {
            const double bprev = j > 0 ? s.beta[j - 1] : 0.0;
            const double *qp = Q + (size_t)(j > 0 ? j - 1 : 0) * m;
            for (int e = tid; e < m; e += T) s.w[e] -= aj * s.q[e] + (j > 0 ? bprev * qp[e] : 0.0);
            if (tid == 0) s.h[j] = aj;
        }
        __syncthreads();
        double n0 = 0.0;
        for (int e = tid; e < m; e += T) n0 += s.w[e] * s.w[e];
        n0 = fpt_block_sum(n0, s.sc.red);
        fpt_cta_cgs_pass(Q, j + 1, m, s.w, s.h, s.ew.lu);            /* lu is free between tridiagonal solves */
        double b2 = 0.0;
        for (int e = tid; e < m; e += T) b2 += s.w[e] * s.w[e];
        b2 = fpt_block_sum(b2, s.sc.red);
        if (!(b2 >= 0.5 * n0)) {
            fpt_cta_cgs_pass(Q, j + 1, m, s.w, s.h, s.ew.lu);
            b2 = 0.0;
            for (int e = tid; e < m; e += T) b2 += s.w[e] * s.w[e];
            b2 = fpt_block_sum(b2, s.sc.red);
        }
        FPT_LZ_MARK(3);
        const double bj = sqrt(b2);
        if (tid == 0) { s.alpha[j] = s.h[j]; s.beta[j] = bj; }
        __syncthreads();
        nvec = j + 1;
        bnorm = fmax(bnorm, fabs(s.alpha[j]) + bj + (j > 0 ? s.beta[j - 1] : 0.0));
        if (!(bnorm < 1e300)) { degenerate = 2; break; }          /* NaN / Inf in the input */
        if (j == 0 && bnorm == 0.0) { degenerate = 1; break; }    /* B q0 = 0: B = 0 (all dissimilarities equal) */
        const bool breakdown = !(bj > 1e-14 * bnorm);             /* invariant subspace: T_j is exact */
        FPT_LZ_MARK(5);
        const bool check = breakdown || nvec == cap || (nvec >= FPT_LANCZOS_FIRST_CHECK && (nvec % FPT_LANCZOS_CHECK) == 0) || nvec == m;
        if (check && nvec >= 2) {
            if (tid < 32) {
                for (int i = tid; i < nvec; i += 32) { s.ew.d[i] = s.alpha[i]; s.ew.e[i] = i < nvec - 1 ? s.beta[i] : 0.0; }
                __syncwarp();
                double l1 = 0.0, l2 = 0.0, l3 = 0.0, tn = 0.0;
                const int ok = fpt_warp_tri_eig(nvec, s.ew, nvec >= 3, l1, l2, l3, tn);
                if (tid == 0) {
                    int done = 0;
                    if (ok) {
                        const double r1 = bj * fabs(s.ew.y[nvec - 1]), r2 = bj * fabs(s.ew.y[2 * nvec - 1]);
                        const double gap1 = l1 - l2, gap2 = fmin(l1 - l2, nvec >= 3 ? l2 - l3 : l1 - l2);
                        const double floor_ = 1e-14 * tn;
                        done = (r1 <= fmax(1e-9 * gap1, floor_)) && (r2 <= fmax(1e-9 * gap2, floor_));
                    }
                    sh_val[0] = l1; sh_val[1] = l2; sh_val[2] = l3; sh_val[3] = tn;
                    sh_flag = ok ? (done ? 2 : 1) : 0;
                }
            }
            __syncthreads();
            lam1 = sh_val[0]; lam2 = sh_val[1]; lam3 = sh_val[2];
            const int flag = sh_flag;
            __syncthreads();
            FPT_LZ_MARK(4);
            have = flag != 0;
            if (flag == 0) { degenerate = sh_val[3] == 0.0 ? 1 : 2; break; }
            if (flag == 2 || breakdown || nvec == cap || nvec == m) break;
        } else if (breakdown || nvec == m) {
            break;                                                  /* nvec == 1 */
        }
        /* next vector */
        const double rb = 1.0 / bj;
        double *qn = Q + (size_t)(j + 1) * m;
        for (int e = tid; e < m; e += T) { const double v = s.w[e] * rb; s.q[e] = v; qn[e] = v; }
        __syncthreads();
    }
    if (steps_out && tid == 0) *steps_out = nvec;
#ifndef FPT_EMU
    if (tid == 0) atomicAdd(&fpt_lanczos_phase_cycles[7], (unsigned long long)nvec);     /* slot 7: Lanczos steps, not cycles */
#endif
    FPT_LZ_MARK(5);

    /* ---- 3. coordinates */
    if (degenerate || !have) {
        /* B = 0 gives X = 0 as the one-warp path does; non-finite input gives NaN; a 1 x 1 Krylov space (m == 1, or q0 an
           eigenvector) has one Ritz pair: alpha_0 along q0 */
        double v = degenerate == 2 ? bnorm - bnorm : 0.0;
        if (!degenerate && nvec == 1) {
            double a0 = s.alpha[0];
            if (a0 < 0.0 && -a0 <= 1e-13 * bnorm) a0 = 0.0;
            const double r = sqrt(a0);
            for (int e = tid; e < m; e += T) { X[2 * e] = s.q[e] * r; X[2 * e + 1] = 0.0; }
            if (tid == 0 && evals3) { evals3[0] = s.alpha[0]; evals3[1] = 0.0; evals3[2] = 0.0; }
        } else {
            for (int e = tid; e < 2 * m; e += T) X[e] = v;
            if (tid == 0 && evals3) { evals3[0] = v; evals3[1] = v; evals3[2] = v; }
        }
        __syncthreads();
        return;
    }
    const double tnorm = sh_val[3];
    if (lam1 < 0.0 && -lam1 <= 1e-13 * tnorm) lam1 = 0.0;
    if (lam2 < 0.0 && -lam2 <= 1e-13 * fmax(fabs(lam1), tnorm * 1e-3)) lam2 = 0.0;
    const double r1 = sqrt(lam1), r2 = sqrt(lam2);
    const double *y0 = s.ew.y, *y1 = s.ew.y + nvec;
    for (int e = tid; e < m; e += T) {
        double a0 = 0.0, a1 = 0.0;
        for (int i = 0; i < nvec; i++) { const double qv = Q[(size_t)i * m + e]; a0 += qv * y0[i]; a1 += qv * y1[i]; }
        X[2 * e] = a0 * r1; X[2 * e + 1] = a1 * r2;
    }
    if (tid == 0 && evals3) { evals3[0] = lam1; evals3[1] = lam2; evals3[2] = lam3; }
    __syncthreads();
    FPT_LZ_MARK(6);
}

/* A (m x m, global) holds the filled dissimilarities on entry and B on exit; Q (>= cap x m doubles, global) receives the
   Lanczos basis. X: 2m doubles (shared or global) written by the CTA; evals3 optional. All threads of the CTA take part. */
FPT_D void fpt_css_cmds_lanczos(double *A, double *Q, int m, double *X, double *evals3, const FptLanczosSmem &s, int *steps_out,
                                int max_form = 2) {
    const int T = blockDim.x, tid = threadIdx.x;
    const size_t mm = (size_t)m * m;
    FPT_LZ_START();
    /* ---- 1. row means of S = D.D and the grand mean (same closed form as the small-cohort paths); on the way, whether
       the matrix has the compact form: every entry a count in 1..65535 or the fill value v0 (the diagonal always is) */
    const double v0 = A[0];
    int compact = (m & 3) == 0, narrow = 1;                     /* narrow: every count fits a byte */
    {
        const int lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
        for (int i = warp; i < m; i += nwarp) {                /* row sums, coalesced */
            const double *row = A + (size_t)i * m;
            double acc = 0.0;
            for (int j = lane; j < m; j += 32) {
                const double d = row[j];
                acc += d * d;
                if (!(d == v0 || (d >= 1.0 && d <= 65535.0 && d == floor(d)))) compact = 0;
                if (d > 255.0) narrow = 0;
            }
            acc = fpt_warp_sum(acc);
            if (lane == 0) s.rmean[i] = acc / m;
        }
    }
    compact = __syncthreads_and(compact);
    narrow = __syncthreads_and(narrow);
    /* product form: 2 8-bit codes, 1 16-bit codes, 0 fp64 matrix; `max_form` caps it (parity tests) */
    const int ok2 = compact && narrow && (m & 7) == 0, ok1 = compact;
    const int form = (ok2 && max_form >= 2) ? 2 : ((ok1 && max_form >= 1) ? 1 : 0);
    narrow = form == 2;
    compact = form > 0;
    double g = 0.0;
    for (int i = tid; i < m; i += T) g += s.rmean[i];
    g = fpt_block_sum(g, s.sc.red) / m;
    unsigned short *codes = reinterpret_cast<unsigned short *>(A);
    unsigned char *codes8 = reinterpret_cast<unsigned char *>(A);
    const double v2 = v0 * v0;
    if (compact) {
        /* in place: the codes of a block of 8 T entries land in the first quarter of the bytes those entries occupied,
           which only the block itself has still to read — hence the barrier between its loads and its stores */
        for (size_t base = 0; base < mm; base += (size_t)8 * T) {
            double d[8];
#pragma unroll
            for (int u = 0; u < 8; u++) { const size_t e = base + tid + (size_t)u * T; d[u] = e < mm ? A[e] : 0.0; }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < 8; u++) {
                const size_t e = base + tid + (size_t)u * T;
                if (e < mm) {
                    const unsigned c = (d[u] >= 1.0 && d[u] == floor(d[u])) ? (unsigned)d[u] : 0u;
                    if (narrow) codes8[e] = (unsigned char)c; else codes[e] = (unsigned short)c;
                }
            }
        }
    } else {
        for (size_t e = tid; e < mm; e += T) {
            const int i = (int)(e / m), j = (int)(e - (size_t)i * m);
            const double d = A[e];
            A[e] = -0.5 * (((d * d - s.rmean[i]) - s.rmean[j]) + g);
        }
    }
    __syncthreads();
    FPT_LZ_MARK(1);
    FptLzMatrix M;
    M.form = form; M.ld = m; M.data = A; M.g = g; M.v2 = v2; M.zlist = nullptr;
    fpt_lanczos_iterate<0>(M, Q, m, X, evals3, s, steps_out);
}

/* Large cohorts, default route. The window's count codes were written by fpt_css_k4_umma_kernel / fpt_css_k4_popc_kernel
   (fpt_css_k4.cuh); here fill_averages (css.c:337-366) and the double centring of cmds (css.c:505-531) are derived from them in
   integers — blanks = codes 0, their replacement = (sum of the counts) / m^2, row sums of squares exact — and the Lanczos
   iteration streams the codes. ARITH = true: the kernel of the arithmetic product — windows of at most 255 SNPs whose blank entries fit
   the list; any other window is marked FPT_WIN_PENDING for a second launch with ARITH = false, only_pending = 1 (table / 16-bit
   products). basis: per CTA fpt_lanczos_cta_scratch_bytes(m). THREADS: 512 (64 registers per thread) or 384 (80:
   room for the software-pipelined loads of the product without spills); two CTAs per SM either way. */
#define FPT_WIN_PENDING 3     /* code route, between its two kernels: left by the arithmetic kernel for the table / 16-bit kernel */

template <int THREADS, bool ARITH>
__global__ void __launch_bounds__(THREADS, 2)
fpt_css_mds_codes_kernel(const unsigned char *__restrict__ codes, size_t stride, int m, const int *__restrict__ wleft,
                         const int *__restrict__ wright, long long nwin, double *__restrict__ basis, double *__restrict__ Xout,
                         double *__restrict__ evals_out, unsigned char *__restrict__ status, int *__restrict__ steps_out,
                         int only_pending) {
    FPT_DYN_SMEM(smem);
    const FptLanczosSmem s = fpt_lanczos_carve(smem, m, 0);
    const int T = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = T >> 5;
    const int ld = (m + 15) & ~15;
    double *Q = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(basis) + (size_t)blockIdx.x * fpt_lanczos_cta_scratch_bytes(m));
    int *zlist = reinterpret_cast<int *>(Q + (size_t)fpt_lanczos_cap(m) * m);
    for (long long w = blockIdx.x; w < nwin; w += gridDim.x) {
        if (!ARITH && only_pending && status[w] != FPT_WIN_PENDING) continue;
        const int l = wleft[w], r = wright[w];
        if (r <= l) { if (tid == 0) status[w] = FPT_WIN_EMPTY; continue; }
        FPT_LZ_START();
        const int wide = (r - l) > 255;                           /* two-byte codes (fpt_k4_code_bytes) */
        if (ARITH && wide) { if (tid == 0) status[w] = FPT_WIN_PENDING; continue; }
        const unsigned char *c8 = codes + (size_t)w * stride;
        const unsigned short *c16 = reinterpret_cast<const unsigned short *>(c8);
        long long blanks = 0, total = 0;
        for (int i = warp; i < m; i += nwarp) {                   /* a warp per row, 16 bytes per lane and load */
            unsigned long long sq = 0;
            unsigned sum = 0, nb = 0;
            if (!wide) {
                for (int j = 16 * lane; j < m; j += 512) {
                    const uint4 v = *reinterpret_cast<const uint4 *>(c8 + (size_t)i * ld + j);
                    const unsigned ww[4] = { v.x, v.y, v.z, v.w };
                    const int valid = min(16, m - j);
#pragma unroll
                    for (int k = 0; k < 16; k++) {
                        const unsigned c = (ww[k >> 2] >> (8 * (k & 3))) & 0xffu;
                        sum += c; sq += c * c; nb += (k < valid && c == 0u);
                    }
                }
            } else {
                for (int j = 8 * lane; j < m; j += 256) {
                    const uint4 v = *reinterpret_cast<const uint4 *>(c16 + (size_t)i * ld + j);
                    const unsigned ww[4] = { v.x, v.y, v.z, v.w };
                    const int valid = min(8, m - j);
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const unsigned c = (ww[k >> 1] >> (16 * (k & 1))) & 0xffffu;
                        sum += c; sq += (unsigned long long)c * c; nb += (k < valid && c == 0u);
                    }
                }
            }
            long long sq_w = fpt_warp_sum_i64((long long)sq), sum_w = fpt_warp_sum_i64((long long)sum), nb_w = fpt_warp_sum_i64((long long)nb);
            sq_w = __shfl_sync(FPT_FULL_MASK, sq_w, 0); sum_w = __shfl_sync(FPT_FULL_MASK, sum_w, 0); nb_w = __shfl_sync(FPT_FULL_MASK, nb_w, 0);
            if (lane == 0) { s.rmean[i] = (double)sq_w; s.w[i] = (double)nb_w; blanks += nb_w; total += sum_w; }
        }
        blanks = fpt_block_sum_i64(blanks, s.sc.redi);
        total = fpt_block_sum_i64(total, s.sc.redi);
        const long long mm = (long long)m * m;
        if (blanks > mm / 2) { if (tid == 0) status[w] = FPT_WIN_DISCARDED; __syncthreads(); continue; }   /* css.c:355 */
        const double avg = __ddiv_rn((double)total, (double)mm);                                            /* css.c:357 */
        const double v2 = avg * avg;
        for (int i = tid; i < m; i += T) s.rmean[i] = (s.rmean[i] + s.w[i] * v2) / m;
        __syncthreads();
        double g = 0.0;
        for (int i = tid; i < m; i += T) g += s.rmean[i];
        g = fpt_block_sum(g, s.sc.red) / m;
        /* the blank entries (code 0, the diagonal among them) as a list of column indices, row by row in column order: the
           arithmetic form of the product skips them and adds their fill value per row. One more pass over the codes. */
        const bool zl_on = ARITH && blanks <= (long long)fpt_lanczos_zcap(m);
        if (ARITH && !zl_on) { if (tid == 0) status[w] = FPT_WIN_PENDING; __syncthreads(); continue; }
        if (zl_on) {
            if (warp == 0) {                                          /* exclusive scan of the rows' blank counts (s.w) */
                const int per = (m + 31) >> 5, b0 = lane * per, b1 = min(m, b0 + per);
                int loc = 0;
                for (int i = b0; i < b1; i++) loc += (int)s.w[i];
                int inc = loc;
                for (int o = 1; o < 32; o <<= 1) { const int yv = __shfl_up_sync(FPT_FULL_MASK, inc, o); if (lane >= o) inc += yv; }
                int run = inc - loc;
                for (int i = b0; i < b1; i++) { s.zoff[i] = run; run += (int)s.w[i]; }
                if (lane == 31) s.zoff[m] = inc;
            }
            __syncthreads();
            for (int i = warp; i < m; i += nwarp) {
                int base = s.zoff[i];
                for (int j0 = 0; j0 < m; j0 += 512) {
                    const int j = j0 + 16 * lane;
                    unsigned zmask = 0u;
                    if (j < m) {
                        const uint4 v = *reinterpret_cast<const uint4 *>(c8 + (size_t)i * ld + j);
                        const unsigned ww[4] = { v.x, v.y, v.z, v.w };
                        const int valid = min(16, m - j);
#pragma unroll
                        for (int k = 0; k < 16; k++) if (k < valid && ((ww[k >> 2] >> (8 * (k & 3))) & 0xffu) == 0u) zmask |= 1u << k;
                    }
                    const int cntl = __popc(zmask);
                    int inc = cntl;
                    for (int o = 1; o < 32; o <<= 1) { const int yv = __shfl_up_sync(FPT_FULL_MASK, inc, o); if (lane >= o) inc += yv; }
                    int at = base + inc - cntl;
                    while (zmask) { const int k = __ffs((int)zmask) - 1; zmask &= zmask - 1u; zlist[at++] = j + k; }
                    base += __shfl_sync(FPT_FULL_MASK, inc, 31);
                }
            }
            __syncthreads();
        }
        FPT_LZ_MARK(1);
        FptLzMatrix M;
        M.form = wide ? 1 : 2; M.ld = ld; M.data = c8; M.g = g; M.v2 = v2; M.zlist = zl_on ? zlist : nullptr;
        fpt_lanczos_iterate<ARITH ? 3 : 0>(M, Q, m, Xout + (size_t)w * 2 * m, evals_out ? evals_out + 3 * w : 0, s, steps_out ? steps_out + w : 0);
        if (tid == 0) status[w] = FPT_WIN_SCORED;
        __syncthreads();
    }
}

/* Legacy route (fpt_set_k4_mode(0), the frequency metric, windows of more than 65535 SNPs): dissimilarities as an fp64 matrix.
   mds 0 (and the first half of mds 2) for large cohorts: one CTA walks windows blockIdx.x, +gridDim.x, ...;
   gscratch: per CTA fpt_css_mats_doubles(m) doubles (B, then the Lanczos basis) */
__global__ void __launch_bounds__(512, 2)
fpt_css_mds_large_kernel(const unsigned *__restrict__ planes, const double *__restrict__ absdiff, int m,
                         const int *__restrict__ wleft, const int *__restrict__ wright, long long nwin, int wch,
                         double *__restrict__ gscratch, double *__restrict__ Xout, double *__restrict__ evals_out,
                         unsigned char *__restrict__ status, int *__restrict__ steps_out, int max_form = 2) {
    FPT_DYN_SMEM(smem);
    const FptLanczosSmem s = fpt_lanczos_carve(smem, m, wch);
    double *M0 = gscratch + (size_t)blockIdx.x * fpt_css_mats_doubles(m), *M1 = M0 + (size_t)m * m;
    for (long long w = blockIdx.x; w < nwin; w += gridDim.x) {
        const int l = wleft[w], r = wright[w];
        if (r <= l) { if (threadIdx.x == 0) status[w] = FPT_WIN_EMPTY; continue; }
        FPT_LZ_START();
        const int keep = fpt_css_dissimilarity(planes, absdiff, m, l, r, M0, s.sc);
        FPT_LZ_MARK(0);                                             /* slot 0: compare_all + fill_averages */
        if (!keep) { if (threadIdx.x == 0) status[w] = FPT_WIN_DISCARDED; __syncthreads(); continue; }
        fpt_css_cmds_lanczos(M0, M1, m, Xout + (size_t)w * 2 * m, evals_out ? evals_out + 3 * w : 0, s, steps_out ? steps_out + w : 0, max_form);
        if (threadIdx.x == 0) status[w] = FPT_WIN_SCORED;
        __syncthreads();
    }
}

#endif
