/*
 * fpt_css_eig_reg.cuh — phase A of the one-warp classical MDS (fpt_css_eig.cuh) for the cohorts the genome scans are made of
 * (3 <= m <= 48): opposite-homozygote counts, fill_averages, double centring and the Householder tridiagonalisation of
 * css/css.c:277-366,505-531 (cmds up to GSL's symmetric reduction) with the whole m x m matrix in the REGISTERS of the warp.
 *
 * Why (ncu of fpt_css_tridiag_kernel, profiles/r1_css_tridiag_raw.csv): 47 k warp instructions per 40 x 40 window with 15.5
 * of 32 lanes active on average — lane = row over a packed triangle in shared memory leaves most lanes idle once the active
 * block is narrower than the warp, and every multiply-add pays two shared-memory loads, an index update and a loop branch.
 *
 * Here the 32 lanes form an 8 x 4 grid (rg = lane / 4, cg = lane % 4) and lane (rg, cg) owns the elements
 * (8 r + rg, 4 c + cg), r < RS, c < CS of the full (both triangles) matrix, padded with zeros to 8 RS = 4 CS: a cyclic layout,
 * so the trailing block of a Householder step stays spread over all 32 lanes however small it gets. Per step:
 *   - column k goes through shared memory once (the owners store it), its norm is a warp reduction, and the reflector v built from
 *     it (0 up to row k, 1 at k + 1, the scaled column below) replaces it there: every lane reads its rows and its columns of v,
 *     no selects in the step; v, d_k, e_k and tau_k stay in shared memory until the window is done and leave in one coalesced sweep;
 *   - p = tau A v is RS x CS register multiply-adds per lane and a two-stage butterfly over the four column groups;
 *   - v'p is a three-stage butterfly over the row groups; w = p - (tau/2)(v'p) v goes through shared memory for its column copy;
 *   - A -= v w' + w v' is 2 RS CS register multiply-adds.
 * Row slots (8 rows) and column slots (4 columns) that lie wholly above / left of the active block are skipped at compile time:
 * the step body is instantiated per first live row slot R0 (columns from slot 2 R0), so the work shrinks with the block.
 * The reflector scale comes from rsqrt and a Newton reciprocal instead of sqrt and two divisions (tau v'v = 2 holds to
 * rounding, which is what the orthogonality of H needs). Same outputs as fpt_css_tridiag_kernel — d, e, tau and the reflectors
 * in the fpt_refl_col layout — so fpt_css_eigvec_kernel follows unchanged; the values differ from that kernel's in rounding only
 * (different summation order), pinned against the oracle by the same tests.
 */
#ifndef FPT_CSS_EIG_REG_CUH
#define FPT_CSS_EIG_REG_CUH

#include "fpt_css_eig.cuh"

#define FPT_TREG_WARPS 4

FPT_HD int fpt_tridiag_reg_ok(int m) { return m >= 3 && m <= 48; }
FPT_HD int fpt_tridiag_reg_pad(int m) { return m <= 32 ? 32 : (m <= 40 ? 40 : 48); }
/* per warp: two vectors of `pad` doubles (column / v, w) — the same bytes first stage two 32-SNP words of both bit-planes —
   then d, e, tau (3 pad) and the reflectors (pad (pad - 1) / 2): they are collected in shared memory and leave for the hand-over
   buffer in one coalesced sweep per window (per-step scalar stores cost a 64-bit address computation each, ~10 % of the loop) */
FPT_HD size_t fpt_tridiag_reg_work_bytes(int m) {
    const size_t pad = (size_t)fpt_tridiag_reg_pad(m);
    return 8 * (2 * pad + 3 * pad + pad * (pad - 1) / 2);
}

/* column k of the register matrix (every row) to vb[row]: the lanes of column group k % 4 hold it in column slot k / 4 */
template <int RS, int CS>
FPT_D void fpt_treg_dump_col(const double (&a)[RS][CS], int k, int rg, int cg, double *vb) {
    if (cg == (k & 3)) {
#define FPT_TREG_CASE(c_)                                                     \
    case c_:                                                                  \
        if (c_ < CS) {                                                        \
            _Pragma("unroll") for (int r = 0; r < RS; r++) vb[8 * r + rg] = a[r][c_ < CS ? c_ : 0]; \
        }                                                                     \
        break;
        switch (k >> 2) {
            FPT_TREG_CASE(0) FPT_TREG_CASE(1) FPT_TREG_CASE(2) FPT_TREG_CASE(3) FPT_TREG_CASE(4) FPT_TREG_CASE(5)
            FPT_TREG_CASE(6) FPT_TREG_CASE(7) FPT_TREG_CASE(8) FPT_TREG_CASE(9) FPT_TREG_CASE(10) FPT_TREG_CASE(11)
            default: break;
        }
#undef FPT_TREG_CASE
    }
}

/* Householder step on the block of row slots >= R0, column slots >= 2 R0 (everything before them lies above / left of row and
   column k + 1). vb holds the reflector v itself (0 up to row k, 1 at k + 1, the scaled column below; rows >= m read zero).
   Rows and columns <= k inside the live slots are NOT masked: they are never read again as matrix entries, their v is zero (so they
   add nothing to p of a live row, nor to v'p), and what the step does to them is the same orthogonal update as everywhere else —
   they stay finite, and skipping the masks saves a compare and two selects per element and step. */
template <int RS, int CS, int R0>
FPT_D void fpt_treg_step(double (&a)[RS][CS], double tau, int rg, int cg, const double *vb, double *wb) {
    constexpr int C0 = 2 * R0;
    double vr[RS], vc[CS], p[RS];
#pragma unroll
    for (int r = R0; r < RS; r++) vr[r] = vb[8 * r + rg];
#pragma unroll
    for (int c = C0; c < CS; c++) vc[c] = vb[4 * c + cg];
    /* p = tau A v: my columns' share of my rows, then the other three column groups' */
#pragma unroll
    for (int r = R0; r < RS; r++) {
        double s = 0.0, s_b = 0.0;
#pragma unroll
        for (int c = C0; c + 1 < CS; c += 2) { s = fma(a[r][c], vc[c], s); s_b = fma(a[r][c + 1], vc[c + 1], s_b); }
        p[r] = s + s_b;
    }
#pragma unroll
    for (int r = R0; r < RS; r++) p[r] += __shfl_xor_sync(FPT_FULL_MASK, p[r], 1);
#pragma unroll
    for (int r = R0; r < RS; r++) p[r] += __shfl_xor_sync(FPT_FULL_MASK, p[r], 2);
    double pv = 0.0;
#pragma unroll
    for (int r = R0; r < RS; r++) {
        p[r] *= tau;
        pv = fma(p[r], vr[r], pv);
    }
    pv += __shfl_xor_sync(FPT_FULL_MASK, pv, 4);
    pv += __shfl_xor_sync(FPT_FULL_MASK, pv, 8);
    pv += __shfl_xor_sync(FPT_FULL_MASK, pv, 16);
    const double K = -0.5 * tau * pv;
#pragma unroll
    for (int r = R0; r < RS; r++) p[r] = fma(K, vr[r], p[r]);                 /* w, by row */
    if (cg == 0) {
#pragma unroll
        for (int r = R0; r < RS; r++) wb[8 * r + rg] = p[r];
    }
    __syncwarp();
    /* A -= v w' + w v' */
#pragma unroll
    for (int c = C0; c < CS; c++) {
        const double wc = wb[4 * c + cg], vcc = vc[c];
#pragma unroll
        for (int r = R0; r < RS; r++) a[r][c] = fma(-p[r], vcc, fma(-vr[r], wc, a[r][c]));
    }
}

template <int RS, int CS>
__global__ void __launch_bounds__(32 * FPT_TREG_WARPS, RS <= 4 ? 4 : (RS == 5 ? 3 : 2))
fpt_css_tridiag_reg_kernel(const unsigned *__restrict__ planes, int m, const int *__restrict__ wleft, const int *__restrict__ wright,
                           long long nwin, double *__restrict__ tri_out, double *__restrict__ refl_out,
                           unsigned char *__restrict__ status) {
    static_assert(CS == 2 * RS, "the lane grid is 8 x 4: 8 RS rows = 4 CS columns");
    constexpr int MP = 8 * RS;
    FPT_DYN_SMEM(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
    const int rg = lane >> 2, cg = lane & 3;
    double *vb = reinterpret_cast<double *>(smem + (size_t)warp * (8 * (5 * MP + MP * (MP - 1) / 2))), *wb = vb + MP;
    double *ts = wb + MP, *rs = ts + 3 * MP;              /* d | e | tau at stride m, reflectors in the fpt_refl_col layout */
    unsigned *words = reinterpret_cast<unsigned *>(vb);
    const size_t nrefl = (size_t)m * (m - 1) / 2;
    const int mm = m * m;
    const double dm = (double)m;
#pragma unroll 1
    for (long long win = (long long)blockIdx.x * nwarp + warp; win < nwin; win += (long long)gridDim.x * nwarp) {
        const int l = wleft[win], rr = wright[win];
        if (rr <= l) { if (lane == 0) status[win] = 0; continue; }
        double a[RS][CS];
        {   /* ---- compare_all (css.c:277-327) on the bit-planes: both triangles, two 32-SNP words per pass */
            int cnt[RS][CS];
#pragma unroll
            for (int r = 0; r < RS; r++)
#pragma unroll
                for (int c = 0; c < CS; c++) cnt[r][c] = 0;
            const int w0 = l >> 5, w1 = (rr - 1) >> 5;
#pragma unroll 1
            for (int wc = w0; wc <= w1; wc += 2) {
                const int nw = min(2, w1 - wc + 1);
                __syncwarp();
#pragma unroll 1
                for (int e = lane; e < nw * 2 * m; e += 32) {
                    const int ww = wc + (e >= 2 * m ? 1 : 0);
                    unsigned mask = 0xffffffffu;
                    if (ww == w0) mask &= 0xffffffffu << (l & 31);
                    if (ww == w1) mask &= 0xffffffffu >> (31 - ((rr - 1) & 31));
                    words[e] = planes[(size_t)wc * 2 * m + e] & mask;
                }
                __syncwarp();
#pragma unroll 1
                for (int q = 0; q < nw; q++) {
                    const unsigned *row = words + (size_t)q * 2 * m;
                    unsigned pr[RS], mr[RS], pc[CS], mc[CS];
#pragma unroll
                    for (int r = 0; r < RS; r++) {
                        const int i = 8 * r + rg;
                        const bool ok = i < m;
                        pr[r] = ok ? row[i] : 0u; mr[r] = ok ? row[m + i] : 0u;
                    }
#pragma unroll
                    for (int c = 0; c < CS; c++) {
                        const int j = 4 * c + cg;
                        const bool ok = j < m;
                        pc[c] = ok ? row[j] : 0u; mc[c] = ok ? row[m + j] : 0u;
                    }
#pragma unroll
                    for (int r = 0; r < RS; r++)
#pragma unroll
                        for (int c = 0; c < CS; c++) cnt[r][c] += __popc(pr[r] & mc[c]) + __popc(mr[r] & pc[c]);
                }
            }
            /* ---- fill_averages (css.c:337-366): blanks (the diagonal always is) take the mean of all m*m entries */
            int blanks = 0;
            long long sum = 0;
#pragma unroll
            for (int r = 0; r < RS; r++)
#pragma unroll
                for (int c = 0; c < CS; c++) {
                    const bool valid = (8 * r + rg < m) && (4 * c + cg < m);
                    blanks += (valid && cnt[r][c] == 0) ? 1 : 0;
                    sum += cnt[r][c];
                }
            for (int o = 16; o > 0; o >>= 1) blanks += __shfl_xor_sync(FPT_FULL_MASK, blanks, o);
            sum = fpt_warp_sum_i64(sum);
            if (blanks > mm / 2) { if (lane == 0) status[win] = 1; continue; }
            const double avg = __ddiv_rn((double)sum, (double)mm);
#pragma unroll
            for (int r = 0; r < RS; r++)
#pragma unroll
                for (int c = 0; c < CS; c++) {
                    const bool valid = (8 * r + rg < m) && (4 * c + cg < m);
                    const double v = cnt[r][c] ? (double)cnt[r][c] : avg;
                    a[r][c] = valid ? v * v : 0.0;
                }
        }
        /* ---- double centring (css.c:505-531): B = -1/2 (S - r 1' - 1 r' + g) */
        {
            double rmean[RS], g = 0.0;
#pragma unroll
            for (int r = 0; r < RS; r++) {
                double s = 0.0;
#pragma unroll
                for (int c = 0; c < CS; c++) s += a[r][c];
                s += __shfl_xor_sync(FPT_FULL_MASK, s, 1);
                s += __shfl_xor_sync(FPT_FULL_MASK, s, 2);
                rmean[r] = s / dm;
                g += rmean[r];
            }
            g += __shfl_xor_sync(FPT_FULL_MASK, g, 4);
            g += __shfl_xor_sync(FPT_FULL_MASK, g, 8);
            g += __shfl_xor_sync(FPT_FULL_MASK, g, 16);
            g /= dm;
            __syncwarp();                                    /* the bit-plane words are done with */
            if (cg == 0) {
#pragma unroll
                for (int r = 0; r < RS; r++) vb[8 * r + rg] = rmean[r];
            }
            __syncwarp();
#pragma unroll
            for (int c = 0; c < CS; c++) {
                const double rc = vb[4 * c + cg];
#pragma unroll
                for (int r = 0; r < RS; r++) {
                    const bool valid = (8 * r + rg < m) && (4 * c + cg < m);
                    a[r][c] = valid ? -0.5 * (((a[r][c] - rmean[r]) - rc) + g) : 0.0;
                }
            }
            __syncwarp();
        }
        /* ---- Householder tridiagonalisation (LAPACK dsytd2, lower): H_k = I - tau v v', v(k+1) = 1 */
#pragma unroll 1
        for (int k = 0; k + 2 < m; k++) {
            fpt_treg_dump_col<RS, CS>(a, k, rg, cg, vb);
            __syncwarp();
            const double dk = vb[k], x0 = vb[k + 1];
            const double xl = vb[lane], xh = lane + 32 < MP ? vb[lane + 32] : 0.0;
            double s2 = (lane >= k + 2 ? xl * xl : 0.0) + (lane + 32 >= k + 2 ? xh * xh : 0.0);
            s2 = fpt_warp_sum(s2);
            if (lane == 0) ts[k] = dk;
            if (s2 == 0.0) {                                 /* column already tridiagonal: H = I */
                if (lane == 0) { ts[m + k] = x0; ts[2 * m + k] = 0.0; }
                __syncwarp();
                continue;
            }
            const double n2 = fma(x0, x0, s2), ax0 = fabs(x0);
            double nrm, tau, scal;
            if (n2 > 1e-60 && n2 < 1e60) {
                const double rn = rsqrt(n2);
                nrm = n2 * rn;
                tau = fma(ax0, rn, 1.0);                     /* (alpha - x0) / alpha, alpha = -sign(x0) nrm */
                const double rc = fpt_fast_rcp(ax0 + nrm);
                scal = x0 >= 0.0 ? rc : -rc;                 /* 1 / (x0 - alpha) */
            } else {
                nrm = sqrt(n2);
                tau = (ax0 + nrm) / nrm;
                scal = (x0 >= 0.0 ? 1.0 : -1.0) / (ax0 + nrm);
            }
            const double alpha = x0 >= 0.0 ? -nrm : nrm;
            {   /* the reflector v (0 up to row k, 1 at k + 1, the scaled column below): to the hand-over buffer (rcol[i] = v_i,
                   i = k+1 .. m-1) and, in place of the column, to vb, where the step reads its row and column copies */
                const double vl = lane >= k + 2 ? xl * scal : (lane == k + 1 ? 1.0 : 0.0);
                const double vh = lane + 32 >= k + 2 ? xh * scal : (lane + 32 == k + 1 ? 1.0 : 0.0);
                double *rcol = rs + fpt_refl_col(m, k) - (k + 1);
                if (lane > k && lane < m) rcol[lane] = vl;
                if (lane + 32 > k && lane + 32 < m) rcol[lane + 32] = vh;
                __syncwarp();                                /* every lane has read the column */
                vb[lane] = vl;
                if (lane + 32 < MP) vb[lane + 32] = vh;
                __syncwarp();
            }
            if (lane == 0) { ts[m + k] = alpha; ts[2 * m + k] = tau; }
            switch ((k + 1) >> 3) {
                case 0: fpt_treg_step<RS, CS, 0>(a, tau, rg, cg, vb, wb); break;
                case 1: fpt_treg_step<RS, CS, 1>(a, tau, rg, cg, vb, wb); break;
                case 2: fpt_treg_step<RS, CS, 2>(a, tau, rg, cg, vb, wb); break;
                case 3: fpt_treg_step<RS, CS, 3>(a, tau, rg, cg, vb, wb); break;
                case 4: if (RS > 4) fpt_treg_step<RS, CS, (RS > 4 ? 4 : 0)>(a, tau, rg, cg, vb, wb); break;
                default: if (RS > 5) fpt_treg_step<RS, CS, (RS > 5 ? 5 : 0)>(a, tau, rg, cg, vb, wb); break;
            }
        }
        /* the last 2 x 2 block: d[m-2], e[m-2], d[m-1] */
        fpt_treg_dump_col<RS, CS>(a, m - 2, rg, cg, vb);
        __syncwarp();
        const double d2 = vb[m - 2], e2 = vb[m - 1];
        __syncwarp();
        fpt_treg_dump_col<RS, CS>(a, m - 1, rg, cg, vb);
        __syncwarp();
        if (lane == 0) {
            ts[m - 2] = d2; ts[m + m - 2] = e2; ts[m - 1] = vb[m - 1];
            ts[m + m - 1] = 0.0; ts[2 * m + m - 1] = 0.0; ts[2 * m + m - 2] = 0.0;   /* entries the reduction does not produce */
            status[win] = 2;
        }
        __syncwarp();
        {   /* hand-over: one coalesced sweep (a column whose reflector was the identity keeps stale entries: tau = 0 says so) */
            double *t = tri_out + (size_t)win * 3 * m;
            double *refl = refl_out + (size_t)win * nrefl;
            for (int e = lane; e < 3 * m; e += 32) t[e] = ts[e];
            for (int e = lane; e < (int)nrefl; e += 32) refl[e] = rs[e];
        }
        __syncwarp();
    }
}

#endif
