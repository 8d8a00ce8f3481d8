/*
 * fpt_css_perm_umma.cuh — the large-cohort permutation test on the 5th-generation tensor cores (250 < m <= 1024,
 * independent shuffles; BASELINE configs[4] has m = 1000).
 *
 * Reference: calc_dist css/css.c:573-587, css css.c:608-647, significance_treshold / random_shuffle css.c:700-752
 * (paths relative to /root/reference/statistics/).
 *
 * Same decisions as fpt_css_perm_large.cuh — the exact integer surrogate, the proven error bound E, the reference-order
 * fp64 re-scoring of the rare permutation within E of the observed score — but the between-group sums of a batch of 128
 * permutations are ONE dense contraction on `tcgen05.mma kind::i8`:
 *
 *     D_d[p][n] = sum_k Z[p][k] * Q_d[n][k]        Z: 128 x K membership rows (0/1, u8), built in shared memory
 *                                                  Q_d: digit d (base 256) of the quantised distance matrix, u8
 *     bet[p]    = sum_d 256^d sum_{n : Z[p][n] = 0} D_d[p][n]
 *
 * Z is the A operand (shared memory, K-major core-matrix layout, fpt_umma.cuh), the digit matrices are written by the
 * distance pass straight in that layout to the CTA's global scratch and streamed through a two-stage shared-memory ring by
 * bulk TMA copies (one warp), a second warp issues the MMAs (128 x 256 x 32 per instruction, s32 accumulators in tensor
 * memory, two 256-column accumulators so that a tile is drained while the next one is computed), and four warps drain the
 * accumulators with `tcgen05.ld`, mask them with the membership row of their own lane and keep four u32 sums.
 */
#ifndef FPT_CSS_PERM_UMMA_CUH
#define FPT_CSS_PERM_UMMA_CUH

#include "fpt_css_perm_large.cuh"
#include "fpt_css_observed.cuh"
#include "fpt_umma.cuh"

#define FPT_UMMA_THREADS 512
#define FPT_UMMA_BATCH 128                 /* permutations per contraction = MMA M */
#define FPT_UMMA_NT 256                    /* MMA N: columns of one accumulator */
#define FPT_UMMA_KC 128                    /* bytes of K per ring stage */
#define FPT_UMMA_STAGE (FPT_UMMA_NT * FPT_UMMA_KC)
#define FPT_UMMA_DIGITS 4
#define FPT_UMMA_LRING (2 * FPT_UMMA_BATCH) /* label rows kept in the CTA's global scratch: shuffle rounds run ahead of the batches */

FPT_HD int fpt_umma_kp(int m) { return ((m + FPT_UMMA_KC - 1) / FPT_UMMA_KC) * FPT_UMMA_KC; }
FPT_HD int fpt_umma_np(int m) { return ((m + FPT_UMMA_NT - 1) / FPT_UMMA_NT) * FPT_UMMA_NT; }
FPT_HD int fpt_umma_rbytes(int m) { return (int)((((size_t)m * 2 + 3) >> 2) | 1) << 2; }     /* label row, odd word count */

/* per-CTA global scratch: digit matrices in tile layout, label rows of one batch */
FPT_HD size_t fpt_umma_scratch_bytes(int m) {
    size_t b = (size_t)FPT_UMMA_DIGITS * fpt_umma_np(m) * fpt_umma_kp(m);
    b += ((size_t)FPT_UMMA_LRING * m * 2 + 255) & ~(size_t)255;
    return b;
}
FPT_HD size_t fpt_umma_smem_bytes(int m) {
    size_t off = 128;                                           /* slack for the 128-byte alignment of the tile space */
    off += (size_t)FPT_UMMA_BATCH * fpt_umma_kp(m) + 2 * FPT_UMMA_STAGE;
    off += (size_t)2 * m * 8;                                   /* X */
    off += (size_t)(m + 1) * 8;                                 /* per-n (limit, magic) of the shuffle draws */
    off += (size_t)2 * FPT_UMMA_LRING * 8;                      /* adjacent-pair sums */
    off += (size_t)(FPT_UMMA_THREADS / 32) * 32 * 8;            /* per-warp staging of the exact re-scoring */
    off += (size_t)FPT_UMMA_BATCH * 4 + 33 * 4 + 16;            /* hits, scan */
    return (off + 15) & ~(size_t)15;
}


/* diagnostic: SM cycles thread 0 spent per phase, summed over CTAs and windows (distance pass, observed score, shuffles,
   membership rows, contraction, decisions); read and reset by fpt_debug_umma_phases() */
__device__ unsigned long long fpt_umma_phase_cycles[8];
#define FPT_UMMA_MARK(slot) do { if (tid == 0) { const long long now_ = clock64(); atomicAdd(&fpt_umma_phase_cycles[slot], (unsigned long long)(now_ - t_mark)); t_mark = now_; } } while (0)

__global__ void __launch_bounds__(FPT_UMMA_THREADS, 1)
fpt_css_perm_umma_kernel(const double *__restrict__ Xall, int m, int asize, int bsize, long long wbase, long long nwin,
                         const unsigned char *__restrict__ status, int treshold, int runs, uint64_t seed,
                         const uint64_t *__restrict__ state_override, unsigned char *__restrict__ gscratch,
                         size_t gscratch_per_cta, int qbits, double *__restrict__ out_score, double *__restrict__ out_p,
                         int *__restrict__ out_hits, int *__restrict__ out_n, unsigned long long *__restrict__ recheck_counter) {
    FPT_DYN_SMEM(smem_raw);
    __shared__ __align__(8) uint64_t bar_full[2], bar_empty[2], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t s_tmem;
    __shared__ double s_dmax, s_box[(FPT_UMMA_THREADS / 32) * 4];
    __shared__ int s_flag;
    const int T = FPT_UMMA_THREADS, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int kp = fpt_umma_kp(m), np = fpt_umma_np(m), nkc = kp / FPT_UMMA_KC, ntn = np / FPT_UMMA_NT;
    unsigned char *smem = smem_raw + ((128u - (fpt_smem_u32(smem_raw) & 127u)) & 127u);
    unsigned char *tileA = smem;                                /* 128 x kp membership rows; with the ring: shuffle space */
    unsigned char *ring = tileA + (size_t)FPT_UMMA_BATCH * kp;
    const size_t shuf_bytes = (size_t)FPT_UMMA_BATCH * kp + 2 * FPT_UMMA_STAGE;
    size_t off = shuf_bytes;
    double *X = (double *)(smem + off); off += (size_t)2 * m * 8;
    uint2 *rtab = (uint2 *)(smem + off); off += (size_t)(m + 1) * 8;
    long long *wsum = (long long *)(smem + off); off += (size_t)2 * FPT_UMMA_LRING * 8;
    double *stage = (double *)(smem + off) + 32 * warp; off += (size_t)(FPT_UMMA_THREADS / 32) * 32 * 8;
    int *hit_s = (int *)(smem + off); off += (size_t)FPT_UMMA_BATCH * 4;
    int *scan = (int *)(smem + off);
    unsigned char *gs = gscratch + (size_t)blockIdx.x * gscratch_per_cta;
    unsigned char *qd = gs; gs += (size_t)FPT_UMMA_DIGITS * np * kp;
    unsigned short *labels = (unsigned short *)gs;
    const size_t qd_digit = (size_t)np * kp;

    for (int n = tid; n <= m; n += T) rtab[n] = fpt_umma_rtab_entry(n);
    if (tid == 0) {
        for (int s = 0; s < 2; s++) {
            fpt_mbar_init(&bar_full[s], 1); fpt_mbar_init(&bar_empty[s], 1);
            fpt_mbar_init(&bar_tfull[s], 1); fpt_mbar_init(&bar_tempty[s], 4);
        }
        fpt_mbar_fence_init();
    }
    if (warp == 0) fpt_tmem_alloc(&s_tmem, 512);
    fpt_tc_fence_before();
    __syncthreads();
    fpt_tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = fpt_umma_idesc_u8(FPT_UMMA_BATCH, FPT_UMMA_NT);
    uint32_t ring_it = 0, acc_it = 0;       /* ring stages / accumulators used so far: each role counts its own copy */
    unsigned long long rechecks = 0;
    const int rbytes = fpt_umma_rbytes(m);
    const int rows_fit = min(FPT_UMMA_BATCH, (int)(shuf_bytes / (size_t)rbytes));

    for (long long w = blockIdx.x; w < nwin; w += gridDim.x) {
        if (status[w] != FPT_WIN_SCORED) continue;
        long long t_mark = clock64();
        for (int e = tid; e < 2 * m; e += T) X[e] = Xall[(size_t)w * 2 * m + e];
        __syncthreads();
        /* surrogate scale from the bounding box of the embedding (known before the distances are) */
        double xlo = 1e308, xhi = -1e308, ylo = 1e308, yhi = -1e308;
        for (int e = tid; e < m; e += T) {
            const double x = X[2 * e], y = X[2 * e + 1];
            xlo = fmin(xlo, x); xhi = fmax(xhi, x); ylo = fmin(ylo, y); yhi = fmax(yhi, y);
        }
        for (int o = 16; o > 0; o >>= 1) {
            xlo = fmin(xlo, __shfl_xor_sync(FPT_FULL_MASK, xlo, o)); xhi = fmax(xhi, __shfl_xor_sync(FPT_FULL_MASK, xhi, o));
            ylo = fmin(ylo, __shfl_xor_sync(FPT_FULL_MASK, ylo, o)); yhi = fmax(yhi, __shfl_xor_sync(FPT_FULL_MASK, yhi, o));
        }
        if (lane == 0) { double *b4 = s_box + 4 * warp; b4[0] = xlo; b4[1] = xhi; b4[2] = ylo; b4[3] = yhi; }
        __syncthreads();
        if (tid == 0) {
            for (int k = 1; k < (T >> 5); k++) {
                xlo = fmin(xlo, s_box[4 * k]); xhi = fmax(xhi, s_box[4 * k + 1]); ylo = fmin(ylo, s_box[4 * k + 2]); yhi = fmax(yhi, s_box[4 * k + 3]);
            }
            const double ex = xhi - xlo, ey = yhi - ylo;
            s_dmax = sqrt(ex * ex + ey * ey) * 1.000000000001;
        }
        __syncthreads();
        const double dmax = s_dmax;
        const bool scale_ok = (dmax > 0.0) && (dmax < 1e300);
        const double S = scale_ok ? (double)(1u << qbits) / dmax : 0.0;
        /* one pass: surrogate distances and the four digit matrices of their quantised values in the tile layout. Thread order: 4 bytes of k, then the row n, then 16 bytes of k:
           a warp fills one 128-byte core matrix per digit; no divisions, the four k-side points stay in registers. */
        int bad = 0;
        {
            const int kq = tid & 3, nfirst = tid >> 2;          /* T / 4 rows per sweep; np is a multiple of 256 */
            for (int kg = 0; kg < (kp >> 4); kg++) {
                const int k4 = (kg << 4) + (kq << 2);
                double xk[4], yk[4];
#pragma unroll
                for (int b = 0; b < 4; b++) { const int k = k4 + b < m ? k4 + b : 0; xk[b] = X[2 * k]; yk[b] = X[2 * k + 1]; }
                /* tile (n / 256, k / 128) of digit d at ((d * ntn + nt) * nkc + kc) * 32 KB */
                const size_t kpart = (size_t)(k4 >> 7) * FPT_UMMA_STAGE + (size_t)((k4 & 127) >> 4) * (FPT_UMMA_NT * 16) + (size_t)(k4 & 15);
                for (int n = nfirst; n < np; n += T >> 2) {
                    unsigned w0 = 0, w1 = 0, w2 = 0, w3 = 0;
                    if (n < m && k4 < m) {
                        const double xn = X[2 * n], yn = X[2 * n + 1];
#pragma unroll
                        for (int b = 0; b < 4; b++) {
                            const int k = k4 + b;
                            if (k < m) {
                                const double d = k != n ? fpt_umma_dist(xn, yn, xk[b], yk[b]) : 0.0;
                                if (!(d == d)) bad = 1;
                                const unsigned qv = (scale_ok && d == d) ? (unsigned)__double2ll_rn(d * S) : 0u;
                                w0 |= (qv & 0xffu) << (8 * b); w1 |= ((qv >> 8) & 0xffu) << (8 * b);
                                w2 |= ((qv >> 16) & 0xffu) << (8 * b); w3 |= (qv >> 24) << (8 * b);
                            }
                        }
                    }
                    const size_t toff = (size_t)(n >> 8) * nkc * FPT_UMMA_STAGE + (size_t)((n & 255) >> 3) * 128 + (size_t)(n & 7) * 16 + kpart;
                    *reinterpret_cast<unsigned *>(qd + toff) = w0;
                    *reinterpret_cast<unsigned *>(qd + qd_digit + toff) = w1;
                    *reinterpret_cast<unsigned *>(qd + 2 * qd_digit + toff) = w2;
                    *reinterpret_cast<unsigned *>(qd + 3 * qd_digit + toff) = w3;
                }
            }
        }
        fpt_fence_proxy_async();                                /* the digit matrices are read back by bulk copies */
        bad = __syncthreads_or(bad);
        FPT_UMMA_MARK(0);
        const double score = out_score[w];                      /* fpt_css_observed_kernel ran before this kernel */
        FPT_UMMA_MARK(1);
        const bool use_surrogate = scale_ok && !bad && (score == score) && (fabs(score) < 1e300);
        const double a_ = (double)asize, b_ = (double)bsize;
        const double wterm = (asize > 1 ? 1.0 / (a_ * a_) : 0.0) + (bsize > 1 ? 1.0 / (b_ * b_) : 0.0);
        /* |surrogate - reference score| <= E, see fpt_css_perm_large.cuh; here q is the rounding of a distance that is itself
           within 2^-50 of the reference's, i.e. |q / S - d| <= (0.5 + 2^-19) / S: the factor 1.00001 covers it */
        const double E = use_surrogate ? (0.5 / S) * (1.0 + (a_ + b_) * wterm) * 1.00001 +
                                         8.0 * 1.2e-16 * dmax * (a_ * b_ + 2.0 * (a_ + b_)) : 0.0;
        const double invS = use_surrogate ? 1.0 / S : 0.0;
        const double c_bet = invS / (a_ * b_);
        const double c_wa = asize > 1 ? invS / (a_ * a_ * (a_ - 1.0)) : 0.0;
        const double c_wb = bsize > 1 ? invS / (b_ * b_ * (b_ - 1.0)) : 0.0;
        const int use_a = asize <= bsize;
        const uint64_t st_win = state_override ? state_override[w] : fpt_stream_state(seed, wbase + w, FPT_STREAM_RESAMPLE);
        const int draws = m - 1;
        int hits = 0, ndone = 0, produced = 0;                 /* produced: permutations shuffled so far (ring rows written) */
        bool stopped = false;
        while (!stopped && hits < treshold && ndone < runs) {
            const int nvalid = min(FPT_UMMA_BATCH, runs - ndone);
            /* independent shuffles: permutation k starts k (m-1) draws into the window's stream. Random swaps want shared
               memory: as many label rows at a time as the tile space holds (98 at m = 1000), copied out to a ring of 256 rows
               in global memory. A round is one dependent chain of m - 1 swaps whatever the number of rows, so the rounds are
               always full and run ahead of the batches: 11 rounds per 1000 permutations, not two per batch of 128. */
            while (produced < ndone + nvalid) {
                const int nb = min(rows_fit, runs - produced);
                /* one row per thread on the first warps: consecutive lanes on consecutive rows (an odd word count apart: no
                   bank conflicts on equal indices), full warps so that the chain's instructions are issued once per 32 rows */
                const int myrow = tid;
                if (myrow < nb) {
                    long long sa = 0, sb = 0;
                    fpt_umma_shuffle(reinterpret_cast<unsigned short *>(tileA + (size_t)myrow * rbytes), m, rtab,
                                     fpt_lcg_skip(st_win, (uint64_t)(produced + myrow) * (uint64_t)draws), use_surrogate, X, S, asize, &sa, &sb);
                    const int slot = (produced + myrow) & (FPT_UMMA_LRING - 1);
                    wsum[slot] = sa; wsum[FPT_UMMA_LRING + slot] = sb;
                }
                __syncthreads();
                FPT_UMMA_MARK(2);
                /* labels out to the ring's global rows, one warp per row, a 32-bit word (two labels) per lane and trip */
                for (int rr = warp; rr < nb; rr += T >> 5) {
                    const unsigned short *row = reinterpret_cast<const unsigned short *>(tileA + (size_t)rr * rbytes);
                    unsigned short *orow = labels + (size_t)((produced + rr) & (FPT_UMMA_LRING - 1)) * m;
                    if ((m & 1) == 0) {
                        const uint32_t *rw = reinterpret_cast<const uint32_t *>(row);
                        uint32_t *ow = reinterpret_cast<uint32_t *>(orow);
#pragma unroll 4
                        for (int e = lane; e < (m >> 1); e += 32) ow[e] = rw[e];
                    } else {
                        for (int col = lane; col < m; col += 32) orow[col] = row[col];
                    }
                }
                __syncthreads();
                FPT_UMMA_MARK(6);
                produced += nb;
            }
            FPT_UMMA_MARK(6);
            /* membership rows of the smaller group (the A operand): one sweep over its labels, a byte store each */
            for (int e = tid; e < (FPT_UMMA_BATCH * kp) >> 4; e += T) reinterpret_cast<uint4 *>(tileA)[e] = make_uint4(0u, 0u, 0u, 0u);
            if (tid < FPT_UMMA_BATCH) hit_s[tid] = 0;
            __syncthreads();
            if (use_surrogate) {
                const int lo = use_a ? 0 : asize, cnt = use_a ? asize : bsize;
                const int total = nvalid * cnt;
                for (int e0 = tid; e0 < total; e0 += 8 * T) {
                    int pp[8], cc[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) {
                        const int e = e0 + u * T;
                        pp[u] = e < total ? e / cnt : -1;
                        cc[u] = e < total ? (int)labels[(size_t)((ndone + pp[u]) & (FPT_UMMA_LRING - 1)) * m + lo + (e - pp[u] * cnt)] : 0;
                    }
#pragma unroll
                    for (int u = 0; u < 8; u++)
                        if (pp[u] >= 0) tileA[fpt_umma_tile_off(FPT_UMMA_BATCH, pp[u], cc[u])] = 1;
                }
            }
            fpt_fence_proxy_async();                            /* rows and ring space: generic writes before async reads / writes */
            __syncthreads();
            FPT_UMMA_MARK(3);
            if (use_surrogate) {
                if (warp == 0) {
                    if (lane == 0) {                            /* producer: digit tiles into the ring */
                        for (int nt = 0; nt < ntn; nt++)
                            for (int d = 0; d < FPT_UMMA_DIGITS; d++)
                                for (int kc = 0; kc < nkc; kc++) {
                                    const uint32_t s = ring_it & 1u;
                                    fpt_mbar_wait(&bar_empty[s], ((ring_it >> 1) & 1u) ^ 1u);
                                    fpt_mbar_expect_tx(&bar_full[s], FPT_UMMA_STAGE);
                                    fpt_bulk_g2s(ring + (size_t)s * FPT_UMMA_STAGE,
                                                 qd + (size_t)d * qd_digit + ((size_t)nt * nkc + kc) * FPT_UMMA_STAGE, FPT_UMMA_STAGE, &bar_full[s]);
                                    ring_it++;
                                }
                    }
                } else if (warp == 1) {
                    if (lane == 0) {                            /* MMA issuer */
                        const uint32_t a_base = fpt_smem_u32(tileA), r_base = fpt_smem_u32(ring);
                        for (int t = 0; t < ntn * FPT_UMMA_DIGITS; t++) {
                            const uint32_t as = acc_it & 1u;
                            fpt_mbar_wait(&bar_tempty[as], ((acc_it >> 1) & 1u) ^ 1u);
                            fpt_tc_fence_after();
                            for (int kc = 0; kc < nkc; kc++) {
                                const uint32_t s = ring_it & 1u;
                                fpt_mbar_wait(&bar_full[s], (ring_it >> 1) & 1u);
                                fpt_tc_fence_after();
#pragma unroll
                                for (int ks = 0; ks < FPT_UMMA_KC / 32; ks++) {
                                    const uint64_t ad = fpt_umma_desc(a_base + (uint32_t)(kc * 8 + ks * 2) * (FPT_UMMA_BATCH * 16), FPT_UMMA_BATCH * 16, 128);
                                    const uint64_t bd = fpt_umma_desc(r_base + s * FPT_UMMA_STAGE + (uint32_t)(ks * 2) * (FPT_UMMA_NT * 16), FPT_UMMA_NT * 16, 128);
                                    fpt_umma_u8(tmem + as * FPT_UMMA_NT, ad, bd, idesc, (kc | ks) != 0);
                                }
                                fpt_umma_commit(&bar_empty[s]);
                                ring_it++;
                            }
                            fpt_umma_commit(&bar_tfull[as]);
                            acc_it++;
                        }
                    }
                } else if (warp >= 4 && warp < 8) {             /* drain: lane p of the accumulator = permutation p */
                    const int wq = warp - 4, p = wq * 32 + lane;
                    const unsigned char *mrow = tileA + (size_t)(p >> 3) * 128 + (size_t)(p & 7) * 16;
                    uint32_t acc[FPT_UMMA_DIGITS] = { 0u, 0u, 0u, 0u };
                    for (int nt = 0; nt < ntn; nt++) {
#pragma unroll
                        for (int d = 0; d < FPT_UMMA_DIGITS; d++) {
                            const uint32_t as = acc_it & 1u;
                            fpt_mbar_wait(&bar_tfull[as], (acc_it >> 1) & 1u);
                            fpt_tc_fence_after();
                            uint32_t a = 0u;
                            for (int c = 0; c < FPT_UMMA_NT; c += 32) {
                                uint32_t v[32];
                                fpt_tmem_ld32(tmem + ((uint32_t)(wq * 32) << 16) + as * FPT_UMMA_NT + (uint32_t)c, v);
                                const int n0 = nt * FPT_UMMA_NT + c;
                                uint4 z0 = make_uint4(0u, 0u, 0u, 0u), z1 = z0;
                                if (n0 < kp) {
                                    z0 = *reinterpret_cast<const uint4 *>(mrow + (size_t)(n0 >> 4) * (FPT_UMMA_BATCH * 16));
                                    z1 = *reinterpret_cast<const uint4 *>(mrow + (size_t)((n0 >> 4) + 1) * (FPT_UMMA_BATCH * 16));
                                }
                                fpt_tmem_ld_wait();
                                const uint32_t zz[8] = { z0.x, z0.y, z0.z, z0.w, z1.x, z1.y, z1.z, z1.w };
#pragma unroll
                                for (int i = 0; i < 32; i++) {
                                    const uint32_t zb = (zz[i >> 2] >> (8 * (i & 3))) & 1u;     /* 1: column inside the group */
                                    a += v[i] & (zb - 1u);
                                }
                            }
                            acc[d] += a;
                            fpt_tc_fence_before();
                            __syncwarp();
                            if (lane == 0) fpt_mbar_arrive(&bar_tempty[as]);
                            acc_it++;
                        }
                    }
                    long long bet = 0;
#pragma unroll
                    for (int d = FPT_UMMA_DIGITS; d--;) bet = (bet << 8) + (long long)acc[d];
                    const int pslot = (ndone + p) & (FPT_UMMA_LRING - 1);
                    const double approx = (double)bet * c_bet - (a_ + b_) * ((double)wsum[pslot] * c_wa + (double)wsum[FPT_UMMA_LRING + pslot] * c_wb);
                    const double diff = approx - score;
                    int hit = diff > 0.0;
                    /* within E of the observed score: the whole warp re-scores that permutation in the reference's order */
                    unsigned need = __ballot_sync(FPT_FULL_MASK, p < nvalid && !(fabs(diff) > E));
                    while (need) {
                        const int src = __ffs(need) - 1;
                        need &= need - 1;
                        const double sc = fpt_warp_css_score(X, stage, labels + (size_t)((ndone + wq * 32 + src) & (FPT_UMMA_LRING - 1)) * m, asize, bsize, lane);
                        if (lane == src) { hit = sc >= score ? 1 : 0; rechecks++; }
                    }
                    if (p < nvalid) hit_s[p] = hit;
                }
            } else {                                            /* no surrogate (NaN embedding ...): every permutation exactly */
                for (int p = warp; p < nvalid; p += T >> 5) {
                    const double sc = fpt_warp_css_score(X, stage, labels + (size_t)((ndone + p) & (FPT_UMMA_LRING - 1)) * m, asize, bsize, lane);
                    if (lane == 0) hit_s[p] = sc >= score ? 1 : 0;
                }
            }
            __syncthreads();
            FPT_UMMA_MARK(4);
            const int hit = tid < FPT_UMMA_BATCH ? hit_s[tid] : 0;
            int chunk_hits = 0;
            const int hincl = fpt_block_scan_incl(hit, scan, &chunk_hits);
            if (tid == 0) s_flag = -1;
            __syncthreads();
            if (hit && hits + hincl == treshold) s_flag = tid;  /* the permutation at which the loop exits */
            __syncthreads();
            if (s_flag >= 0) { ndone += s_flag + 1; hits = treshold; stopped = true; }
            else { hits += chunk_hits; ndone += nvalid; }
            __syncthreads();
            FPT_UMMA_MARK(5);
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
    fpt_tc_fence_before();
    __syncthreads();
    if (warp == 0) fpt_tmem_free(tmem, 512);
}

#endif
