/*
 * fpt_css_k4.cuh — the genotype-distance matrix of a large-cohort window as a dense integer contraction.
 *
 * compare_all (css/css.c:277-327, paths relative to /root/reference/statistics/) counts, for every pair of individuals, the
 * SNPs of the window at which one is homozygous major (3) and the other homozygous minor (-3). With the 0/1 indicator rows
 * P_i ("is 3") and M_i ("is -3") over the window's SNPs that count is
 *
 *         D = P M' + M P'  =  [P | M] [M | P]'                      (m x m, K = 2 npos)
 *
 * i.e. ONE u8 x u8 -> s32 GEMM of the genotype indicators with themselves: the "int8 genotype GEMM G G'" of the north star.
 * Two interchangeable kernels produce it (fpt_set_k4_mode; same bits out, the parity tests require it):
 *
 *   fpt_css_k4_umma_kernel   tcgen05.mma kind::i8, 128 x 256 x 32 per instruction, accumulators in tensor memory. One CTA per SM
 *                            walks windows; per window the bit-planes are expanded to 0/1 bytes straight into the shared-memory
 *                            operand tiles (K-major core-matrix layout of fpt_umma.cuh): the B tile (256 individuals x K) stays
 *                            resident while the A tiles (128 x K) alternate between two buffers, so that the expansion of the next
 *                            tile and the drain of the previous accumulator overlap the MMAs in flight. Windows with more than
 *                            192 SNPs run the same loop over K chunks with accumulation.
 *   fpt_css_k4_popc_kernel   bit-plane AND + popcount (exact too, 32 SNPs per instruction, no tensor core), a warp per row.
 *
 * (The diagonal needs no special case: P_i and M_i are disjoint, so D_ii = 0.)
 *
 * Output: COUNT CODES, row-major with the row stride padded to a multiple of 16 (`fpt_k4_ld`): one byte per pair when the window
 * holds <= 255 SNPs (no count can exceed 255), else two. Code 0 = "blank" (fill_averages, css.c:337-366, replaces it by the
 * mean). Nothing else is written: blanks, their mean and the double centring are derived from the codes by the Lanczos kernel
 * (fpt_css_lanczos.cuh), so the m x m fp64 matrix of the reference (8 MB per window at m = 1000) never exists.
 */
#ifndef FPT_CSS_K4_CUH
#define FPT_CSS_K4_CUH

#include "fpt_css.cuh"
#ifndef FPT_EMU
#include "fpt_umma.cuh"
#endif

FPT_HD int fpt_k4_ld(int m) { return (m + 15) & ~15; }
FPT_HD int fpt_k4_code_bytes(int npos) { return npos <= 255 ? 1 : 2; }
/* bytes reserved per window in the code buffer (two-byte codes, 256-byte aligned) */
FPT_HD size_t fpt_k4_window_stride(int m) { return (((size_t)m * fpt_k4_ld(m) * 2) + 255) & ~(size_t)255; }

/* diagnostic: SM cycles per phase of the tcgen05 kernel, summed over CTAs and windows (0 operand expansion, 1 waiting for MMAs,
   2 accumulator drain + code stores), read and reset by fpt_debug_k4_phases() */
#ifndef FPT_EMU
__device__ unsigned long long fpt_k4_phase_cycles[4];
#endif

/* 16 consecutive SNPs [s, s + 16) of individual i's bit-plane `plane`, bit b = SNP s + b; SNPs at or beyond `s_end` read as 0 */
FPT_D unsigned fpt_k4_bits16(const unsigned *__restrict__ planes, int m, int i, int plane, int s, int s_end) {
    if (s >= s_end) return 0u;
    const int wq = s >> 5, sh = s & 31;
    const unsigned lo = planes[((size_t)wq * 2 + plane) * m + i];
    unsigned bits = lo >> sh;
    if (sh > 16 && ((s + 15) >> 5) > wq && (((wq + 1) << 5) < s_end))
        bits |= planes[((size_t)(wq + 1) * 2 + plane) * m + i] << (32 - sh);
    bits &= 0xffffu;
    const int left = s_end - s;
    if (left < 16) bits &= (1u << left) - 1u;
    return bits;
}

/* ============================================================================================ popcount form */
/* One CTA per window. The window's bit-plane words are staged in shared memory eight at a time (masked to the window once);
   a warp owns row i with its 16 words in registers, lanes run across j: two conflict-free shared loads, two ANDs and two
   POPCs per word and pair. Windows longer than 256 SNPs take several passes and accumulate through the code buffer. */
#define FPT_K4_POPC_WORDS 8
FPT_HD size_t fpt_k4_popc_smem(int m) { return (size_t)FPT_K4_POPC_WORDS * 2 * m * sizeof(unsigned); }

__global__ void __launch_bounds__(256)
fpt_css_k4_popc_kernel(const unsigned *__restrict__ planes, int m, const int *__restrict__ wleft, const int *__restrict__ wright,
                       long long nwin, unsigned char *__restrict__ codes, size_t stride) {
    FPT_DYN_SMEM(smem);
    unsigned *sw = reinterpret_cast<unsigned *>(smem);          /* [word][plane][individual] */
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int ld = fpt_k4_ld(m);
    for (long long w = blockIdx.x; w < nwin; w += gridDim.x) {
        const int l = wleft[w], r = wright[w];
        if (r <= l) continue;
        const int w0 = l >> 5, w1 = (r - 1) >> 5;
        const int esz = fpt_k4_code_bytes(r - l);
        unsigned char *c8 = codes + (size_t)w * stride;
        unsigned short *c16 = reinterpret_cast<unsigned short *>(c8);
        for (int wb = w0; wb <= w1; wb += FPT_K4_POPC_WORDS) {
            const int nw = min(FPT_K4_POPC_WORDS, w1 - wb + 1);
            __syncthreads();
            for (int e = threadIdx.x; e < nw * 2 * m; e += blockDim.x) {
                const int wc = wb + e / (2 * m);
                unsigned mask = 0xffffffffu;
                if (wc == w0) mask &= 0xffffffffu << (l & 31);
                if (wc == w1) mask &= 0xffffffffu >> (31 - ((r - 1) & 31));
                sw[e] = planes[(size_t)wb * 2 * m + e] & mask;
            }
            __syncthreads();
            for (int i = warp; i < m; i += nwarp) {
                unsigned pi[FPT_K4_POPC_WORDS], mi[FPT_K4_POPC_WORDS];
#pragma unroll
                for (int k = 0; k < FPT_K4_POPC_WORDS; k++) { pi[k] = k < nw ? sw[(size_t)k * 2 * m + i] : 0u; mi[k] = k < nw ? sw[(size_t)k * 2 * m + m + i] : 0u; }
                for (int j0 = 0; j0 < ld; j0 += 32) {
                    const int j = j0 + lane;
                    if (j >= ld) continue;
                    int cnt = 0;
                    if (j < m) {
#pragma unroll
                        for (int k = 0; k < FPT_K4_POPC_WORDS; k++)
                            if (k < nw) cnt += __popc(pi[k] & sw[(size_t)k * 2 * m + m + j]) + __popc(mi[k] & sw[(size_t)k * 2 * m + j]);
                    }
                    if (esz == 1) c8[(size_t)i * ld + j] = (unsigned char)(cnt + (wb > w0 ? (int)c8[(size_t)i * ld + j] : 0));
                    else c16[(size_t)i * ld + j] = (unsigned short)(cnt + (wb > w0 ? (int)c16[(size_t)i * ld + j] : 0));
                }
            }
        }
    }
}

#ifndef FPT_EMU
/* ============================================================================================ tcgen05 form */
#define FPT_K4_THREADS 256
#define FPT_K4_KH_MAX 192                      /* SNPs per K chunk: K = 2 x 192 = 384 bytes per operand row */
#define FPT_K4_KMAX (2 * FPT_K4_KH_MAX)
#define FPT_K4_SMEM ((size_t)(256 + 2 * 128) * FPT_K4_KMAX + 1024)

/* rows [row0, row0 + R) of the operand ([P | M] when swap = 0, [M | P] when swap = 1) for SNPs [s0, s0 + nk) of the window, as
   0/1 bytes in the canonical K-major tile of R rows; kh = nk rounded up to 16, K = 2 kh. Consecutive threads take consecutive
   rows of one 16-byte k-group: coalesced plane loads, conflict-free 16-byte shared-memory stores. */
template <int R>
FPT_D void fpt_k4_expand(unsigned char *tile, const unsigned *__restrict__ planes, int m, int row0, int swap, int s0, int nk, int kh) {
    const int ngroups = (2 * kh) >> 4;
    for (int item = threadIdx.x; item < ngroups * R; item += blockDim.x) {
        const int kg = item / R, rr = item - kg * R;
        const int i = row0 + rr;
        const int half = (kg << 4) >= kh;
        const int s = s0 + (kg << 4) - (half ? kh : 0);
        unsigned bits = 0u;
        if (i < m) bits = fpt_k4_bits16(planes, m, i, half ^ swap, s, s0 + nk);
        uint4 v;
        v.x = ((bits & 0xfu) * 0x00204081u) & 0x01010101u;
        v.y = (((bits >> 4) & 0xfu) * 0x00204081u) & 0x01010101u;
        v.z = (((bits >> 8) & 0xfu) * 0x00204081u) & 0x01010101u;
        v.w = (((bits >> 12) & 0xfu) * 0x00204081u) & 0x01010101u;
        *reinterpret_cast<uint4 *>(tile + (size_t)kg * ((size_t)R * 16) + (size_t)(rr >> 3) * 128 + (size_t)(rr & 7) * 16) = v;
    }
}

__global__ void __launch_bounds__(FPT_K4_THREADS, 1)
fpt_css_k4_umma_kernel(const unsigned *__restrict__ planes, int m, const int *__restrict__ wleft, const int *__restrict__ wright,
                       long long nwin, unsigned char *__restrict__ codes, size_t stride) {
    FPT_DYN_SMEM(smem_raw);
    __shared__ __align__(8) uint64_t bar_buf[2], bar_tile[2];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char *smem = smem_raw + ((1024u - (fpt_smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char *sB = smem;                                   /* 256 x K */
    unsigned char *sA[2] = { smem + (size_t)256 * FPT_K4_KMAX, smem + (size_t)(256 + 128) * FPT_K4_KMAX };
    if (tid == 0) {
        fpt_mbar_init(&bar_buf[0], 1); fpt_mbar_init(&bar_buf[1], 1);
        fpt_mbar_init(&bar_tile[0], 1); fpt_mbar_init(&bar_tile[1], 1);
        fpt_mbar_fence_init();
    }
    if (warp == 0) fpt_tmem_alloc(&s_tmem, 512);
    fpt_tc_fence_before();
    __syncthreads();
    fpt_tc_fence_after();
    const uint32_t tmem = s_tmem;
    const uint32_t idesc = fpt_umma_idesc_u8(128, 256);
    const int ld = fpt_k4_ld(m);
    const int NT = (m + 255) >> 8, MT = (m + 127) >> 7;
    /* commits made so far on each barrier (every thread keeps the same counts: the loop structure is uniform) */
    uint32_t used_buf[2] = { 0u, 0u }, used_tile[2] = { 0u, 0u };
    uint32_t it = 0, tile_no = 0;
    long long t_mark = clock64();
#define FPT_K4_MARK(slot) do { if (tid == 0) { const long long now_ = clock64(); atomicAdd(&fpt_k4_phase_cycles[slot], (unsigned long long)(now_ - t_mark)); t_mark = now_; } } while (0)

    for (long long w = blockIdx.x; w < nwin; w += gridDim.x) {
        const int l = wleft[w], r = wright[w];
        if (r <= l) continue;
        const int npos = r - l, esz = fpt_k4_code_bytes(npos);
        const int nchunks = (npos + FPT_K4_KH_MAX - 1) / FPT_K4_KH_MAX;
        unsigned char *cw = codes + (size_t)w * stride;
        int pend_n = -1, pend_mi = 0;                           /* tile whose accumulator waits to be drained */
        uint32_t pend_tile = 0;
        int cached_n = -1, cached_c = -1;
        if (tid == 0) t_mark = clock64();
        for (int n = 0; n <= NT; n++) {
            for (int mi = 0; mi < MT; mi++) {
                const bool real = n < NT;
                if (!real && mi > 0) break;                     /* one extra round: drain of the last tile only */
                if (real) {
                    const uint32_t acc = tile_no & 1u;
                    for (int c = 0; c < nchunks; c++) {
                        const int s0 = l + c * FPT_K4_KH_MAX, nk = min(FPT_K4_KH_MAX, r - s0), kh = (nk + 15) & ~15;
                        const uint32_t b = it & 1u;
                        /* the MMAs that read this A buffer two iterations ago (and, for B, every MMA issued so far) are done */
                        if (used_buf[b]) fpt_mbar_wait(&bar_buf[b], (used_buf[b] - 1u) & 1u);
                        if (cached_n != n || cached_c != c) {
                            if (used_buf[b ^ 1u]) fpt_mbar_wait(&bar_buf[b ^ 1u], (used_buf[b ^ 1u] - 1u) & 1u);
                            fpt_k4_expand<256>(sB, planes, m, n << 8, 1, s0, nk, kh);
                            cached_n = n; cached_c = c;
                        }
                        fpt_k4_expand<128>(sA[b], planes, m, mi << 7, 0, s0, nk, kh);
                        fpt_fence_proxy_async();                /* generic-proxy stores -> visible to the tensor core's async reads */
                        fpt_tc_fence_before();
                        __syncthreads();
                        FPT_K4_MARK(0);
                        if (tid == 0) {
                            fpt_tc_fence_after();
                            const uint32_t a_base = fpt_smem_u32(sA[b]), b_base = fpt_smem_u32(sB);
                            for (int ks = 0; ks < (2 * kh) >> 5; ks++) {
                                const uint64_t ad = fpt_umma_desc(a_base + (uint32_t)ks * 2u * 128u * 16u, 128 * 16, 128);
                                const uint64_t bd = fpt_umma_desc(b_base + (uint32_t)ks * 2u * 256u * 16u, 256 * 16, 128);
                                fpt_umma_u8(tmem + acc * 256u, ad, bd, idesc, (c | ks) != 0);
                            }
                            fpt_umma_commit(&bar_buf[b]);
                            if (c == nchunks - 1) fpt_umma_commit(&bar_tile[acc]);
                        }
                        used_buf[b]++;
                        if (c == nchunks - 1) used_tile[acc]++;
                        it++;
                    }
                }
                /* drain the previous tile while this one's MMAs run */
                if (pend_n >= 0) {
                    const uint32_t pacc = pend_tile & 1u;
                    fpt_mbar_wait(&bar_tile[pacc], (used_tile[pacc] - 1u) & 1u);     /* consecutive tiles alternate accumulators: not committed again yet */
                    fpt_tc_fence_after();
                    FPT_K4_MARK(1);
                    const int q = warp & 3, hcol = warp >> 2;
                    const int i = (pend_mi << 7) + 32 * q + lane;
                    for (int cc = 0; cc < 128; cc += 32) {
                        uint32_t v[32];
                        fpt_tmem_ld32(tmem + ((uint32_t)(32 * q) << 16) + pacc * 256u + (uint32_t)(128 * hcol + cc), v);
                        fpt_tmem_ld_wait();
                        const int j0 = (pend_n << 8) + 128 * hcol + cc;
                        if (i < m) {
                            if (esz == 1) {
                                unsigned pk[8];
#pragma unroll
                                for (int k = 0; k < 8; k++) pk[k] = (v[4 * k] & 0xffu) | ((v[4 * k + 1] & 0xffu) << 8) | ((v[4 * k + 2] & 0xffu) << 16) | (v[4 * k + 3] << 24);
                                unsigned char *dst = cw + (size_t)i * ld + j0;
                                if (j0 < ld) *reinterpret_cast<uint4 *>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                                if (j0 + 16 < ld) *reinterpret_cast<uint4 *>(dst + 16) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                            } else {
                                unsigned pk[16];
#pragma unroll
                                for (int k = 0; k < 16; k++) pk[k] = (v[2 * k] & 0xffffu) | (v[2 * k + 1] << 16);
                                unsigned short *dst = reinterpret_cast<unsigned short *>(cw) + (size_t)i * ld + j0;
#pragma unroll
                                for (int g = 0; g < 4; g++)
                                    if (j0 + 8 * g < ld) *reinterpret_cast<uint4 *>(dst + 8 * g) = make_uint4(pk[4 * g], pk[4 * g + 1], pk[4 * g + 2], pk[4 * g + 3]);
                            }
                        }
                    }
                    fpt_tc_fence_before();                      /* the accumulator may be overwritten after the next barrier */
                    FPT_K4_MARK(2);
                }
                if (real) { pend_n = n; pend_mi = mi; pend_tile = tile_no; tile_no++; }
                else pend_n = -1;
            }
        }
        __syncthreads();
    }
#undef FPT_K4_MARK
    fpt_tc_fence_before();
    __syncthreads();
    if (warp == 0) fpt_tmem_free(tmem, 512);
}
#endif /* FPT_EMU */

#endif
