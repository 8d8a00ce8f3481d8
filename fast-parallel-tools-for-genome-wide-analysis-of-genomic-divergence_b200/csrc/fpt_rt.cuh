/*
 * fpt_rt.cuh — small device-side runtime shared by the FET and CSS kernels:
 * the reference's 48-bit LCG with O(log n) skip-ahead, window-keyed stream states,
 * and CTA-wide reductions / scans.
 *
 * Compiled by nvcc for sm_100a (the product). The same header also compiles under g++ against
 * tests/emu/cuda_emu.h (FPT_EMU), which the CPU-only test-suite uses to exercise kernel logic.
 */
#ifndef FPT_RT_CUH
#define FPT_RT_CUH

#ifndef FPT_EMU
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>
#define FPT_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#else
#define FPT_DYN_SMEM(name) unsigned char *name = emu::g_dyn_smem
#endif

#define FPT_HD __host__ __device__ __forceinline__
#define FPT_D __device__ __forceinline__

/* --------------------------------------------------------------------------------------------
 * Random streams. The reference draws bootstrap indices and label shuffles with glibc nrand48 and
 * SMACOF starts with drand48 (fisher/cFisher.c:547-554, css/css.c:675-690, css/css.c:863-864):
 * X <- (0x5DEECE66D X + 0xB) mod 2^48, nrand48 = X >> 17, drand48 = X / 2^48.
 * We keep that generator (so a window reproduces the reference's per-window functions bit for bit
 * from the same state) but key the state per (seed, window, stream) instead of time(NULL), and use
 * the LCG's affine skip-ahead so that every (window, replicate) is addressable by counter.
 */
#define FPT_LCG_A 0x5DEECE66DULL
#define FPT_LCG_C 0xBULL
#define FPT_MASK48 0xFFFFFFFFFFFFULL
#define FPT_STREAM_RESAMPLE 0
#define FPT_STREAM_INIT 1

FPT_HD uint64_t fpt_stream_state(uint64_t seed, long long window, int stream) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ULL * (uint64_t)(4ULL * (uint64_t)window + (uint64_t)stream + 1ULL);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    z ^= z >> 31;
    return z & FPT_MASK48;
}

FPT_HD uint64_t fpt_lcg_next(uint64_t &s) {
    s = (s * FPT_LCG_A + FPT_LCG_C) & FPT_MASK48;
    return s;
}

/* state after n steps: x -> a^n x + c (a^n - 1)/(a - 1), by square-and-multiply on the affine map */
FPT_HD uint64_t fpt_lcg_skip(uint64_t s, uint64_t n) {
    uint64_t ra = 1, rc = 0, ba = FPT_LCG_A, bc = FPT_LCG_C;
    while (n) {
        if (n & 1) { rc = (ba * rc + bc) & FPT_MASK48; ra = (ba * ra) & FPT_MASK48; }
        bc = (ba * bc + bc) & FPT_MASK48;
        ba = (ba * ba) & FPT_MASK48;
        n >>= 1;
    }
    return (ra * s + rc) & FPT_MASK48;
}

/* random_int_nrand48: rejection above RAND_MAX - (RAND_MAX+1) % n, then modulo; counts the draws used */
FPT_HD uint32_t fpt_randint(uint32_t n, uint64_t &s, int &used) {
    uint32_t limit = 2147483647u - (2147483648u % n);
    uint32_t r = (uint32_t)(fpt_lcg_next(s) >> 17);
    used++;
    while (r > limit) { r = (uint32_t)(fpt_lcg_next(s) >> 17); used++; }
    return r % n;
}

/* the same draw with the two per-n constants precomputed: limit = RAND_MAX - (RAND_MAX+1) % n and
   magic = floor(2^32 / n), which turns r % n (r < 2^31) into one multiply-high and a conditional subtract:
   floor(r * magic / 2^32) is floor(r / n) or one less, because r * (2^32 mod n) / (n 2^32) < 1/2 */
FPT_HD uint32_t fpt_randint_limit(uint32_t n) { return 2147483647u - (2147483648u % n); }
FPT_HD uint32_t fpt_randint_magic(uint32_t n) {            /* floor(2^32 / n) in 32-bit arithmetic */
    if (n <= 1) return 0u;
    const uint32_t q = 0xFFFFFFFFu / n, r = 0xFFFFFFFFu - q * n;
    return r == n - 1 ? q + 1 : q;
}
FPT_D uint32_t fpt_randint_fast(uint32_t n, uint32_t limit, uint32_t magic, uint64_t &s, int &used) {
    uint32_t r = (uint32_t)(fpt_lcg_next(s) >> 17);
    used++;
    while (r > limit) { r = (uint32_t)(fpt_lcg_next(s) >> 17); used++; }
    uint32_t rem = r - __umulhi(r, magic) * n;
    if (rem >= n) rem -= n;
    return n > 1 ? rem : 0u;
}

FPT_HD double fpt_drand48(uint64_t &s) {
    /* X / 2^48: X < 2^48 converts exactly, the power-of-two scale is exact */
    return (double)(long long)fpt_lcg_next(s) * 3.5527136788005009e-15;   /* 2^-48 */
}

/* -------------------------------------------------------------------------------------------- */
#define FPT_FULL_MASK 0xffffffffu

/* CTA-wide sum in a fixed order (lane tree, then warp totals in warp order); all threads get it.
   `scratch` must hold 33 doubles of shared memory. blockDim.x must be a multiple of 32. */
FPT_D double fpt_block_sum(double v, double *scratch) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FPT_FULL_MASK, v, o);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    __syncthreads();                         /* scratch may still be read from a previous call */
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0;
        for (int w = 0; w < nwarp; w++) t += scratch[w];
        scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

FPT_D long long fpt_block_sum_i64(long long v, long long *scratch) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FPT_FULL_MASK, v, o);
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long t = 0;
        for (int w = 0; w < nwarp; w++) t += scratch[w];
        scratch[32] = t;
    }
    __syncthreads();
    return scratch[32];
}

/* CTA-wide inclusive prefix sum of one int per thread; `scratch` holds 33 ints; *total = grand sum */
FPT_D int fpt_block_scan_incl(int v, int *scratch, int *total) {
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    int x = v;
    for (int o = 1; o < 32; o <<= 1) {
        int y = __shfl_up_sync(FPT_FULL_MASK, x, o);
        if (lane >= o) x += y;
    }
    __syncthreads();
    if (lane == 31) scratch[warp] = x;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int w = 0; w < nwarp; w++) { int t = scratch[w]; scratch[w] = run; run += t; }
        scratch[32] = run;
    }
    __syncthreads();
    x += scratch[warp];
    if (total) *total = scratch[32];
    return x;
}

#endif
