/*
 * cuda_emu.h — a tiny CPU stand-in for the CUDA execution model, TEST INFRASTRUCTURE ONLY.
 *
 * The build container has nvcc but no GPU. To debug kernel LOGIC here (index arithmetic, barriers,
 * scans, fixed-point loops) the product's kernel headers (csrc/fpt_*.cuh) are also compiled with g++
 * against this shim: one CTA at a time, one pthread per CUDA thread, __syncthreads() = a pthread
 * barrier, warp collectives = slot exchange + a per-warp barrier. It is slow and only used by
 * tests/emu/emu_driver.cpp at tiny sizes. Nothing under the product package includes this file;
 * the shipped library is the nvcc build of the same headers.
 */
#ifndef FPT_CUDA_EMU_H
#define FPT_CUDA_EMU_H

#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <algorithm>
#include <vector>

#define FPT_EMU 1

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __restrict__
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

namespace emu {
struct Cta {
    pthread_barrier_t bar;
    pthread_barrier_t *warp_bar;
    uint64_t (*slots)[32];
    uint32_t (*wide)[32][8];     /* one 8-register record per lane: warp-wide publish for emulated MMA */
    unsigned nthreads;
};
extern Cta *g_cta;
extern dim3 g_blockDim, g_gridDim;
extern unsigned char *g_dyn_smem;
extern thread_local dim3 t_threadIdx, t_blockIdx;
}  // namespace emu

#define threadIdx (emu::t_threadIdx)
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::g_blockDim)
#define gridDim (emu::g_gridDim)
#define warpSize 32

static inline void __syncthreads() { pthread_barrier_wait(&emu::g_cta->bar); }
static inline void __syncwarp(unsigned = 0xffffffffu) { pthread_barrier_wait(&emu::g_cta->warp_bar[emu::t_threadIdx.x / 32]); }
static inline void __threadfence() { __sync_synchronize(); }
static inline void __threadfence_block() { __sync_synchronize(); }

template <typename T>
static inline T emu_exchange(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "emu exchange");
    unsigned w = emu::t_threadIdx.x / 32, l = emu::t_threadIdx.x % 32;
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    emu::g_cta->slots[w][l] = raw;
    pthread_barrier_wait(&emu::g_cta->warp_bar[w]);
    uint64_t got = emu::g_cta->slots[w][src_lane & 31];
    pthread_barrier_wait(&emu::g_cta->warp_bar[w]);
    T out;
    memcpy(&out, &got, sizeof(T));
    return out;
}
template <typename T> static inline T __shfl_sync(unsigned, T v, int lane) { return emu_exchange(v, lane); }
template <typename T> static inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_exchange(v, (int)(emu::t_threadIdx.x % 32) ^ m); }
template <typename T> static inline T __shfl_down_sync(unsigned, T v, int d) {
    int l = emu::t_threadIdx.x % 32; return emu_exchange(v, l + d < 32 ? l + d : l);
}
template <typename T> static inline T __shfl_up_sync(unsigned, T v, int d) {
    int l = emu::t_threadIdx.x % 32; return emu_exchange(v, l - d >= 0 ? l - d : l);
}
/* every lane publishes n <= 8 registers; returns a pointer to the warp's [32][8] table (valid until emu_warp_release) */
static inline const uint32_t (*emu_warp_publish(const uint32_t *regs, int n))[8] {
    unsigned w = emu::t_threadIdx.x / 32, l = emu::t_threadIdx.x % 32;
    for (int i = 0; i < n; i++) emu::g_cta->wide[w][l][i] = regs[i];
    pthread_barrier_wait(&emu::g_cta->warp_bar[w]);
    return emu::g_cta->wide[w];
}
static inline void emu_warp_release() { pthread_barrier_wait(&emu::g_cta->warp_bar[emu::t_threadIdx.x / 32]); }

static inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned w = emu::t_threadIdx.x / 32, l = emu::t_threadIdx.x % 32;
    emu::g_cta->slots[w][l] = pred ? 1 : 0;
    pthread_barrier_wait(&emu::g_cta->warp_bar[w]);
    unsigned r = 0;
    for (int i = 0; i < 32; i++) r |= (unsigned)(emu::g_cta->slots[w][i] & 1) << i;
    pthread_barrier_wait(&emu::g_cta->warp_bar[w]);
    return r;
}
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
static inline int __all_sync(unsigned m, int p) { return __ballot_sync(m, p) == 0xffffffffu; }
static inline int __syncthreads_or(int p) {
    static int flag;
    __syncthreads();
    if (emu::t_threadIdx.x == 0) flag = 0;
    __syncthreads();
    if (p) __atomic_store_n(&flag, 1, __ATOMIC_RELAXED);
    __syncthreads();
    int r = flag;
    __syncthreads();
    return r;
}

static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline double rsqrt(double x) { return 1.0 / sqrt(x); }
static inline int __syncthreads_and(int p) { return !__syncthreads_or(!p); }
static inline double __hiloint2double(int hi, int lo) {
    unsigned long long bits = ((unsigned long long)(unsigned)hi << 32) | (unsigned long long)(unsigned)lo;
    double d;
    memcpy(&d, &bits, 8);
    return d;
}

/* arithmetic intrinsics (the emu build uses -ffp-contract=off, so plain operators round once) */
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return sqrt(a); }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline double __ull2double_rn(unsigned long long v) { return (double)v; }
static inline double __int2double_rn(int v) { return (double)v; }
static inline long long __double2ll_rn(double v) { return llrint(v); }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
    return (unsigned long long)(((unsigned __int128)a * b) >> 64);
}
static inline long long __double_as_longlong(double d) { long long r; memcpy(&r, &d, 8); return r; }
static inline int __double2hiint(double d) { long long r; memcpy(&r, &d, 8); return (int)(r >> 32); }
static inline double __longlong_as_double(long long v) { double r; memcpy(&r, &v, 8); return r; }
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline int atomicMax(int *p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
static inline int atomicMin(int *p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}
static inline int atomicOr(int *p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicOr(unsigned *p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }

using std::max;
using std::min;
static inline long long max(long long a, int b) { return a > b ? a : (long long)b; }
static inline long long min(long long a, int b) { return a < b ? a : (long long)b; }

struct int4 { int x, y, z, w; };
struct int2 { int x, y; };
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
struct ulonglong2 { unsigned long long x, y; };
static inline ulonglong2 make_ulonglong2(unsigned long long a, unsigned long long b) { ulonglong2 r = { a, b }; return r; }
struct double2 { double x, y; };
static inline int4 make_int4(int a, int b, int c, int d) { int4 r = { a, b, c, d }; return r; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { uint4 r = { a, b, c, d }; return r; }
static inline uint2 make_uint2(unsigned a, unsigned b) { uint2 r = { a, b }; return r; }
static inline int2 make_int2(int a, int b) { int2 r = { a, b }; return r; }
static inline double2 make_double2(double a, double b) { double2 r = { a, b }; return r; }

namespace emu {

struct LaunchCtx {
    void (*body)(void *);
    void *arg;
    unsigned nblocks, nthreads;
    pthread_barrier_t gate;
};

struct ThreadArg { LaunchCtx *ctx; unsigned tid; };

inline void *thread_main(void *p) {
    ThreadArg *ta = (ThreadArg *)p;
    LaunchCtx *c = ta->ctx;
    for (unsigned b = 0; b < c->nblocks; b++) {
        t_threadIdx = dim3(ta->tid, 0, 0);
        t_blockIdx = dim3(b, 0, 0);
        pthread_barrier_wait(&c->gate);
        c->body(c->arg);
        pthread_barrier_wait(&c->gate);
    }
    return 0;
}

/* run `body` as a grid of nblocks CTAs of nthreads threads with dyn_smem bytes of dynamic shared memory */
inline void launch(unsigned nblocks, unsigned nthreads, size_t dyn_smem, void (*body)(void *), void *arg) {
    if (nblocks == 0) return;
    Cta cta;
    cta.nthreads = nthreads;
    unsigned nwarps = (nthreads + 31) / 32;
    pthread_barrier_init(&cta.bar, 0, nthreads);
    cta.warp_bar = (pthread_barrier_t *)malloc(sizeof(pthread_barrier_t) * nwarps);
    for (unsigned w = 0; w < nwarps; w++) {
        unsigned cnt = std::min(32u, nthreads - w * 32);
        pthread_barrier_init(&cta.warp_bar[w], 0, cnt);
    }
    cta.slots = (uint64_t(*)[32])calloc(nwarps, sizeof(uint64_t[32]));
    cta.wide = (uint32_t(*)[32][8])calloc(nwarps, sizeof(uint32_t[32][8]));
    g_cta = &cta;
    g_blockDim = dim3(nthreads, 1, 1);
    g_gridDim = dim3(nblocks, 1, 1);
    g_dyn_smem = (unsigned char *)aligned_alloc(128, ((dyn_smem + 127) / 128 + 1) * 128);
    memset(g_dyn_smem, 0xA5, dyn_smem);   /* shared memory starts as garbage, like on the device */
    LaunchCtx ctx;
    ctx.body = body; ctx.arg = arg; ctx.nblocks = nblocks; ctx.nthreads = nthreads;
    pthread_barrier_init(&ctx.gate, 0, nthreads);
    std::vector<pthread_t> th(nthreads);
    std::vector<ThreadArg> ta(nthreads);
    pthread_attr_t at;
    pthread_attr_init(&at);
    pthread_attr_setstacksize(&at, 512 * 1024);
    for (unsigned t = 0; t < nthreads; t++) {
        ta[t].ctx = &ctx; ta[t].tid = t;
        pthread_create(&th[t], &at, thread_main, &ta[t]);
    }
    for (unsigned t = 0; t < nthreads; t++) pthread_join(th[t], 0);
    pthread_attr_destroy(&at);
    pthread_barrier_destroy(&ctx.gate);
    pthread_barrier_destroy(&cta.bar);
    for (unsigned w = 0; w < nwarps; w++) pthread_barrier_destroy(&cta.warp_bar[w]);
    free(cta.warp_bar); free(cta.slots); free(cta.wide); free(g_dyn_smem);
    g_cta = 0; g_dyn_smem = 0;
}

}  // namespace emu

#define FPT_EMU_DEFINE_GLOBALS                                             \
    namespace emu {                                                        \
    Cta *g_cta = 0; dim3 g_blockDim, g_gridDim; unsigned char *g_dyn_smem = 0; \
    thread_local dim3 t_threadIdx, t_blockIdx;                             \
    }

#endif
