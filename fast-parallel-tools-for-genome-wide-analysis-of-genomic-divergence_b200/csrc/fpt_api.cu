/*
 * fpt_api.cu — host side of libfpt_b200.so: the C ABI declared in include/fpt_b200.h.
 *
 * Replaces, for the two hot paths only, the reference's drivers
 *   fisher/threadfisher.c:47-251 (threadcompute/mycompute), fisher/cFisher.c:38-115 (compute),
 *   css/threadcss.c:52-293,       css/css.c:49-156
 * (paths relative to /root/reference/statistics/): instead of 64 pthreads pulling tasks of 100 windows
 * off a mutex-protected counter, the whole chromosome goes to the GPU once and every SNP / window is a
 * thread / CTA of a handful of kernel launches.
 *
 * No CPU fallback: every compute entry point needs a CUDA device.
 */
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <chrono>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/fpt_b200.h"
#include "fpt_css.cuh"
#include "fpt_css_eig.cuh"
#include "fpt_css_eig_reg.cuh"
#include "fpt_css_lanczos.cuh"
#include "fpt_css_k4.cuh"
#include "fpt_css_perm.cuh"
#include "fpt_css_perm3.cuh"
#include "fpt_css_perm_large.cuh"
#include "fpt_css_perm_umma.cuh"
#include "fpt_fet.cuh"
#include "fpt_rt.cuh"
#include "fpt_tables.h"

/* ================================================================================================ state */
static thread_local char g_err[512] = "";
/* Process-wide settings. They are atomics, and every entry point takes ONE snapshot of them when it starts (struct Knobs):
   a setter racing a scan in another thread affects the next call, never the kernel route of a call in flight. */
static std::atomic<uint64_t> g_seed{20261018ULL};
static std::atomic<int> g_device{-1};           /* -1: whatever device is current */
static std::atomic<int> g_lanczos_form{3};      /* large-cohort MDS: highest product form allowed (3 8-bit codes with arithmetic squares + blank list, 2 8-bit codes through the table, 1 16-bit codes, 0 fp64 matrix) */
static std::atomic<int> g_lanczos_threads{256};  /* large-cohort MDS on count codes: threads per CTA (512 or 384) */
static std::atomic<int> g_perm_umma{1};         /* large cohorts: 1 = tcgen05 permutation kernel, 0 = the general (mma.sync) kernel */
static std::atomic<int> g_perm_chain{0};        /* CSS label shuffles: 0 = independent per permutation, 1 = the reference's chain */
static std::atomic<int> g_perm_small{1};        /* cohorts of 8..64, independent shuffles: 1 = fpt_css_perm3_kernel, 0 = the round-1 kernel (fpt_css_perm2_kernel) */
static std::atomic<int> g_mds_small{1};         /* cohorts of 3..48, classical MDS: 1 = tridiagonalisation in registers (fpt_css_tridiag_reg_kernel), 0 = the shared-memory kernel */
static std::atomic<int> g_k4_mode{2};           /* large cohorts, genotype-distance matrix: 2 = tcgen05 u8 GEMM, 1 = popcounts, 0 = legacy fp64 matrix */

struct Knobs { int lanczos_form, perm_umma, perm_chain, k4_mode, perm_small, mds_small, lanczos_threads; };
static Knobs knobs_now() {
    Knobs k;
    k.lanczos_form = g_lanczos_form.load(); k.perm_umma = g_perm_umma.load(); k.perm_chain = g_perm_chain.load(); k.k4_mode = g_k4_mode.load(); k.perm_small = g_perm_small.load(); k.mds_small = g_mds_small.load(); k.lanczos_threads = g_lanczos_threads.load();
    return k;
}

static int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CU(call)                                                                                     \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess)                                                                       \
            return fail(FPT_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
    } while (0)
#define CHECK(call)                     \
    do {                                \
        int rc_ = (call);               \
        if (rc_ != FPT_OK) return rc_;  \
    } while (0)

struct DeviceCtx {
    int device = -1;
    int sms = 0;
    int smem_optin = 0;
    unsigned long long *binom = nullptr;
    double *lf = nullptr;
    int lf_maxn = -1;
    std::vector<double *> lf_retired;
    unsigned long long *rechecks = nullptr;
    bool pool_ready = false;
    void *pinned[8] = { nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr };   /* grow-only host staging */
    size_t pinned_bytes[8] = { 0, 0, 0, 0, 0, 0, 0, 0 };
};
static DeviceCtx g_ctx[64];
static std::mutex g_mu;
static std::mutex g_host_mu[64];         /* host entry points that use a device's staging buffers run one at a time */

static int get_ctx(DeviceCtx **out) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(FPT_ERR_NO_DEVICE, "no CUDA device available (%s); libfpt_b200 has no CPU path",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    int dev = g_device.load();
    if (dev >= 0) CU(cudaSetDevice(dev));
    else CU(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(FPT_ERR_ARG, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lk(g_mu);
    DeviceCtx &c = g_ctx[dev];
    if (c.device < 0) {
        cudaDeviceProp pr;
        CU(cudaGetDeviceProperties(&pr, dev));
        c.sms = pr.multiProcessorCount;
        c.smem_optin = (int)pr.sharedMemPerBlockOptin;
        std::vector<unsigned long long> b = fpt_build_binom_table();
        CU(cudaMalloc(&c.binom, b.size() * sizeof(unsigned long long)));
        CU(cudaMemcpy(c.binom, b.data(), b.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
        CU(cudaMalloc(&c.rechecks, sizeof(unsigned long long)));
        CU(cudaMemset(c.rechecks, 0, sizeof(unsigned long long)));
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ULL;                 /* keep freed blocks cached between calls */
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            c.pool_ready = true;
        }
        cudaGetLastError();
        c.device = dev;
    }
    *out = &c;
    return FPT_OK;
}

/* log-factorial table, grown on demand (rare: once per process for a given coverage) */
static int ensure_lf(DeviceCtx *c, int maxn, const double **table) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (maxn <= c->lf_maxn) { *table = c->lf; return FPT_OK; }
    int want = std::max(maxn, 1024);
    std::vector<double> t = fpt_build_lfact_table(want);
    double *d = nullptr;
    CU(cudaMalloc(&d, t.size() * sizeof(double)));
    CU(cudaMemcpy(d, t.data(), t.size() * sizeof(double), cudaMemcpyHostToDevice));
    /* grow-only and never freed before fpt_release(): a launch in flight on another stream (or another thread that read the
       pointer a moment ago) keeps a valid table; the old one is a prefix of the new one */
    if (c->lf) c->lf_retired.push_back(c->lf);
    c->lf = d;
    c->lf_maxn = want;
    *table = d;
    return FPT_OK;
}

/* page-locked staging buffers are expensive to create (cudaMallocHost pins pages), so the host entry points keep a few
   per device and only ever grow them; the host entry points that use them hold g_host_mu[device] for the whole call, so
   concurrent callers serialise per device (the reference's own entry points share global state and are not re-entrant) */
static int pinned_slot(DeviceCtx *c, int slot, size_t bytes, void **out) {
    if (bytes > c->pinned_bytes[slot]) {
        if (c->pinned[slot]) cudaFreeHost(c->pinned[slot]);
        c->pinned[slot] = nullptr; c->pinned_bytes[slot] = 0;
        size_t want = std::max<size_t>(bytes + bytes / 4, 1 << 16);
        cudaError_t e = cudaMallocHost(&c->pinned[slot], want);
        if (e != cudaSuccess) return fail(FPT_ERR_CUDA, "cudaMallocHost(%zu bytes): %s", want, cudaGetErrorString(e));
        c->pinned_bytes[slot] = want;
    }
    *out = c->pinned[slot];
    return FPT_OK;
}

/* stream-ordered scratch allocations of one host-level call */
struct Arena {
    cudaStream_t st;
    std::vector<void *> blocks;
    std::vector<void *> pinned;
    explicit Arena(cudaStream_t s) : st(s) {}
    ~Arena() {
        for (void *p : blocks) cudaFreeAsync(p, st);
        cudaStreamSynchronize(st);
        for (void *p : pinned) cudaFreeHost(p);
    }
    template <typename T>
    int get(T **out, size_t count) {
        void *p = nullptr;
        size_t bytes = std::max<size_t>(count * sizeof(T), 16);
        cudaError_t e = cudaMallocAsync(&p, bytes, st);
        if (e != cudaSuccess) return fail(FPT_ERR_CUDA, "cudaMallocAsync(%zu bytes): %s", bytes, cudaGetErrorString(e));
        blocks.push_back(p);
        *out = (T *)p;
        return FPT_OK;
    }
    template <typename T>
    int host(T **out, size_t count) {
        void *p = nullptr;
        size_t bytes = std::max<size_t>(count * sizeof(T), 16);
        cudaError_t e = cudaMallocHost(&p, bytes);
        if (e != cudaSuccess) return fail(FPT_ERR_CUDA, "cudaMallocHost(%zu bytes): %s", bytes, cudaGetErrorString(e));
        pinned.push_back(p);
        *out = (T *)p;
        return FPT_OK;
    }
};

template <typename K>
static int persistent_grid(DeviceCtx *c, K kernel, int block, size_t smem, long long items, int *grid) {
    if (smem > 48 * 1024) {
        if ((int)smem > c->smem_optin) return fail(FPT_ERR_ARG, "kernel needs %zu bytes of shared memory, device allows %d", smem, c->smem_optin);
        CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem));
    if (per_sm < 1) per_sm = 1;
    long long g = (long long)per_sm * c->sms;
    if (items < g) g = items;
    if (g < 1) g = 1;
    *grid = (int)g;
    return FPT_OK;
}

/* ------------------------------------------------------------------------------------------------
 * Optional per-kernel timing: when enabled, every launch made through this library is bracketed by a
 * pair of CUDA events on the launching stream; fpt_profile_summary() synchronises and reports the
 * summed device time per kernel. Used by bench.py for the roofline of the dominant kernel.
 */
struct ProfRec { const char *name; cudaEvent_t a, b; };
static std::atomic<bool> g_prof_on{false};
static std::mutex g_prof_mu;                    /* guards the two vectors */
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_prof_free;

static cudaEvent_t prof_event() {
    cudaEvent_t e;
    {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        if (!g_prof_free.empty()) { e = g_prof_free.back(); g_prof_free.pop_back(); return e; }
    }
    cudaEventCreate(&e);
    return e;
}

struct ProfScope {
    const char *name; cudaStream_t st; cudaEvent_t a = nullptr;
    ProfScope(const char *n, cudaStream_t s) : name(n), st(s) {
        if (g_prof_on) { a = prof_event(); cudaEventRecord(a, st); }
    }
    ~ProfScope() {
        if (a) { cudaEvent_t b = prof_event(); cudaEventRecord(b, st); std::lock_guard<std::mutex> lk(g_prof_mu); g_prof.push_back({ name, a, b }); }
    }
};

extern "C" int fpt_profile_enable(int on) {
    g_prof_on = on != 0;
    return FPT_OK;
}

extern "C" int fpt_profile_summary(char *buf, size_t buflen) {
    struct Acc { const char *name; double ms; long n; };
    std::vector<Acc> acc;
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (ProfRec &r : g_prof) {
        cudaEventSynchronize(r.b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, r.a, r.b);
        bool found = false;
        for (Acc &a : acc) if (!strcmp(a.name, r.name)) { a.ms += ms; a.n++; found = true; break; }
        if (!found) acc.push_back({ r.name, (double)ms, 1 });
        g_prof_free.push_back(r.a); g_prof_free.push_back(r.b);
    }
    g_prof.clear();
    cudaGetLastError();
    size_t off = 0;
    if (buf && buflen) {
        off += snprintf(buf + off, buflen - off, "{");
        for (size_t i = 0; i < acc.size() && off < buflen; i++)
            off += snprintf(buf + off, buflen - off, "%s\"%s\": {\"launches\": %ld, \"ms\": %.6f}", i ? ", " : "", acc[i].name,
                            acc[i].n, acc[i].ms);
        if (off < buflen) snprintf(buf + off, buflen - off, "}");
    }
    return FPT_OK;
}

/* ================================================================================================ library state */
extern "C" const char *fpt_last_error(void) { return g_err; }

extern "C" int fpt_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int fpt_set_device(int device) {
    int n = fpt_device_count();
    if (device < 0 || device >= n) return fail(FPT_ERR_ARG, "device %d not in [0,%d)", device, n);
    g_device.store(device);
    CU(cudaSetDevice(device));
    return FPT_OK;
}

extern "C" void fpt_set_perm_mode(int chain) { g_perm_chain.store(chain != 0); }
extern "C" int fpt_get_perm_mode(void) { return g_perm_chain.load(); }
extern "C" void fpt_set_perm_large_kernel(int tensor_memory) { g_perm_umma.store(tensor_memory); }
extern "C" void fpt_set_lanczos_form(int max_form) { g_lanczos_form.store(max_form < 0 ? 0 : (max_form > 3 ? 3 : max_form)); }
extern "C" void fpt_set_perm_small_kernel(int v) { g_perm_small.store(v != 0); }
extern "C" void fpt_set_mds_small_kernel(int v) { g_mds_small.store(v != 0); }
extern "C" void fpt_set_lanczos_threads(int threads) { g_lanczos_threads.store(threads == 384 || threads == 256 ? threads : 512); }
extern "C" void fpt_set_k4_mode(int mode) { g_k4_mode.store(mode < 0 ? 0 : (mode > 2 ? 2 : mode)); }
extern "C" int fpt_debug_k4_phases(unsigned long long *out4) {
    unsigned long long zero[4] = { 0 };
    if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpyFromSymbol(out4, fpt_k4_phase_cycles, sizeof zero) != cudaSuccess ||
        cudaMemcpyToSymbol(fpt_k4_phase_cycles, zero, sizeof zero) != cudaSuccess) return FPT_ERR_CUDA;
    return FPT_OK;
}
extern "C" int fpt_debug_lanczos_phases(unsigned long long *out8) {
    unsigned long long zero[8] = { 0 };
    if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpyFromSymbol(out8, fpt_lanczos_phase_cycles, sizeof zero) != cudaSuccess ||
        cudaMemcpyToSymbol(fpt_lanczos_phase_cycles, zero, sizeof zero) != cudaSuccess) return FPT_ERR_CUDA;
    return FPT_OK;
}
extern "C" int fpt_debug_umma_phases(unsigned long long *out8) {
    unsigned long long zero[8] = { 0 };
    if (cudaDeviceSynchronize() != cudaSuccess || cudaMemcpyFromSymbol(out8, fpt_umma_phase_cycles, sizeof zero) != cudaSuccess ||
        cudaMemcpyToSymbol(fpt_umma_phase_cycles, zero, sizeof zero) != cudaSuccess) return FPT_ERR_CUDA;
    return FPT_OK;
}

/* exact re-evaluations the permutation kernel needed since the last call (its integer surrogate could not
   decide `permuted >= observed`); synchronises the device */
extern "C" long long fpt_css_perm_rechecks(void) {
    DeviceCtx *c;
    if (get_ctx(&c) != FPT_OK) return -1;
    unsigned long long v = 0;
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpy(&v, c->rechecks, sizeof v, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    cudaMemset(c->rechecks, 0, sizeof v);
    return (long long)v;
}

extern "C" void fpt_set_seed(uint64_t seed) { g_seed.store(seed); }
extern "C" uint64_t fpt_get_seed(void) { return g_seed.load(); }
extern "C" uint64_t fpt_window_state(uint64_t seed, int64_t window, int stream) {
    return fpt_stream_state(seed, (long long)window, stream);
}

extern "C" void fpt_release(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    for (int d = 0; d < 64; d++) {
        DeviceCtx &c = g_ctx[d];
        if (c.device < 0) continue;
        cudaSetDevice(d);
        cudaDeviceSynchronize();
        if (c.binom) cudaFree(c.binom);
        if (c.lf) cudaFree(c.lf);
        for (double *q : c.lf_retired) cudaFree(q);
        if (c.rechecks) cudaFree(c.rechecks);
        for (int k = 0; k < 8; k++) if (c.pinned[k]) cudaFreeHost(c.pinned[k]);
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, d) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
        c = DeviceCtx();
    }
    cudaGetLastError();
}

/* ================================================================================================ device API: FET */
static int count_tile(DeviceCtx *c, int asize, int bsize, size_t *smem) {
    /* SNPs per shared-memory tile: one byte per genotype, <= 40 KB, even (keeps 16-byte tile starts) */
    long long per = (long long)asize + bsize;
    long long tile = (40 * 1024) / per;
    if (tile > 512) tile = 512;
    tile &= ~1LL;
    if (tile < 2) tile = 2;
    *smem = (((size_t)tile * asize + 15) & ~(size_t)15) + (size_t)tile * bsize + 16;
    if ((int)*smem > c->smem_optin) return -1;
    return (int)tile;
}

template <typename T>
static int dev_fet_count(const T *a, const T *b, int64_t nsnp, int asize, int bsize, int32_t *tables, cudaStream_t st) {
    DeviceCtx *c;
    CHECK(get_ctx(&c));
    if (nsnp < 0 || asize <= 0 || bsize <= 0) return fail(FPT_ERR_ARG, "fet_count: nsnp=%lld asize=%d bsize=%d", (long long)nsnp, asize, bsize);
    if (nsnp == 0) return FPT_OK;
    size_t smem;
    int tile = count_tile(c, asize, bsize, &smem);
    if (tile < 0) return fail(FPT_ERR_ARG, "populations of %d+%d individuals exceed the shared-memory tile", asize, bsize);
    int grid;
    CHECK(persistent_grid(c, fpt_fet_count_kernel<T>, 256, smem, (nsnp + tile - 1) / tile, &grid));
    { ProfScope ps_("fet_count", st); fpt_fet_count_kernel<T><<<grid, 256, smem, st>>>(a, b, nsnp, asize, bsize, tile, (int4 *)tables); }
    CU(cudaGetLastError());
    return FPT_OK;
}

extern "C" int fpt_dev_fet_count_f64(const double *a, const double *b, int64_t nsnp, int asize, int bsize,
                                     int32_t *tables, void *stream) {
    return dev_fet_count<double>(a, b, nsnp, asize, bsize, tables, (cudaStream_t)stream);
}
extern "C" int fpt_dev_fet_count_i8(const int8_t *a, const int8_t *b, int64_t nsnp, int asize, int bsize,
                                    int32_t *tables, void *stream) {
    return dev_fet_count<signed char>((const signed char *)a, (const signed char *)b, nsnp, asize, bsize, tables,
                                      (cudaStream_t)stream);
}

extern "C" int fpt_dev_fet_score(const int32_t *tables, int64_t n, int max_n, int force_log, double *out, void *stream) {
    DeviceCtx *c;
    CHECK(get_ctx(&c));
    cudaStream_t st = (cudaStream_t)stream;
    if (n < 0) return fail(FPT_ERR_ARG, "fet_score: n=%lld", (long long)n);
    if (n == 0) return FPT_OK;
    if (max_n <= 0) {
        int *d_max;
        CU(cudaMallocAsync((void **)&d_max, sizeof(int), st));
        CU(cudaMemsetAsync(d_max, 0, sizeof(int), st));
        int g = (int)std::min<long long>((n + 255) / 256, (long long)c->sms * 8);
        { ProfScope ps_("fet_maxn", st); fpt_fet_maxn_kernel<<<g, 256, 0, st>>>((const int4 *)tables, n, d_max); }
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(&max_n, d_max, sizeof(int), cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        CU(cudaFreeAsync(d_max, st));
    }
    const double *lf = nullptr;
    CHECK(ensure_lf(c, max_n, &lf));
    size_t lf_bytes = ((size_t)max_n + 1) * sizeof(double);
    int lf_in_smem = lf_bytes <= 96 * 1024;
    size_t smem = FPT_BINOM_ENTRIES * sizeof(unsigned long long) + (lf_in_smem ? lf_bytes : 0);
    int grid;
    const size_t smem_sorted = smem + fpt_fet_sorted_smem_extra();
    if ((max_n > FPT_FET_EXACT_MAX_N || force_log) && n >= 4 * FPT_FET_TILE && smem_sorted <= (size_t)c->smem_optin) {
        /* tables beyond the exact domain walk for tens to hundreds of terms: sort each tile by estimated walk length first */
        CHECK(persistent_grid(c, fpt_fet_score_sorted_kernel, FPT_FET_SORT_THREADS, smem_sorted, (n + FPT_FET_TILE - 1) / FPT_FET_TILE, &grid));
        ProfScope ps_("fet_score", st);
        fpt_fet_score_sorted_kernel<<<grid, FPT_FET_SORT_THREADS, smem_sorted, st>>>((const int4 *)tables, n, c->binom, lf, max_n, lf_in_smem, force_log, out);
    } else {
        CHECK(persistent_grid(c, fpt_fet_score_kernel, 256, smem, (n + 255) / 256, &grid));
        ProfScope ps_("fet_score", st);
        fpt_fet_score_kernel<<<grid, 256, smem, st>>>((const int4 *)tables, n, c->binom, lf, max_n, lf_in_smem, force_log, out);
    }
    CU(cudaGetLastError());
    return FPT_OK;
}

static int check_range(const fpt_scan_range *r, long long *nwin) {
    if (!r) return fail(FPT_ERR_ARG, "scan range is NULL");
    if (r->wsize <= 0 || r->wstep <= 0 || r->regend < 0) return fail(FPT_ERR_ARG, "bad window geometry regend=%d wsize=%d wstep=%d", r->regend, r->wsize, r->wstep);
    if (r->window_begin < 0 || r->window_end < r->window_begin) return fail(FPT_ERR_ARG, "bad window range [%lld,%lld)", (long long)r->window_begin, (long long)r->window_end);
    if (r->semantics != FPT_SCAN_SERIAL && r->semantics != FPT_SCAN_THREADED) return fail(FPT_ERR_ARG, "unknown scan semantics %d", r->semantics);
    *nwin = r->window_end - r->window_begin;
    return FPT_OK;
}

extern "C" int fpt_dev_window_table(const int32_t *pos, int64_t nsnp, const fpt_scan_range *r, int32_t *wleft,
                                    int32_t *wright, int32_t *max_npos, void *stream) {
    DeviceCtx *c;
    CHECK(get_ctx(&c));
    long long nwin;
    CHECK(check_range(r, &nwin));
    if (nwin == 0) return FPT_OK;
    unsigned grid = (unsigned)((nwin + 255) / 256);
    { ProfScope ps_("window_table", (cudaStream_t)stream); fpt_window_table_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pos, nsnp, r->window_begin, nwin, r->regend, r->wsize,
                                                                   r->wstep, r->semantics, wleft, wright, max_npos); }
    CU(cudaGetLastError());
    return FPT_OK;
}

extern "C" int fpt_dev_fet_windows(const double *snp_scores, const int32_t *wleft, const int32_t *wright,
                                   const fpt_scan_range *r, int max_npos, double perc, const uint64_t *states,
                                   double *scores, double *stddev, uint8_t *written, void *stream) {
    DeviceCtx *c;
    CHECK(get_ctx(&c));
    long long nwin;
    CHECK(check_range(r, &nwin));
    if (nwin == 0 || max_npos <= 0) return FPT_OK;
    if (!(perc >= 0.0 && perc <= 1.0)) return fail(FPT_ERR_ARG, "percentile %g not in [0,1]", perc);
    int npad = 2;
    while (npad < max_npos) npad <<= 1;
    int use_hist = max_npos <= FPT_FET_HIST_MAX_NPOS;
    size_t smem = (size_t)npad * sizeof(double) + (use_hist ? (size_t)FPT_FET_HIST_LD * max_npos * sizeof(unsigned short) : 0);
    if ((int)smem > c->smem_optin)
        return fail(FPT_ERR_WINDOW_TOO_LARGE, "a window holds %d SNPs; at most %d fit one CTA's shared memory", max_npos,
                    c->smem_optin / 8);
    int grid;
    CHECK(persistent_grid(c, fpt_fet_window_kernel, 128, smem, nwin, &grid));
    { ProfScope ps_("fet_window", (cudaStream_t)stream); fpt_fet_window_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>(snp_scores, wleft, wright, r->window_begin, nwin, perc,
                                                                    r->seed, states, npad, use_hist, scores, stddev, written); }
    CU(cudaGetLastError());
    return FPT_OK;
}

/* ================================================================================================ device API: CSS */
extern "C" size_t fpt_dev_css_planes_bytes(int64_t nsnp, int m) {
    return (size_t)((nsnp + 31) / 32) * 2 * (size_t)m * sizeof(uint32_t);
}

template <typename T>
static int dev_css_pack(const T *a, const T *b, int64_t nsnp, int asize, int bsize, uint32_t *planes, cudaStream_t st) {
    DeviceCtx *c;
    CHECK(get_ctx(&c));
    if (nsnp < 0 || asize <= 0 || bsize <= 0) return fail(FPT_ERR_ARG, "css_pack: nsnp=%lld asize=%d bsize=%d", (long long)nsnp, asize, bsize);
    if (nsnp == 0) return FPT_OK;
    long long per = (long long)asize + bsize;
    long long wpt = (40 * 1024) / (per * 32);
    if (wpt > 8) wpt = 8;
    if (wpt < 1) wpt = 1;
    size_t smem = (((size_t)wpt * 32 * asize + 15) & ~(size_t)15) + (size_t)wpt * 32 * bsize + 16;
    if ((int)smem > c->smem_optin) return fail(FPT_ERR_ARG, "populations of %d+%d individuals exceed the shared-memory tile", asize, bsize);
    long long nwords = (nsnp + 31) / 32;
    int grid;
    CHECK(persistent_grid(c, fpt_css_pack_kernel<T>, 256, smem, (nwords + wpt - 1) / wpt, &grid));
    { ProfScope ps_("css_pack", st); fpt_css_pack_kernel<T><<<grid, 256, smem, st>>>(a, b, nsnp, asize, bsize, (int)wpt, planes); }
    CU(cudaGetLastError());
    return FPT_OK;
}

extern "C" int fpt_dev_css_pack_f64(const double *a, const double *b, int64_t nsnp, int asize, int bsize,
                                    uint32_t *planes, void *stream) {
    return dev_css_pack<double>(a, b, nsnp, asize, bsize, planes, (cudaStream_t)stream);
}
extern "C" int fpt_dev_css_pack_i8(const int8_t *a, const int8_t *b, int64_t nsnp, int asize, int bsize,
                                   uint32_t *planes, void *stream) {
    return dev_css_pack<signed char>((const signed char *)a, (const signed char *)b, nsnp, asize, bsize, planes,
                                     (cudaStream_t)stream);
}

extern "C" int fpt_dev_css_absdiff(const double *a, const double *b, int64_t nsnp, double *out, void *stream) {
    DeviceCtx *c;
    CHECK(get_ctx(&c));
    if (nsnp <= 0) return FPT_OK;
    int g = (int)std::min<long long>((nsnp + 255) / 256, (long long)c->sms * 8);
    { ProfScope ps_("css_absdiff", (cudaStream_t)stream); fpt_css_absdiff_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(a, b, nsnp, out); }
    CU(cudaGetLastError());
    return FPT_OK;
}

/* launch geometry of the per-window CSS kernels for a cohort of m individuals */
struct CssPlan {
    int m, wch, mats_in_smem;
    size_t smem_win;                 /* CTA-per-window kernels (SMACOF) */
    size_t smem_large; int wch_large; /* large cohorts: Lanczos kernel, one 512-thread CTA per window */
    int mds_warps;                   /* > 0: classical MDS runs one warp per window, this many warps per CTA */
    size_t smem_mds_warp;
    int perm_threads, wide_tracks, dist_in_smem, tracks_in_smem, perm_sur;
    size_t smem_perm, perm_scratch_per_cta;
    int perm2, qbits;                /* second-generation permutation kernel (labels in bytes, everything in smem) */
    size_t smem_perm2;
    int perm_umma; size_t smem_umma; /* large cohorts, independent shuffles: tcgen05 contraction, one CTA per SM */
    int max_ctas;                    /* upper bound on persistent CTAs (sizes the global scratch) */
};

static CssPlan css_plan(const DeviceCtx *c, int m, const Knobs &kn) {
    CssPlan p;
    p.m = m;
    p.wch = 8;
    const size_t budget = (size_t)c->smem_optin - 1024;
    p.mats_in_smem = fpt_css_smem_bytes(m, p.wch, 1) <= budget;
    if (!p.mats_in_smem && fpt_css_smem_bytes(m, p.wch, 0) > budget) p.wch = 1;
    p.smem_win = fpt_css_smem_bytes(m, p.wch, p.mats_in_smem);
    const size_t per_warp = fpt_tridiag_work_bytes(m, 3);   /* three 32-SNP words per pass fill the vectors' space exactly */
    p.mds_warps = 2 * per_warp <= 32 * 1024 ? 2 : (per_warp <= budget ? 1 : 0);   /* small CTAs: many resident warps */
    p.smem_mds_warp = per_warp * (p.mds_warps > 0 ? p.mds_warps : 1);
    p.perm_threads = 256;
    p.wide_tracks = m > 256;
    const int tb = p.wide_tracks ? 2 : 1;
    p.dist_in_smem = (size_t)m * m * 8 + (size_t)4 * 1024 <= budget;
    /* general kernel: with the distance matrix in global memory the exact evaluation is a quarter of a million scattered
       loads per permutation, so the integer surrogate goes on (its membership rows want 128-thread CTAs) */
    p.perm_sur = !p.dist_in_smem && fpt_css_perm_smem_bytes(m, 128, tb, 0, 0, 1) <= budget;
    if (p.perm_sur) p.perm_threads = 128;
    p.tracks_in_smem = fpt_css_perm_smem_bytes(m, p.perm_threads, tb, p.dist_in_smem, 1, p.perm_sur) <= budget;
    p.smem_perm = fpt_css_perm_smem_bytes(m, p.perm_threads, tb, p.dist_in_smem, p.tracks_in_smem, p.perm_sur);
    size_t per = (p.dist_in_smem ? 0 : (size_t)m * m * 8) +
                 (p.tracks_in_smem ? 0 : (((size_t)2 * p.perm_threads * m * tb + 15) & ~(size_t)15)) +
                 (p.perm_sur ? fpt_perm_large_sur_scratch(m) : 0);
    p.perm_scratch_per_cta = (per + 255) & ~(size_t)255;
    /* cohorts beyond the one-warp path keep two m x m matrices per CTA in global scratch: bound the persistent grid */
    p.max_ctas = p.mds_warps > 0 ? c->sms * 16 : c->sms * 2;
    p.wch_large = 2;
    p.smem_large = fpt_lanczos_smem_bytes(m, p.wch_large);
    p.smem_perm2 = fpt_css_perm2_smem_bytes(m, p.perm_threads, kn.perm_chain);
    p.perm2 = m <= 250 && p.smem_perm2 <= budget;
    p.qbits = 8;
    p.smem_umma = fpt_umma_smem_bytes(m);
    p.perm_umma = !p.perm2 && p.perm_sur && !kn.perm_chain && kn.perm_umma && m <= 1024 && p.smem_umma <= budget;
    if (p.perm_umma) p.perm_scratch_per_cta = std::max(p.perm_scratch_per_cta, (fpt_umma_scratch_bytes(m) + 1023) & ~(size_t)1023);
    return p;
}

#define FPT_K4_BATCH 4096                         /* windows per pass of the code route: bounds the code buffer (2 MB per window at m = 1000, so 8 GB at most) */
struct CssWorkspace {
    double *X, *Xruns, *sigma, *evals, *gscratch, *tri, *refl, *basis;
    unsigned char *codes;
    int *iters;
    unsigned char *perm_scratch;
    size_t total;
};

static CssWorkspace css_carve(const CssPlan &p, long long nwin, int mds, unsigned char *base) {
    CssWorkspace w;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const int nruns = mds == 1 ? 4 : (mds == 2 ? 1 : 0);
    const size_t m2 = (size_t)2 * p.m;
    size_t oX = take((size_t)nwin * m2 * 8);
    size_t oXr = take((size_t)nwin * nruns * m2 * 8);
    size_t oS = take((size_t)nwin * nruns * 8);
    size_t oE = take((size_t)nwin * 3 * 8);
    size_t oI = take((size_t)nwin * nruns * 4);
    const bool need_g = !p.mats_in_smem || p.mds_warps == 0;
    /* large cohorts: the legacy matrices (per CTA) and the code route's buffers (codes of one batch of windows + a Lanczos
       basis per CTA) are never live together (SMACOF's matrices are used after classical MDS is done): one region */
    const size_t legacy_g = need_g ? (size_t)p.max_ctas * fpt_css_mats_doubles(p.m) * 8 : 0;
    const size_t codes_b = p.mds_warps == 0 ? (size_t)std::min<long long>(nwin, FPT_K4_BATCH) * fpt_k4_window_stride(p.m) : 0;
    const size_t basis_b = p.mds_warps == 0 ? (size_t)p.max_ctas * fpt_lanczos_cta_scratch_bytes(p.m) : 0;
    size_t oG = take(std::max(legacy_g, codes_b + basis_b));
    size_t oP = take((size_t)p.max_ctas * p.perm_scratch_per_cta);
    const bool warp_mds = p.mds_warps > 0 && mds != 1;      /* tridiagonal + reflectors handed from phase A to phase B */
    size_t oT = take(warp_mds ? (size_t)nwin * 3 * p.m * 8 : 0);
    size_t oR = take(warp_mds ? (size_t)nwin * ((size_t)p.m * (p.m - 1) / 2) * 8 : 0);
    w.total = off + 256;
    if (base) {
        w.X = (double *)(base + oX); w.Xruns = (double *)(base + oXr); w.sigma = (double *)(base + oS);
        w.evals = (double *)(base + oE); w.iters = (int *)(base + oI);
        w.gscratch = need_g ? (double *)(base + oG) : nullptr;
        w.basis = (double *)(base + oG);
        w.codes = base + oG + basis_b;
        w.perm_scratch = p.perm_scratch_per_cta ? base + oP : nullptr;
        w.tri = (double *)(base + oT); w.refl = (double *)(base + oR);
    }
    return w;
}

extern "C" size_t fpt_dev_css_workspace_bytes(int m, int64_t nwin, int mds) {
    DeviceCtx *c;
    if (get_ctx(&c) != FPT_OK || m <= 0 || nwin < 0) return 0;
    /* the largest layout over the kernel routes the process-wide switches can select, so that a workspace sized before
       fpt_set_perm_mode / fpt_set_perm_large_kernel is still accepted afterwards */
    size_t total = 0;
    for (int chain = 0; chain < 2; chain++)
        for (int umma = 0; umma < 2; umma++) {
            Knobs kn = knobs_now();
            kn.perm_chain = chain; kn.perm_umma = umma;
            total = std::max(total, css_carve(css_plan(c, m, kn), nwin, mds, nullptr).total);
        }
    return total;
}

template <typename TrackT>
static int launch_perm(DeviceCtx *c, const Knobs &kn, const CssPlan &p, const CssWorkspace &ws, int asize, int bsize, long long wbase,
                       long long nwin, const uint8_t *status, int treshold, int runs, uint64_t seed, const uint64_t *states,
                       double *scores, double *pv, int *hits, int *nperm, cudaStream_t st) {
    int grid;
    CHECK(persistent_grid(c, fpt_css_perm_kernel<TrackT>, p.perm_threads, p.smem_perm, nwin, &grid));
    grid = std::min(grid, p.max_ctas);
    { ProfScope ps_("css_perm", st); fpt_css_perm_kernel<TrackT><<<grid, p.perm_threads, p.smem_perm, st>>>(
        ws.X, p.m, asize, bsize, wbase, nwin, status, treshold, runs, seed, states, kn.perm_chain, p.dist_in_smem, p.tracks_in_smem,
        (double *)ws.perm_scratch, p.perm_scratch_per_cta, p.perm_sur ? 31 : 0, scores, pv, hits, nperm, c->rechecks); }
    CU(cudaGetLastError());
    return FPT_OK;
}

static int launch_perm_umma(DeviceCtx *c, const Knobs &kn, const CssPlan &p, const CssWorkspace &ws, int asize, int bsize, long long wbase,
                            long long nwin, const uint8_t *status, int treshold, int runs, uint64_t seed, const uint64_t *states,
                            double *scores, double *pv, int *hits, int *nperm, cudaStream_t st) {
    /* observed scores first: one warp per window, its result handed over in `scores` */
    const size_t smem_obs = fpt_css_observed_smem_bytes(p.m);
    int grid_obs;
    CHECK(persistent_grid(c, fpt_css_observed_kernel, FPT_OBS_WARPS * 32, smem_obs, (nwin + FPT_OBS_WARPS - 1) / FPT_OBS_WARPS, &grid_obs));
    { ProfScope ps_("css_observed", st); fpt_css_observed_kernel<<<grid_obs, FPT_OBS_WARPS * 32, smem_obs, st>>>(
        ws.X, p.m, asize, bsize, nwin, status, scores); }
    CU(cudaGetLastError());
    /* The kernel allocates all 512 columns of tensor memory, so two of its CTAs must never share an SM (the second would
       sit in tcgen05.alloc until the first exits — also across streams). Enforced here, not assumed: the dynamic shared
       memory request is padded above half of an SM's capacity, which makes co-residency impossible whatever m is. */
    const size_t smem_launch = std::max(p.smem_umma, (size_t)c->smem_optin / 2 + 2048);
    CU(cudaFuncSetAttribute(fpt_css_perm_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_launch));
    int per_sm = 0;
    CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fpt_css_perm_umma_kernel, FPT_UMMA_THREADS, smem_launch));
    if (per_sm != 1) return fail(FPT_ERR_CUDA, "tensor-memory permutation kernel: %d CTAs per SM (exactly 1 required)", per_sm);
    const int grid = (int)std::max(1LL, std::min<long long>(std::min(c->sms, p.max_ctas), nwin));
    { ProfScope ps_("css_perm", st); fpt_css_perm_umma_kernel<<<grid, FPT_UMMA_THREADS, smem_launch, st>>>(
        ws.X, p.m, asize, bsize, wbase, nwin, status, treshold, runs, seed, states, ws.perm_scratch, p.perm_scratch_per_cta,
        kn.perm_umma == 2 ? 10 : 31, scores, pv, hits, nperm, c->rechecks); }
    CU(cudaGetLastError());
    return FPT_OK;
}

static int dev_css_windows(const Knobs &kn, const uint32_t *planes, const double *absdiff, int asize, int bsize,
                           const int32_t *wleft, const int32_t *wright, const fpt_scan_range *r, int treshold,
                           int runs, int mds, void *workspace, size_t workspace_bytes, double *scores, double *pv,
                           uint8_t *status, const fpt_css_probes *probes, void *stream) {
    DeviceCtx *c;
    CHECK(get_ctx(&c));
    cudaStream_t st = (cudaStream_t)stream;
    long long nwin;
    CHECK(check_range(r, &nwin));
    if (asize <= 0 || bsize <= 0) return fail(FPT_ERR_ARG, "css: asize=%d bsize=%d", asize, bsize);
    if (mds < 0 || mds > 2) return fail(FPT_ERR_ARG, "css: mds must be 0, 1 or 2 (got %d)", mds);
    if (absdiff && (asize != 1 || bsize != 1)) return fail(FPT_ERR_ARG, "the frequency metric needs one track per population (got %d+%d)", asize, bsize);
    if (!absdiff && !planes) return fail(FPT_ERR_ARG, "css: neither bit-planes nor frequency differences given");
    if (nwin == 0) return FPT_OK;
    const int m = asize + bsize;
    CssPlan p = css_plan(c, m, kn);
    if (p.smem_win > (size_t)c->smem_optin || p.smem_perm > (size_t)c->smem_optin)
        return fail(FPT_ERR_ARG, "cohort of %d individuals does not fit the kernels' shared-memory plan", m);
    CssWorkspace ws = css_carve(p, nwin, mds, (unsigned char *)workspace);
    if (!workspace || workspace_bytes < ws.total)
        return fail(FPT_ERR_ARG, "css workspace too small: %zu < %zu bytes", workspace_bytes, ws.total);
    const uint64_t *st_perm = r->states_resample, *st_init = r->states_init;

    int grid;
    if (mds == 0 || mds == 2) {
        if (p.mds_warps > 0) {
            if (kn.mds_small && planes && !absdiff && fpt_tridiag_reg_ok(m)) {
                /* the genome scans' cohorts: the whole matrix in the registers of the warp */
                const int block = 32 * FPT_TREG_WARPS, pad = fpt_tridiag_reg_pad(m);
                const size_t smem_r = fpt_tridiag_reg_work_bytes(m) * FPT_TREG_WARPS;
                const long long items = (nwin + FPT_TREG_WARPS - 1) / FPT_TREG_WARPS;
                ProfScope ps_("css_tridiag", st);
                if (pad == 32) {
                    CHECK(persistent_grid(c, fpt_css_tridiag_reg_kernel<4, 8>, block, smem_r, items, &grid));
                    fpt_css_tridiag_reg_kernel<4, 8><<<grid, block, smem_r, st>>>(planes, m, wleft, wright, nwin, ws.tri, ws.refl, status);
                } else if (pad == 40) {
                    CHECK(persistent_grid(c, fpt_css_tridiag_reg_kernel<5, 10>, block, smem_r, items, &grid));
                    fpt_css_tridiag_reg_kernel<5, 10><<<grid, block, smem_r, st>>>(planes, m, wleft, wright, nwin, ws.tri, ws.refl, status);
                } else {
                    CHECK(persistent_grid(c, fpt_css_tridiag_reg_kernel<6, 12>, block, smem_r, items, &grid));
                    fpt_css_tridiag_reg_kernel<6, 12><<<grid, block, smem_r, st>>>(planes, m, wleft, wright, nwin, ws.tri, ws.refl, status);
                }
            } else {
                const int block = 32 * p.mds_warps;
                CHECK(persistent_grid(c, fpt_css_tridiag_kernel, block, p.smem_mds_warp, (nwin + p.mds_warps - 1) / p.mds_warps, &grid));
                ProfScope ps_("css_tridiag", st);
                fpt_css_tridiag_kernel<<<grid, block, p.smem_mds_warp, st>>>(planes, absdiff, m, wleft, wright, nwin, 3, ws.tri, ws.refl, status);
            }
            CU(cudaGetLastError());
            const size_t smem_b = fpt_eigvec_work_bytes(m) * 4;
            CHECK(persistent_grid(c, fpt_css_eigvec_kernel, 128, smem_b, (nwin + 3) / 4, &grid));
            { ProfScope ps_("css_eigvec", st); fpt_css_eigvec_kernel<<<grid, 128, smem_b, st>>>(m, nwin, ws.tri, ws.refl, status, ws.X,
                                                                                                (probes && probes->evals) ? ws.evals : nullptr); }
        } else {                                           /* cohorts too large for a warp's shared-memory slice: Lanczos, CTA per window */
            if (p.smem_large > (size_t)c->smem_optin)
                return fail(FPT_ERR_ARG, "cohort of %d individuals does not fit the large-cohort kernel's shared memory", m);
            /* code route: the genotype-distance matrix as count codes from the u8 GEMM (or the popcount kernel), then Lanczos on
               the codes. Needs bit-planes and counts that fit 16 bits (a window of wsize bp holds at most wsize + 1 positions). */
            const bool code_route = kn.k4_mode > 0 && planes && !absdiff && r->wsize < 65535 && kn.lanczos_form >= 2;
            if (code_route) {
                const size_t stride = fpt_k4_window_stride(m);
                const size_t smem_lz = fpt_lanczos_smem_bytes(m, 0);
                const int lzt = kn.lanczos_threads;
                const bool arith = kn.lanczos_form >= 3;
                int grid_lz, grid_lza = 0;
                if (lzt == 384) CHECK(persistent_grid(c, fpt_css_mds_codes_kernel<384, false>, 384, smem_lz, nwin, &grid_lz));
                else if (lzt == 256) CHECK(persistent_grid(c, fpt_css_mds_codes_kernel<256, false>, 256, smem_lz, nwin, &grid_lz));
                else CHECK(persistent_grid(c, fpt_css_mds_codes_kernel<512, false>, 512, smem_lz, nwin, &grid_lz));
                if (arith) {
                    if (lzt == 384) CHECK(persistent_grid(c, fpt_css_mds_codes_kernel<384, true>, 384, smem_lz, nwin, &grid_lza));
                    else if (lzt == 256) CHECK(persistent_grid(c, fpt_css_mds_codes_kernel<256, true>, 256, smem_lz, nwin, &grid_lza));
                    else CHECK(persistent_grid(c, fpt_css_mds_codes_kernel<512, true>, 512, smem_lz, nwin, &grid_lza));
                    grid_lza = std::min(grid_lza, p.max_ctas);
                }
                grid_lz = std::min(grid_lz, p.max_ctas);
                /* the Lanczos CTAs walk a pass's windows at a fixed stride (window blockIdx.x, + gridDim.x, ...), so a pass whose window
                   count is not a multiple of the grid ends with a round in which part of the SMs idle: passes of whole rounds
                   (with a cap of 1024 windows: 888 for 296 CTAs, and a 2600-window chromosome took 9 rounds instead of 4 + 4 + 2; with the cap of
                   4096 a configs[4] chromosome is one pass of 8-9 windows per CTA) */
                const long long grid_main = std::max(1, arith ? grid_lza : grid_lz);
                const long long pass = grid_main <= FPT_K4_BATCH ? (FPT_K4_BATCH / grid_main) * grid_main : FPT_K4_BATCH;
                for (long long w0 = 0; w0 < nwin; w0 += pass) {
                    const long long nb = std::min<long long>(pass, nwin - w0);
                    if (kn.k4_mode == 2) {
                        CU(cudaFuncSetAttribute(fpt_css_k4_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FPT_K4_SMEM));
                        const int g4 = (int)std::max(1LL, std::min<long long>(c->sms, nb));     /* 512 columns of tensor memory: one CTA per SM */
                        { ProfScope ps_("css_k4", st); fpt_css_k4_umma_kernel<<<g4, FPT_K4_THREADS, FPT_K4_SMEM, st>>>(planes, m, wleft + w0, wright + w0, nb, ws.codes, stride); }
                    } else {
                        const size_t smem4 = fpt_k4_popc_smem(m);
                        if (smem4 > (size_t)c->smem_optin) return fail(FPT_ERR_ARG, "cohort of %d individuals exceeds the popcount kernel's shared memory", m);
                        int g4;
                        CHECK(persistent_grid(c, fpt_css_k4_popc_kernel, 256, smem4, nb, &g4));
                        { ProfScope ps_("css_k4", st); fpt_css_k4_popc_kernel<<<g4, 256, smem4, st>>>(planes, m, wleft + w0, wright + w0, nb, ws.codes, stride); }
                    }
                    CU(cudaGetLastError());
                    {
                        ProfScope ps_("css_mds_large", st);
                        const int g_ = (int)std::min<long long>(grid_lz, nb), ga_ = (int)std::min<long long>(grid_lza, nb);
                        double *X_ = ws.X + (size_t)w0 * 2 * m, *ev_ = ws.evals + (size_t)w0 * 3;
                        /* arithmetic-product kernel first; the windows it leaves pending (more than 255 SNPs, too many blanks) go to the general one */
#define FPT_LZ_LAUNCH(T_, A_, G_, P_) fpt_css_mds_codes_kernel<T_, A_><<<G_, T_, smem_lz, st>>>(ws.codes, stride, m, wleft + w0, wright + w0, nb, ws.basis, X_, ev_, status + w0, nullptr, P_)
                        if (arith) { if (lzt == 384) FPT_LZ_LAUNCH(384, true, ga_, 0); else if (lzt == 256) FPT_LZ_LAUNCH(256, true, ga_, 0); else FPT_LZ_LAUNCH(512, true, ga_, 0); }
                        const int pend_ = arith ? 1 : 0;
                        if (lzt == 384) FPT_LZ_LAUNCH(384, false, g_, pend_); else if (lzt == 256) FPT_LZ_LAUNCH(256, false, g_, pend_); else FPT_LZ_LAUNCH(512, false, g_, pend_);
#undef FPT_LZ_LAUNCH
                    }
                    CU(cudaGetLastError());
                }
            } else {
            CHECK(persistent_grid(c, fpt_css_mds_large_kernel, 512, p.smem_large, nwin, &grid));
            grid = std::min(grid, p.max_ctas);
            { ProfScope ps_("css_mds_large", st); fpt_css_mds_large_kernel<<<grid, 512, p.smem_large, st>>>(planes, absdiff, m, wleft, wright, nwin, p.wch_large,
                                                             ws.gscratch, ws.X, ws.evals, status, nullptr, kn.lanczos_form); }
            }
        }
        CU(cudaGetLastError());
    }
    if (mds == 1 || mds == 2) {
        const int nruns = mds == 1 ? 4 : 1;
        /* block size: the pair pass runs ceil(npairs / threads) rounds, so pick the multiple of 32 (64..256) that wastes
           the fewest thread-rounds (m = 40: 780 pairs -> 160 threads x 5 rounds instead of 128 x 7) */
        int sthreads = 128;
        {
            const long long npairs = (long long)m * (m - 1) / 2;
            long long best = -1;
            for (int t = 64; t <= 256; t += 32) {
                const long long cost = ((npairs + t - 1) / t) * t;
                if (best < 0 || cost < best) { best = cost; sthreads = t; }
            }
        }
        if (p.mats_in_smem) CHECK(persistent_grid(c, fpt_css_smacof_kernel<true>, sthreads, p.smem_win, nwin * nruns, &grid));
        else CHECK(persistent_grid(c, fpt_css_smacof_kernel<false>, sthreads, p.smem_win, nwin * nruns, &grid));
        grid = std::min(grid, p.max_ctas);
        {
            ProfScope ps_("css_smacof", st);
            if (p.mats_in_smem) fpt_css_smacof_kernel<true><<<grid, sthreads, p.smem_win, st>>>(planes, absdiff, m, wleft, wright, r->window_begin, nwin, p.wch,
                                                            1, ws.gscratch, nruns, mds == 1, r->seed, st_init, 300, 0.000001, ws.X, ws.Xruns, ws.sigma, ws.iters, status);
            else fpt_css_smacof_kernel<false><<<grid, sthreads, p.smem_win, st>>>(planes, absdiff, m, wleft, wright, r->window_begin, nwin, p.wch,
                                                            0, ws.gscratch, nruns, mds == 1, r->seed, st_init, 300, 0.000001, ws.X, ws.Xruns, ws.sigma, ws.iters, status);
        }
        CU(cudaGetLastError());
        int g2 = (int)std::min<long long>(nwin, (long long)c->sms * 16);
        { ProfScope ps_("css_pick", st); fpt_css_pick_kernel<<<g2, 64, 0, st>>>(ws.Xruns, ws.sigma, m, nruns, nwin, status, ws.X); }
        CU(cudaGetLastError());
    }
    int *hits = probes ? probes->hits : nullptr, *nperm = probes ? probes->nperm : nullptr;
    if (p.perm2) {
        /* quantisation width: |smaller group| * m * 2^qbits must stay below 2^31 (integer surrogate sums) */
        const long long terms = (long long)std::min(asize, bsize) * m + 1;
        int qb = fpt_css_perm2_uses_mma(m) ? 23 : 22;     /* tensor-core path: three base-256 digits (q <= 2^23 fits 24 bits) */
        while (qb > 4 && (terms << qb) >= (1LL << 31)) qb--;
        const size_t smem3 = fpt_css_perm3_smem_bytes(m);
        if (kn.perm_small && fpt_css_perm3_ok(m, kn.perm_chain) && ((qb + 8) >> 3) == 3 && smem3 <= (size_t)c->smem_optin) {
            /* observed scores first (one warp per window, handed over in `scores`), then the permutation test */
            const size_t smem_obs = fpt_css_observed_smem_bytes(m);
            int grid_obs;
            CHECK(persistent_grid(c, fpt_css_observed_kernel, FPT_OBS_WARPS * 32, smem_obs, (nwin + FPT_OBS_WARPS - 1) / FPT_OBS_WARPS, &grid_obs));
            { ProfScope ps_("css_observed", st); fpt_css_observed_kernel<<<grid_obs, FPT_OBS_WARPS * 32, smem_obs, st>>>(ws.X, m, asize, bsize, nwin, status, scores); }
            CU(cudaGetLastError());
            if (fpt_css_perm3_ksteps(m) == 2) {
                CHECK(persistent_grid(c, fpt_css_perm3_kernel<2>, FPT_P3_T, smem3, nwin, &grid));
                { ProfScope ps_("css_perm", st); fpt_css_perm3_kernel<2><<<grid, FPT_P3_T, smem3, st>>>(
                      ws.X, m, asize, bsize, r->window_begin, nwin, status, treshold, runs, r->seed, st_perm, qb, scores, pv, hits, nperm, c->rechecks); }
            } else {
                CHECK(persistent_grid(c, fpt_css_perm3_kernel<1>, FPT_P3_T, smem3, nwin, &grid));
                { ProfScope ps_("css_perm", st); fpt_css_perm3_kernel<1><<<grid, FPT_P3_T, smem3, st>>>(
                      ws.X, m, asize, bsize, r->window_begin, nwin, status, treshold, runs, r->seed, st_perm, qb, scores, pv, hits, nperm, c->rechecks); }
            }
        } else {
            CHECK(persistent_grid(c, fpt_css_perm2_kernel, p.perm_threads, p.smem_perm2, nwin, &grid));
            { ProfScope ps_("css_perm", st); fpt_css_perm2_kernel<<<grid, p.perm_threads, p.smem_perm2, st>>>(
                  ws.X, m, asize, bsize, r->window_begin, nwin, status, treshold, runs, r->seed, st_perm, kn.perm_chain, qb, scores, pv,
                  hits, nperm, c->rechecks); }
        }
        CU(cudaGetLastError());
    } else if (p.perm_umma) {
        CHECK(launch_perm_umma(c, kn, p, ws, asize, bsize, r->window_begin, nwin, status, treshold, runs, r->seed, st_perm, scores, pv,
                               hits, nperm, st));
    } else if (p.wide_tracks)
        CHECK(launch_perm<unsigned short>(c, kn, p, ws, asize, bsize, r->window_begin, nwin, status, treshold, runs, r->seed,
                                          st_perm, scores, pv, hits, nperm, st));
    else
        CHECK(launch_perm<unsigned char>(c, kn, p, ws, asize, bsize, r->window_begin, nwin, status, treshold, runs, r->seed,
                                         st_perm, scores, pv, hits, nperm, st));
    if (probes) {
        const int nruns = mds == 1 ? 4 : (mds == 2 ? 1 : 0);
        if (probes->X) CU(cudaMemcpyAsync(probes->X, ws.X, (size_t)nwin * 2 * m * 8, cudaMemcpyDeviceToDevice, st));
        if (probes->evals) {
            if (mds == 1) CU(cudaMemsetAsync(probes->evals, 0, (size_t)nwin * 3 * 8, st));
            else CU(cudaMemcpyAsync(probes->evals, ws.evals, (size_t)nwin * 3 * 8, cudaMemcpyDeviceToDevice, st));
        }
        if (probes->smacof_iters && nruns) CU(cudaMemcpyAsync(probes->smacof_iters, ws.iters, (size_t)nwin * nruns * 4, cudaMemcpyDeviceToDevice, st));
        if (probes->smacof_sigma && nruns) CU(cudaMemcpyAsync(probes->smacof_sigma, ws.sigma, (size_t)nwin * nruns * 8, cudaMemcpyDeviceToDevice, st));
    }
    return FPT_OK;
}

extern "C" int fpt_dev_css_windows(const uint32_t *planes, const double *absdiff, int asize, int bsize,
                                   const int32_t *wleft, const int32_t *wright, const fpt_scan_range *r, int treshold,
                                   int runs, int mds, void *workspace, size_t workspace_bytes, double *scores, double *pv,
                                   uint8_t *status, const fpt_css_probes *probes, void *stream) {
    return dev_css_windows(knobs_now(), planes, absdiff, asize, bsize, wleft, wright, r, treshold, runs, mds, workspace, workspace_bytes,
                           scores, pv, status, probes, stream);
}

/* ================================================================================================ host entry points */
static int check_genotypes(const fpt_genotypes *g) {
    if (!g) return fail(FPT_ERR_ARG, "genotypes are NULL");
    const bool f64 = g->avals && g->bvals, i8 = g->acodes && g->bcodes;
    if (f64 == i8) return fail(FPT_ERR_ARG, "give either float64 values or int8 codes for both populations");
    if (g->nsnp < 0 || g->asize <= 0 || g->bsize <= 0) return fail(FPT_ERR_ARG, "nsnp=%lld asize=%d bsize=%d", (long long)g->nsnp, g->asize, g->bsize);
    if (g->nsnp > 0x7fffffffLL) return fail(FPT_ERR_ARG, "more than 2^31-1 SNPs in one scan");
    return FPT_OK;
}

/* copy both populations to the device in the layout the caller has them */
struct DevGenotypes { const double *a64 = nullptr, *b64 = nullptr; const signed char *a8 = nullptr, *b8 = nullptr; };

/* The genotype upload is cut into SNP chunks on a copy stream of its own, one event per chunk, so that the per-SNP
   kernels (counting / packing) and every window whose SNPs have all arrived can run while later chunks are still on the
   bus. Chunk boundaries are multiples of 4096 SNPs (whole bit-plane words, whole staging tiles). */
#define FPT_MAX_CHUNKS 8
#define FPT_CHUNK_BYTES_FET ((size_t)16 << 20)    /* the copy dominates a FET scan: many chunks, windows released early */
#define FPT_CHUNK_BYTES_CSS ((size_t)96 << 20)    /* compute dominates a CSS scan: sub-range launches cost more in tails than the
                                                     copy they hide until the input is several hundred MB */
/* A few parked helper threads: starting a std::thread costs ~30 us, and a dozen of them started one after the other cost more
   than the gather they were meant to speed up (measured: 0.4 ms per chromosome either way). One job at a time (the mutex). */
class HelperPool {
    std::mutex job_mu, mu;
    std::condition_variable cv, done_cv;
    std::vector<std::thread> th;
    std::function<void(int)> fn;
    int nt = 0, gen = 0, left = 0;
    bool stop = false;
    void loop(int id) {
        int seen = 0;
        for (;;) {
            std::function<void(int)> f;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || gen != seen; });
                if (stop) return;
                seen = gen;
                if (id >= nt) continue;
                f = fn;
            }
            f(id);
            { std::lock_guard<std::mutex> lk(mu); if (--left == 0) done_cv.notify_all(); }
        }
    }
public:
    explicit HelperPool(int n) { for (int i = 0; i < n; i++) th.emplace_back(&HelperPool::loop, this, i + 1); }
    ~HelperPool() { { std::lock_guard<std::mutex> lk(mu); stop = true; } cv.notify_all(); for (auto &x : th) x.join(); }
    int helpers() const { return (int)th.size(); }
    /* f(0 .. parts-1), part 0 on the caller; parts - 1 <= helpers() */
    void run(int parts, const std::function<void(int)> &f) {
        std::lock_guard<std::mutex> job(job_mu);
        if (parts > 1) {
            std::lock_guard<std::mutex> lk(mu);
            fn = f; nt = parts; left = parts - 1; gen++;
        }
        if (parts > 1) cv.notify_all();
        f(0);
        if (parts > 1) { std::unique_lock<std::mutex> lk(mu); done_cv.wait(lk, [&] { return left == 0; }); }
    }
};
static HelperPool &helper_pool() {
    static HelperPool pool((int)std::max(1u, std::min(11u, std::thread::hardware_concurrency() ? std::thread::hardware_concurrency() - 1 : 1u)));
    return pool;
}

struct UploadPlan {
    int n = 0;                                   /* chunks */
    int queued = 0;                              /* chunks already handed to the copy stream */
    long long snp_end[FPT_MAX_CHUNKS];           /* exclusive end of chunk c */
    cudaEvent_t ev[FPT_MAX_CHUNKS];
    cudaStream_t cs = nullptr;
    char *da = nullptr, *db = nullptr;           /* device arrays */
    const char *ha = nullptr, *hb = nullptr;     /* host arrays */
    char *sa = nullptr, *sb = nullptr;           /* pageable host arrays: page-locked staging copies (non-null = staged upload) */
    size_t rowa = 0, rowb = 0;                   /* bytes per SNP */
    ~UploadPlan() {
        if (cs) cudaStreamSynchronize(cs);       /* error paths: no copy may outlive the buffers it writes */
        for (int c = 0; c < n; c++) cudaEventDestroy(ev[c]);
        if (cs) cudaStreamDestroy(cs);
    }
};

/* hand chunks [queued, upto) to the copy stream */
static int upload_enqueue(UploadPlan *p, int upto) {
    for (; p->queued < upto && p->queued < p->n; p->queued++) {
        const long long s0 = p->queued ? p->snp_end[p->queued - 1] : 0, s1 = p->snp_end[p->queued];
        const char *srca = p->ha, *srcb = p->hb;
        if (p->sa) {
            /* pageable caller arrays (what numpy hands over): a cudaMemcpyAsync from them is staged by the driver through one
               thread at ~10 GB/s and blocks the caller; copying the chunk into our own page-locked buffer with the helper threads
               runs at several times that and leaves the DMA asynchronous */
            const size_t ba = (size_t)(s1 - s0) * p->rowa, bb = (size_t)(s1 - s0) * p->rowb, oa = (size_t)s0 * p->rowa, ob = (size_t)s0 * p->rowb;
            HelperPool &pool = helper_pool();
            const int parts = (int)std::max<size_t>(1, std::min<size_t>((size_t)pool.helpers() + 1, (ba + bb) >> 20));
            char *sa = p->sa, *sb = p->sb;
            const char *ha = p->ha, *hb = p->hb;
            pool.run(parts, [=](int t) {
                const size_t a0 = ba * t / parts, a1 = ba * (t + 1) / parts, b0 = bb * t / parts, b1 = bb * (t + 1) / parts;
                memcpy(sa + oa + a0, ha + oa + a0, a1 - a0);
                memcpy(sb + ob + b0, hb + ob + b0, b1 - b0);
            });
            srca = p->sa; srcb = p->sb;
        }
        CU(cudaMemcpyAsync(p->da + (size_t)s0 * p->rowa, srca + (size_t)s0 * p->rowa, (size_t)(s1 - s0) * p->rowa, cudaMemcpyHostToDevice, p->cs));
        CU(cudaMemcpyAsync(p->db + (size_t)s0 * p->rowb, srcb + (size_t)s0 * p->rowb, (size_t)(s1 - s0) * p->rowb, cudaMemcpyHostToDevice, p->cs));
        CU(cudaEventRecord(p->ev[p->queued], p->cs));
    }
    return FPT_OK;
}

/* Allocates the device arrays, lays out the chunks and starts the first quarter of them. The caller enqueues whatever small
   host->device copies the scan needs early (positions) and then calls upload_enqueue(plan, plan->n): copies are served in
   enqueue order by the DMA engine, so anything enqueued after the last chunk would wait for the whole genome. */
static int upload_genotypes(Arena &ar, const fpt_genotypes *g, DevGenotypes *d, UploadPlan *plan, size_t chunk_bytes) {
    const size_t na = (size_t)g->nsnp * g->asize, nb = (size_t)g->nsnp * g->bsize;
    const size_t esz = g->avals ? sizeof(double) : 1;
    if (g->avals) { double *pa, *pb; CHECK(ar.get(&pa, na)); CHECK(ar.get(&pb, nb)); plan->da = (char *)pa; plan->db = (char *)pb; d->a64 = pa; d->b64 = pb; }
    else { signed char *pa, *pb; CHECK(ar.get(&pa, na)); CHECK(ar.get(&pb, nb)); plan->da = (char *)pa; plan->db = (char *)pb; d->a8 = pa; d->b8 = pb; }
    plan->ha = g->avals ? (const char *)g->avals : (const char *)g->acodes;
    plan->hb = g->avals ? (const char *)g->bvals : (const char *)g->bcodes;
    plan->rowa = (size_t)g->asize * esz; plan->rowb = (size_t)g->bsize * esz;
    if ((na + nb) * esz >= ((size_t)1 << 20)) {
        cudaPointerAttributes at;
        const bool pageable = cudaPointerGetAttributes(&at, plan->ha) == cudaSuccess && at.type == cudaMemoryTypeUnregistered;
        cudaGetLastError();
        DeviceCtx *c;
        if (pageable && get_ctx(&c) == FPT_OK) {
            void *s4 = nullptr, *s5 = nullptr;
            if (pinned_slot(c, 4, na * esz, &s4) == FPT_OK && pinned_slot(c, 5, nb * esz, &s5) == FPT_OK) { plan->sa = (char *)s4; plan->sb = (char *)s5; }
        }
    }
    /* one chunk per `chunk_bytes` (FPT_UPLOAD_CHUNK_BYTES overrides: tests use it to cut small inputs), at most FPT_MAX_CHUNKS */
    const size_t bytes = (na + nb) * esz;
    if (const char *e = getenv("FPT_UPLOAD_CHUNK_BYTES")) { const long long v = atoll(e); if (v > 0) chunk_bytes = (size_t)v; }
    if (plan->sa) chunk_bytes = std::min(chunk_bytes, (size_t)16 << 20);    /* staged: the host copy of chunk k+1 overlaps the DMA of chunk k */
    const int want = (int)std::min<size_t>(FPT_MAX_CHUNKS, std::max<size_t>(1, bytes / chunk_bytes));
    long long per = ((g->nsnp + want - 1) / want + 4095) & ~4095LL;
    if (per <= 0) per = 4096;
    /* a single chunk of tens of MB still leaves the GPU idle for a millisecond: lead with an eighth of it, then a quarter, then
       the rest. Where the compute per SNP outlasts its upload (the CSS scan: ~6 ms of kernels for ~1.3 ms of DMA per chromosome)
       each piece arrives while the windows of the one before are being scored, and only the first eighth's upload is exposed;
       with two pieces (1/8, 7/8) the GPU finished the first before the second had arrived. */
    long long first = per, second = per;
    if (want == 1 && bytes >= ((size_t)32 << 20)) {
        first = std::max<long long>(4096, (g->nsnp / 8 + 4095) & ~4095LL);
        second = std::max<long long>(4096, (g->nsnp / 4 + 4095) & ~4095LL);
    }
    CU(cudaStreamCreateWithFlags(&plan->cs, cudaStreamNonBlocking));
    cudaEvent_t ready;                           /* the copy stream starts once the stream-ordered allocations exist */
    CU(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    CU(cudaEventRecord(ready, ar.st));
    CU(cudaStreamWaitEvent(plan->cs, ready, 0));
    CU(cudaEventDestroy(ready));
    for (long long s0 = 0; s0 < g->nsnp && plan->n < FPT_MAX_CHUNKS;) {
        const long long s1 = (plan->n == FPT_MAX_CHUNKS - 1) ? g->nsnp : std::min<long long>(g->nsnp, s0 + (plan->n == 0 ? first : (plan->n == 1 ? second : per)));
        CU(cudaEventCreateWithFlags(&plan->ev[plan->n], cudaEventDisableTiming));
        plan->snp_end[plan->n++] = s1;
        s0 = s1;
    }
    return upload_enqueue(plan, (plan->n + 3) / 4);       /* a quarter now: the DMA stays busy while the host prepares the rest */
}

/* positions ascending? (the windows-as-they-arrive schedule needs it; unsorted input falls back to one pass at the end) */
static bool positions_sorted(const int32_t *pos, long long n) {
    for (long long k = 1; k < n; k++) if (pos[k] < pos[k - 1]) return false;
    return true;
}

/* windows [0, returned) of the range have every SNP among the first `s1` ones (positions ascending): window gw covers
   positions <= gw*wstep + wsize (fpt_window_table_kernel), so it is complete iff that is below pos[s1] */
static long long windows_complete(const fpt_genotypes *g, const fpt_scan_range *r, long long nwin, long long s1) {
    if (s1 >= g->nsnp) return nwin;
    const long long lim = (long long)g->pos[s1] - r->wsize - 1;
    if (lim < 0) return 0;
    const long long gw_excl = lim / r->wstep + 1;
    return std::max<long long>(0, std::min<long long>(nwin, gw_excl - r->window_begin));
}

struct HostStream {
    cudaStream_t st = nullptr;
    int make() { CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking)); return FPT_OK; }
    ~HostStream() { if (st) cudaStreamDestroy(st); }
};

extern "C" int fpt_fet_per_snp(const fpt_genotypes *g, int32_t *tables, double *neglog10p) {
    DeviceCtx *c;
    CHECK(get_ctx(&c));
    CHECK(check_genotypes(g));
    if (g->nsnp == 0) return FPT_OK;
    std::lock_guard<std::mutex> host_lock(g_host_mu[c->device]);   /* the staging buffers of a pageable upload are per device */
    HostStream hs;
    CHECK(hs.make());
    Arena ar(hs.st);
    DevGenotypes d;
    UploadPlan plan;
    CHECK(upload_genotypes(ar, g, &d, &plan, FPT_CHUNK_BYTES_FET));
    CHECK(upload_enqueue(&plan, plan.n));
    CU(cudaStreamWaitEvent(hs.st, plan.ev[plan.n - 1], 0));
    int32_t *d_tab; double *d_sc;
    CHECK(ar.get(&d_tab, (size_t)g->nsnp * 4));
    CHECK(ar.get(&d_sc, (size_t)g->nsnp));
    if (d.a64) CHECK(fpt_dev_fet_count_f64(d.a64, d.b64, g->nsnp, g->asize, g->bsize, d_tab, hs.st));
    else CHECK(fpt_dev_fet_count_i8((const int8_t *)d.a8, (const int8_t *)d.b8, g->nsnp, g->asize, g->bsize, d_tab, hs.st));
    if (neglog10p) {
        CHECK(fpt_dev_fet_score(d_tab, g->nsnp, g->asize + g->bsize, 0, d_sc, hs.st));
        CU(cudaMemcpyAsync(neglog10p, d_sc, (size_t)g->nsnp * sizeof(double), cudaMemcpyDeviceToHost, hs.st));
    }
    if (tables) CU(cudaMemcpyAsync(tables, d_tab, (size_t)g->nsnp * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, hs.st));
    CU(cudaStreamSynchronize(hs.st));
    return FPT_OK;
}

extern "C" int fpt_fet_tables(const int32_t *tables, int64_t n, int force_log, double *neglog10p) {
    DeviceCtx *c;
    CHECK(get_ctx(&c));
    if (n < 0 || (n > 0 && (!tables || !neglog10p))) return fail(FPT_ERR_ARG, "fet_tables: bad arguments");
    if (n == 0) return FPT_OK;
    HostStream hs;
    CHECK(hs.make());
    Arena ar(hs.st);
    int32_t *d_tab; double *d_sc;
    CHECK(ar.get(&d_tab, (size_t)n * 4));
    CHECK(ar.get(&d_sc, (size_t)n));
    CU(cudaMemcpyAsync(d_tab, tables, (size_t)n * 4 * sizeof(int32_t), cudaMemcpyHostToDevice, hs.st));
    CHECK(fpt_dev_fet_score(d_tab, n, 0, force_log, d_sc, hs.st));
    CU(cudaMemcpyAsync(neglog10p, d_sc, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, hs.st));
    CU(cudaStreamSynchronize(hs.st));
    return FPT_OK;
}

/* shared front half of both scans: positions to the device, window table, largest window */
struct ScanFront {
    int32_t *d_pos = nullptr, *d_wl = nullptr, *d_wr = nullptr, *d_max = nullptr;
    long long nwin = 0;
    int max_npos = 0;
};

static int scan_front(Arena &ar, const fpt_genotypes *g, const fpt_scan_range *r, ScanFront *f) {
    CHECK(check_range(r, &f->nwin));
    if (f->nwin > 0x7fffffffLL) return fail(FPT_ERR_ARG, "more than 2^31-1 windows in one scan");
    CHECK(ar.get(&f->d_pos, (size_t)g->nsnp));
    CHECK(ar.get(&f->d_wl, (size_t)f->nwin));
    CHECK(ar.get(&f->d_wr, (size_t)f->nwin));
    CHECK(ar.get(&f->d_max, 1));
    CU(cudaMemcpyAsync(f->d_pos, g->pos, (size_t)g->nsnp * sizeof(int32_t), cudaMemcpyHostToDevice, ar.st));
    CU(cudaMemsetAsync(f->d_max, 0, sizeof(int32_t), ar.st));
    CHECK(fpt_dev_window_table(f->d_pos, g->nsnp, r, f->d_wl, f->d_wr, f->d_max, ar.st));
    return FPT_OK;
}

/* The genotype upload is in flight on plan.cs. Per chunk: wait for it, count + score its SNPs, then run every window whose
   SNPs have all arrived (positions ascending; otherwise all windows after the last chunk). */
static int fet_scan_core(DeviceCtx *c, Arena &ar, const fpt_genotypes *g, const DevGenotypes &d, UploadPlan &plan,
                         const fpt_scan_range *r, double perc, double *scores, double *stddev, uint8_t *written) {
    cudaStream_t st = ar.st;
    long long nwin;
    CHECK(check_range(r, &nwin));
    ScanFront f;
    CHECK(scan_front(ar, g, r, &f));
    int32_t *d_tab = nullptr; double *d_snp = nullptr, *d_sc = nullptr, *d_sd = nullptr; uint8_t *d_fl = nullptr;
    uint64_t *d_states = nullptr;
    CHECK(ar.get(&d_tab, (size_t)g->nsnp * 4));
    CHECK(ar.get(&d_snp, (size_t)g->nsnp));
    CHECK(ar.get(&d_sc, (size_t)nwin)); CHECK(ar.get(&d_sd, (size_t)nwin)); CHECK(ar.get(&d_fl, (size_t)nwin));
    if (r->states_resample) {
        CHECK(ar.get(&d_states, (size_t)nwin));
        CU(cudaMemcpyAsync(d_states, r->states_resample, (size_t)nwin * 8, cudaMemcpyHostToDevice, st));
    }
    /* the small host->device copies above (positions, stream states) are now ahead of the remaining chunks in the DMA queue */
    if (!plan.sa) CHECK(upload_enqueue(&plan, plan.n));         /* staged uploads copy on the host: one chunk ahead of the launches, below */
    CU(cudaMemsetAsync(d_fl, 0, (size_t)nwin, st));
    CU(cudaMemcpyAsync(&f.max_npos, f.d_max, sizeof(int), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));                          /* positions and window table only: the genotypes are on plan.cs */
    const bool stream_windows = plan.n > 1 && positions_sorted(g->pos, g->nsnp);
    long long wdone = 0, s0 = 0;
    for (int k = 0; k < plan.n; k++) {
        const long long s1 = plan.snp_end[k];
        if (plan.sa) CHECK(upload_enqueue(&plan, k + 2));
        CU(cudaStreamWaitEvent(st, plan.ev[k], 0));
        if (d.a64) CHECK(fpt_dev_fet_count_f64(d.a64 + (size_t)s0 * g->asize, d.b64 + (size_t)s0 * g->bsize, s1 - s0, g->asize, g->bsize, d_tab + 4 * s0, st));
        else CHECK(fpt_dev_fet_count_i8((const int8_t *)d.a8 + (size_t)s0 * g->asize, (const int8_t *)d.b8 + (size_t)s0 * g->bsize, s1 - s0, g->asize,
                                        g->bsize, d_tab + 4 * s0, st));
        CHECK(fpt_dev_fet_score(d_tab + 4 * s0, s1 - s0, g->asize + g->bsize, 0, d_snp + s0, st));
        const long long wav = (k == plan.n - 1) ? nwin : (stream_windows ? windows_complete(g, r, nwin, s1) : 0);
        if (wav > wdone) {
            fpt_scan_range sub = *r;
            sub.window_begin = r->window_begin + wdone;
            sub.window_end = r->window_begin + wav;
            CHECK(fpt_dev_fet_windows(d_snp, f.d_wl + wdone, f.d_wr + wdone, &sub, f.max_npos, perc, d_states ? d_states + wdone : nullptr,
                                      d_sc + wdone, d_sd + wdone, d_fl + wdone, st));
            wdone = wav;
        }
        s0 = s1;
    }
    void *h0, *h1, *h2;
    CHECK(pinned_slot(c, 0, (size_t)nwin * 8, &h0)); CHECK(pinned_slot(c, 1, (size_t)nwin * 8, &h1));
    CHECK(pinned_slot(c, 2, (size_t)nwin, &h2));
    double *h_sc = (double *)h0, *h_sd = (double *)h1; uint8_t *h_fl = (uint8_t *)h2;
    CU(cudaMemcpyAsync(h_sc, d_sc, (size_t)nwin * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_sd, d_sd, (size_t)nwin * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_fl, d_fl, (size_t)nwin, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (long long w = 0; w < nwin; w++) {                  /* un-scored windows stay as the caller left them */
        if (h_fl[w]) { scores[w] = h_sc[w]; stddev[w] = h_sd[w]; }
        if (written) written[w] = h_fl[w];
    }
    return FPT_OK;
}

extern "C" int fpt_fet_scan(const fpt_genotypes *g, const fpt_scan_range *r, double perc, double *scores,
                            double *stddev, uint8_t *written) {
    DeviceCtx *c;
    CHECK(get_ctx(&c));
    CHECK(check_genotypes(g));
    long long nwin;
    CHECK(check_range(r, &nwin));
    if (nwin == 0 || g->nsnp == 0) return FPT_OK;
    if (!g->pos || !scores || !stddev) return fail(FPT_ERR_ARG, "fet_scan: positions and both outputs are required");
    std::lock_guard<std::mutex> host_lock(g_host_mu[c->device]);
    HostStream hs;
    CHECK(hs.make());
    Arena ar(hs.st);
    DevGenotypes d;
    UploadPlan plan;
    CHECK(upload_genotypes(ar, g, &d, &plan, FPT_CHUNK_BYTES_FET));
    return fet_scan_core(c, ar, g, d, plan, r, perc, scores, stddev, written);
}

static int css_scan_core(DeviceCtx *c, Arena &ar, const fpt_genotypes *g, const DevGenotypes &d, UploadPlan &plan,
                         const fpt_scan_range *r, int treshold, int runs, int drosophila, int mds, double *scores, double *p,
                         uint8_t *written, const fpt_css_probes *probes) {
    cudaStream_t st = ar.st;
    long long nwin;
    CHECK(check_range(r, &nwin));
    const int m = g->asize + g->bsize;
    const Knobs kn = knobs_now();                 /* one snapshot for every launch of this call */
    ScanFront f;
    CHECK(scan_front(ar, g, r, &f));
    uint32_t *d_planes = nullptr; double *d_abs = nullptr;
    if (drosophila) CHECK(ar.get(&d_abs, (size_t)g->nsnp));
    else CHECK(ar.get(&d_planes, fpt_dev_css_planes_bytes(g->nsnp, m) / 4));
    const int nruns = mds == 1 ? 4 : (mds == 2 ? 1 : 0);
    size_t ws_bytes = fpt_dev_css_workspace_bytes(m, nwin, mds);
    unsigned char *d_ws; double *d_sc, *d_p; uint8_t *d_st;
    CHECK(ar.get(&d_ws, ws_bytes));
    CHECK(ar.get(&d_sc, (size_t)nwin)); CHECK(ar.get(&d_p, (size_t)nwin)); CHECK(ar.get(&d_st, (size_t)nwin));
    CU(cudaMemsetAsync(d_st, 0, (size_t)nwin, st));
    uint64_t *d_s0 = nullptr, *d_s1 = nullptr;
    if (r->states_resample) { CHECK(ar.get(&d_s0, (size_t)nwin)); CU(cudaMemcpyAsync(d_s0, r->states_resample, (size_t)nwin * 8, cudaMemcpyHostToDevice, st)); }
    if (r->states_init) { CHECK(ar.get(&d_s1, (size_t)nwin)); CU(cudaMemcpyAsync(d_s1, r->states_init, (size_t)nwin * 8, cudaMemcpyHostToDevice, st)); }
    fpt_css_probes dp;
    memset(&dp, 0, sizeof dp);
    if (probes) {
        if (probes->X) CHECK(ar.get(&dp.X, (size_t)nwin * 2 * m));
        if (probes->evals) CHECK(ar.get(&dp.evals, (size_t)nwin * 3));
        if (probes->hits) CHECK(ar.get(&dp.hits, (size_t)nwin));
        if (probes->nperm) CHECK(ar.get(&dp.nperm, (size_t)nwin));
        if (probes->smacof_iters && nruns) CHECK(ar.get(&dp.smacof_iters, (size_t)nwin * nruns));
        if (probes->smacof_sigma && nruns) CHECK(ar.get(&dp.smacof_sigma, (size_t)nwin * nruns));
        if (dp.hits) CU(cudaMemsetAsync(dp.hits, 0, (size_t)nwin * 4, st));
        if (dp.nperm) CU(cudaMemsetAsync(dp.nperm, 0, (size_t)nwin * 4, st));
        if (dp.X) CU(cudaMemsetAsync(dp.X, 0, (size_t)nwin * 2 * m * 8, st));
    }
    /* the small host->device copies above (positions, stream states) are now ahead of the remaining chunks in the DMA queue */
    if (!plan.sa) CHECK(upload_enqueue(&plan, plan.n));         /* staged uploads: one chunk ahead of the launches, in the loop */
    /* per upload chunk: pack its SNPs, then score every window whose SNPs have all arrived (see fet_scan_core). Only where
       a sub-range still fills the GPU many times over: the CTA-per-window kernels of large cohorts run a few hundred windows
       at a time, and a short extra launch costs them a whole wave. */
    const bool stream_windows = plan.n > 1 && nwin >= 8192 && css_plan(c, m, kn).mds_warps > 0 && positions_sorted(g->pos, g->nsnp);
    long long wdone = 0, s0 = 0;
    for (int k = 0; k < plan.n; k++) {
        const long long s1 = plan.snp_end[k];
        if (plan.sa) CHECK(upload_enqueue(&plan, k + 2));
        CU(cudaStreamWaitEvent(st, plan.ev[k], 0));
        if (drosophila) CHECK(fpt_dev_css_absdiff(d.a64 + s0, d.b64 + s0, s1 - s0, d_abs + s0, st));
        else {
            uint32_t *pl = d_planes + (size_t)(s0 / 32) * 2 * m;        /* chunk starts are multiples of 32 SNPs */
            if (d.a64) CHECK(fpt_dev_css_pack_f64(d.a64 + (size_t)s0 * g->asize, d.b64 + (size_t)s0 * g->bsize, s1 - s0, g->asize, g->bsize, pl, st));
            else CHECK(fpt_dev_css_pack_i8((const int8_t *)d.a8 + (size_t)s0 * g->asize, (const int8_t *)d.b8 + (size_t)s0 * g->bsize, s1 - s0, g->asize,
                                           g->bsize, pl, st));
        }
        const long long wav = (k == plan.n - 1) ? nwin : (stream_windows ? windows_complete(g, r, nwin, s1) : 0);
        if (wav > wdone) {
            const long long nw = wav - wdone;
            fpt_scan_range sub = *r;
            sub.window_begin = r->window_begin + wdone;
            sub.window_end = r->window_begin + wav;
            sub.states_resample = d_s0 ? d_s0 + wdone : nullptr;
            sub.states_init = d_s1 ? d_s1 + wdone : nullptr;
            fpt_css_probes dq = dp;
            if (dq.X) dq.X += (size_t)wdone * 2 * m;
            if (dq.evals) dq.evals += (size_t)wdone * 3;
            if (dq.hits) dq.hits += wdone;
            if (dq.nperm) dq.nperm += wdone;
            if (dq.smacof_iters) dq.smacof_iters += (size_t)wdone * nruns;
            if (dq.smacof_sigma) dq.smacof_sigma += (size_t)wdone * nruns;
            CHECK(dev_css_windows(kn, d_planes, d_abs, g->asize, g->bsize, f.d_wl + wdone, f.d_wr + wdone, &sub, treshold, runs, mds, d_ws,
                                      fpt_dev_css_workspace_bytes(m, nw, mds), d_sc + wdone, d_p + wdone, d_st + wdone, probes ? &dq : nullptr, st));
            wdone = wav;
        }
        s0 = s1;
    }
    double *h_sc, *h_p; uint8_t *h_st;
    void *h0, *h1, *h2;
    CHECK(pinned_slot(c, 0, (size_t)nwin * 8, &h0)); CHECK(pinned_slot(c, 1, (size_t)nwin * 8, &h1));
    CHECK(pinned_slot(c, 2, (size_t)nwin, &h2));
    h_sc = (double *)h0; h_p = (double *)h1; h_st = (uint8_t *)h2;
    CU(cudaMemcpyAsync(h_sc, d_sc, (size_t)nwin * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_p, d_p, (size_t)nwin * 8, cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(h_st, d_st, (size_t)nwin, cudaMemcpyDeviceToHost, st));
    if (probes) {
        if (probes->X) CU(cudaMemcpyAsync(probes->X, dp.X, (size_t)nwin * 2 * m * 8, cudaMemcpyDeviceToHost, st));
        if (probes->evals) CU(cudaMemcpyAsync(probes->evals, dp.evals, (size_t)nwin * 3 * 8, cudaMemcpyDeviceToHost, st));
        if (probes->hits) CU(cudaMemcpyAsync(probes->hits, dp.hits, (size_t)nwin * 4, cudaMemcpyDeviceToHost, st));
        if (probes->nperm) CU(cudaMemcpyAsync(probes->nperm, dp.nperm, (size_t)nwin * 4, cudaMemcpyDeviceToHost, st));
        if (probes->smacof_iters && nruns) CU(cudaMemcpyAsync(probes->smacof_iters, dp.smacof_iters, (size_t)nwin * nruns * 4, cudaMemcpyDeviceToHost, st));
        if (probes->smacof_sigma && nruns) CU(cudaMemcpyAsync(probes->smacof_sigma, dp.smacof_sigma, (size_t)nwin * nruns * 8, cudaMemcpyDeviceToHost, st));
    }
    CU(cudaStreamSynchronize(st));
    for (long long w = 0; w < nwin; w++) {
        /* css.c:126-132: score and p are stored only when the scorer did not return -1. -1 is the reference's
           "window discarded" sentinel, so a window whose score happens to BE exactly -1.0 is dropped too. */
        const bool store = h_st[w] == FPT_WIN_SCORED && h_sc[w] != -1.0;
        if (store) { scores[w] = h_sc[w]; p[w] = h_p[w]; }
        if (written) written[w] = store;
        if (probes && probes->status) probes->status[w] = h_st[w];
    }
    return FPT_OK;
}


extern "C" int fpt_css_scan(const fpt_genotypes *g, const fpt_scan_range *r, int treshold, int runs, int drosophila,
                            int mds, double *scores, double *p, uint8_t *written, const fpt_css_probes *probes) {
    DeviceCtx *c;
    CHECK(get_ctx(&c));
    if (!g) return fail(FPT_ERR_ARG, "genotypes are NULL");
    if (drosophila) {
        if (!g->avals || !g->bvals || g->asize != 1 || g->bsize != 1)
            return fail(FPT_ERR_ARG, "the frequency metric takes float64 frequencies, one track per population");
        if (g->nsnp < 0 || g->nsnp > 0x7fffffffLL) return fail(FPT_ERR_ARG, "bad nsnp");
    } else {
        CHECK(check_genotypes(g));
    }
    long long nwin;
    CHECK(check_range(r, &nwin));
    if (mds < 0 || mds > 2) return fail(FPT_ERR_ARG, "css: mds must be 0, 1 or 2 (got %d)", mds);
    if (nwin == 0 || g->nsnp == 0) return FPT_OK;
    if (!g->pos || !scores || !p) return fail(FPT_ERR_ARG, "css_scan: positions and both outputs are required");
    std::lock_guard<std::mutex> host_lock(g_host_mu[c->device]);
    HostStream hs;
    CHECK(hs.make());
    Arena ar(hs.st);
    DevGenotypes d;
    UploadPlan plan;
    CHECK(upload_genotypes(ar, g, &d, &plan, FPT_CHUNK_BYTES_CSS));
    return css_scan_core(c, ar, g, d, plan, r, treshold, runs, drosophila, mds, scores, p, written, probes);
}

/* parity probe: the count codes of ONE window (global index `window`) as the chosen K4 kernel writes them, expanded to int32
   counts[m * m] on the host (mode 2 tcgen05 GEMM, 1 popcounts) */
extern "C" int fpt_debug_k4_counts(const fpt_genotypes *g, const fpt_scan_range *r, int mode, int64_t window, int32_t *counts) {
    DeviceCtx *c;
    CHECK(get_ctx(&c));
    CHECK(check_genotypes(g));
    long long nwin;
    CHECK(check_range(r, &nwin));
    if (!counts || !g->pos || window < r->window_begin || window >= r->window_end) return fail(FPT_ERR_ARG, "k4 probe: bad arguments");
    if (mode != 1 && mode != 2) return fail(FPT_ERR_ARG, "k4 probe: mode must be 1 (popcounts) or 2 (tcgen05)");
    const int m = g->asize + g->bsize;
    std::lock_guard<std::mutex> host_lock(g_host_mu[c->device]);
    HostStream hs;
    CHECK(hs.make());
    Arena ar(hs.st);
    DevGenotypes d;
    UploadPlan plan;
    CHECK(upload_genotypes(ar, g, &d, &plan, FPT_CHUNK_BYTES_CSS));
    ScanFront f;
    CHECK(scan_front(ar, g, r, &f));
    CHECK(upload_enqueue(&plan, plan.n));
    CU(cudaStreamWaitEvent(hs.st, plan.ev[plan.n - 1], 0));
    uint32_t *d_planes;
    CHECK(ar.get(&d_planes, fpt_dev_css_planes_bytes(g->nsnp, m) / 4));
    if (d.a64) CHECK(fpt_dev_css_pack_f64(d.a64, d.b64, g->nsnp, g->asize, g->bsize, d_planes, hs.st));
    else CHECK(fpt_dev_css_pack_i8((const int8_t *)d.a8, (const int8_t *)d.b8, g->nsnp, g->asize, g->bsize, d_planes, hs.st));
    const size_t stride = fpt_k4_window_stride(m);
    unsigned char *d_codes;
    CHECK(ar.get(&d_codes, stride));
    CU(cudaMemsetAsync(d_codes, 0xEE, stride, hs.st));
    const long long wi = window - r->window_begin;
    if (mode == 2) {
        CU(cudaFuncSetAttribute(fpt_css_k4_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FPT_K4_SMEM));
        fpt_css_k4_umma_kernel<<<1, FPT_K4_THREADS, FPT_K4_SMEM, hs.st>>>(d_planes, m, f.d_wl + wi, f.d_wr + wi, 1, d_codes, stride);
    } else {
        const size_t smem4 = fpt_k4_popc_smem(m);
        if (smem4 > 48 * 1024) CU(cudaFuncSetAttribute(fpt_css_k4_popc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem4));
        fpt_css_k4_popc_kernel<<<1, 256, smem4, hs.st>>>(d_planes, m, f.d_wl + wi, f.d_wr + wi, 1, d_codes, stride);
    }
    CU(cudaGetLastError());
    std::vector<unsigned char> h(stride);
    int32_t lr[2];
    CU(cudaMemcpyAsync(h.data(), d_codes, stride, cudaMemcpyDeviceToHost, hs.st));
    CU(cudaMemcpyAsync(&lr[0], f.d_wl + wi, 4, cudaMemcpyDeviceToHost, hs.st));
    CU(cudaMemcpyAsync(&lr[1], f.d_wr + wi, 4, cudaMemcpyDeviceToHost, hs.st));
    CU(cudaStreamSynchronize(hs.st));
    const int ld = fpt_k4_ld(m), esz = fpt_k4_code_bytes(lr[1] - lr[0]);
    for (int i = 0; i < m; i++)
        for (int j = 0; j < m; j++)
            counts[(size_t)i * m + j] = lr[1] <= lr[0] ? -1 : (esz == 1 ? (int32_t)h[(size_t)i * ld + j] : (int32_t)((const unsigned short *)h.data())[(size_t)i * ld + j]);
    return FPT_OK;
}

/* ------------------------------------------------------------------------------------------------ drop-ins */
/* comparative.c:25-34 get_population_size, bounded by the array length (the reference runs off the end
   when every position is equal) */
static int population_size(const int *pos, int len) {
    int n = 0;
    while (n < len && pos[n] == pos[0]) n++;
    return n;
}

/* reference layout -> one position per SNP (into pinned staging). The reads are strided (one int per asize individuals), i.e. one
   cache line per SNP: on a genome-sized call this is ~64 bytes of host memory traffic per SNP and population, so it is split over
   a few threads. Only population A's positions are gathered on the critical path (nothing can be launched before the window
   table has them); that B lists the same SNPs (the reference silently mis-pairs them otherwise) is verified by
   PositionCheck below while the GPU is already scoring. */
static void gather_positions(const int *apos, int asize, long long nsnp, int32_t *pos) {
    HelperPool &pool = helper_pool();
    const int nt = (int)std::max<long long>(1, std::min<long long>(pool.helpers() + 1, nsnp / 16384));
    pool.run(nt, [=](int t) {
        const long long lo = nsnp * t / nt, hi = nsnp * (t + 1) / nt;
        for (long long k = lo; k < hi; k++) pos[k] = apos[k * asize];
    });
}

/* bpos[k * bsize] == pos[k] for every SNP, checked on helper threads; result() joins them */
struct PositionCheck {
    std::vector<std::thread> th;
    std::vector<long long> bad;
    void start(const int32_t *pos, const int *bpos, int bsize, long long nsnp) {
        const unsigned hw = std::thread::hardware_concurrency();
        const int nt = (int)std::max<long long>(1, std::min<long long>(std::min<long long>(2, hw ? hw : 1), nsnp / 1000000));   /* it has the whole scan to finish */
        bad.assign((size_t)nt, -1);
        for (int t = 0; t < nt; t++)
            th.emplace_back([this, pos, bpos, bsize, nsnp, nt, t]() {
                const long long lo = nsnp * t / nt, hi = nsnp * (t + 1) / nt;
                for (long long k = lo; k < hi; k++)
                    if (pos[k] != bpos[k * bsize]) { bad[(size_t)t] = k; return; }
            });
    }
    long long result() {                                   /* first mismatching SNP, or -1 */
        for (auto &x : th) x.join();
        th.clear();
        for (long long k : bad) if (k >= 0) return k;
        return -1;
    }
    ~PositionCheck() { for (auto &x : th) if (x.joinable()) x.join(); }
};

static fpt_scan_range full_range(int regend, int wsize, int wstep, int semantics) {
    fpt_scan_range r;
    memset(&r, 0, sizeof r);
    r.regend = regend; r.wsize = wsize; r.wstep = wstep; r.semantics = semantics;
    r.window_begin = 0;
    r.window_end = wstep > 0 ? regend / wstep : 0;       /* length of the caller's output arrays (Q17) */
    r.seed = g_seed.load();
    return r;
}

/* Shared by the four drop-ins: population sizes from the run length of the first position (get_population_size,
   comparative.c:25-34), the float64 genotype upload enqueued FIRST so that the host-side gather of one position per
   SNP (and the A/B consistency check) overlaps the DMA, then the scan proper. */
static int dropin(int css, double *avals, double *bvals, int *apos, int *bpos, int regend, int wsize, int wstep, int alen,
                  int blen, double perc, int treshold, int runs, int drosophila, int mds, double *out0, double *out1,
                  int semantics) {
    DeviceCtx *c;
    CHECK(get_ctx(&c));
    if (wsize <= 0 || wstep <= 0) return fail(FPT_ERR_ARG, "bad window geometry wsize=%d wstep=%d", wsize, wstep);
    if (alen <= 0 || blen <= 0 || !apos || !bpos || !avals || !bvals || !out0 || !out1) return fail(FPT_ERR_ARG, "empty or NULL arrays");
    fpt_genotypes g;
    memset(&g, 0, sizeof g);
    g.asize = population_size(apos, alen);
    g.bsize = population_size(bpos, blen);
    const long long na = alen / g.asize, nb = blen / g.bsize;
    if (na != nb)
        return fail(FPT_ERR_POSITIONS, "population A has %lld SNPs (%d individuals), B has %lld (%d individuals)", na, g.asize, nb, g.bsize);
    if (css && drosophila && (g.asize != 1 || g.bsize != 1))
        return fail(FPT_ERR_ARG, "the frequency metric takes one track per population (got %d+%d)", g.asize, g.bsize);
    g.avals = avals; g.bvals = bvals; g.nsnp = na;
    fpt_scan_range r = full_range(regend, wsize, wstep, semantics);
    if (r.window_end == 0) return FPT_OK;
    std::lock_guard<std::mutex> host_lock(g_host_mu[c->device]);
    static const bool trace = getenv("FPT_TRACE") != nullptr;     /* host-side phase times on stderr */
    const auto t0 = std::chrono::steady_clock::now();
    HostStream hs;
    CHECK(hs.make());
    Arena ar(hs.st);
    DevGenotypes d;
    UploadPlan plan;
    void *hpos;
    CHECK(pinned_slot(c, 3, (size_t)na * sizeof(int32_t), &hpos));
    CHECK(upload_genotypes(ar, &g, &d, &plan, css ? FPT_CHUNK_BYTES_CSS : FPT_CHUNK_BYTES_FET));            /* asynchronous when the caller's arrays are page-locked */
    const auto t1 = std::chrono::steady_clock::now();
    gather_positions(apos, g.asize, na, (int32_t *)hpos);
    g.pos = (const int32_t *)hpos;
    PositionCheck pc;
    pc.start(g.pos, bpos, g.bsize, na);
    const auto t2 = std::chrono::steady_clock::now();
    int rc = css ? css_scan_core(c, ar, &g, d, plan, &r, treshold, runs, drosophila, mds, out0, out1, nullptr, nullptr)
                 : fet_scan_core(c, ar, &g, d, plan, &r, perc, out0, out1, nullptr);
    const long long badk = pc.result();
    if (badk >= 0) {
        /* the outputs were written from mis-paired SNPs: put the caller's pre-zeroed arrays back */
        for (long long w = 0; w < r.window_end; w++) { out0[w] = 0.0; out1[w] = 0.0; }
        rc = fail(FPT_ERR_POSITIONS, "SNP %lld: position %d in A but %d in B", badk, apos[badk * g.asize], bpos[badk * g.bsize]);
    }
    if (trace) {
        const auto t3 = std::chrono::steady_clock::now();
        auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
            return std::chrono::duration<double, std::milli>(b - a).count(); };
        fprintf(stderr, "[fpt] %s drop-in: enqueue upload %.3f ms, gather positions %.3f ms, scan %.3f ms\n", css ? "css" : "fet",
                ms(t0, t1), ms(t1, t2), ms(t2, t3));
    }
    return rc;
}

static int fet_dropin(double *avals, double *bvals, int *apos, int *bpos, int regend, int wsize, int wstep, int alen,
                      int blen, double perc, double *scores, double *stddev, int semantics) {
    return dropin(0, avals, bvals, apos, bpos, regend, wsize, wstep, alen, blen, perc, 0, 0, 0, 0, scores, stddev, semantics);
}

static int css_dropin(double *avals, double *bvals, int *apos, int *bpos, int regend, int wsize, int wstep, int alen,
                      int blen, int treshold, int runs, int drosophila, int mds, double *scores, double *p, int semantics) {
    if (mds < 0 || mds > 2) return fail(FPT_ERR_ARG, "css: mds must be 0, 1 or 2 (got %d)", mds);
    return dropin(1, avals, bvals, apos, bpos, regend, wsize, wstep, alen, blen, 0.0, treshold, runs, drosophila, mds, scores, p, semantics);
}

/* regstart is accepted and ignored, as in the reference (windows always start at 0) */
extern "C" int fpt_fet_threadcompute(double *avals, double *bvals, int *apos, int *bpos, int regstart, int regend,
                                     int wsize, int wstep, int alen, int blen, double perc, double *scores, double *stddev) {
    (void)regstart;
    return fet_dropin(avals, bvals, apos, bpos, regend, wsize, wstep, alen, blen, perc, scores, stddev, FPT_SCAN_THREADED);
}
extern "C" int fpt_fet_compute(double *avals, double *bvals, int *apos, int *bpos, int regstart, int regend, int wsize,
                               int wstep, int alen, int blen, double perc, double *scores, double *stddev) {
    (void)regstart;
    return fet_dropin(avals, bvals, apos, bpos, regend, wsize, wstep, alen, blen, perc, scores, stddev, FPT_SCAN_SERIAL);
}
extern "C" int fpt_css_threadcompute(double *avals, double *bvals, int *apos, int *bpos, int regstart, int regend,
                                     int wsize, int wstep, int alen, int blen, int treshold, int runs, int drosophila,
                                     int mds, double *scores, double *p) {
    (void)regstart;
    return css_dropin(avals, bvals, apos, bpos, regend, wsize, wstep, alen, blen, treshold, runs, drosophila, mds, scores, p, FPT_SCAN_THREADED);
}
extern "C" int fpt_css_compute(double *avals, double *bvals, int *apos, int *bpos, int regstart, int regend, int wsize,
                               int wstep, int alen, int blen, int treshold, int runs, int drosophila, int mds,
                               double *scores, double *p) {
    (void)regstart;
    return css_dropin(avals, bvals, apos, bpos, regend, wsize, wstep, alen, blen, treshold, runs, drosophila, mds, scores, p, FPT_SCAN_SERIAL);
}

/* ================================================================================================ text ingest */
#include "fpt_ingest.h"
