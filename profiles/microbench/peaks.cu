/*
 * peaks.cu — the roofline denominators bench.py needs that MEASURED_PEAKS.json does not carry ("fp64 peak must be
 * measured on the box with a DFMA microbenchmark", "int8 tensor peak must be measured"), measured on the B200 the bench runs on:
 *
 *   fp64      DFMA issue rate: 8 independent fma chains per thread, no memory                       -> TFLOP/s (2 flop per fma)
 *   smem      conflict-free 16-byte shared-memory loads, all SMs                                    -> TB/s and bytes/clk/SM
 *   issue     warp instructions per second of a pure integer ALU loop (4 schedulers x 148 SMs)      -> Gwarp-inst/s
 *   imma_sync mma.sync.m16n8k32 u8 x u8 -> s32 (the tensor-core form of the m <= 64 permutation kernel) -> TOP/s
 *   umma_i8   tcgen05.mma kind::i8 128 x 256 x 32, BOTH operands resident in shared memory (no TMA, no loads in the loop),
 *             two alternating tensor-memory accumulators                                             -> TOP/s, cycles per MMA
 *
 * Prints one JSON object. Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o peaks peaks.cu
 * Measurement infrastructure, not product code; the tcgen05 plumbing is the product's own header.
 */
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../fast-parallel-tools-for-genome-wide-analysis-of-genomic-divergence_b200/csrc/fpt_umma.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); exit(2); } } while (0)

template <typename F>
static float best_ms(F launch, int reps = 5) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    launch();                                              /* warm-up */
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(a));
        launch();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

/* ------------------------------------------------------------------ fp64 */
__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

/* ------------------------------------------------------------------ shared memory */
__global__ void __launch_bounds__(1024) smem_kernel(unsigned *out, int iters) {
    __shared__ uint4 buf[2048];                            /* 32 KB */
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) buf[i] = make_uint4(i, i + 1, i + 2, i + 3);
    __syncthreads();
    unsigned acc = 0;
    const unsigned base = (unsigned)__cvta_generic_to_shared(buf);
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            /* volatile asm: the loads cannot be hoisted or merged. Consecutive lanes -> consecutive 16-byte words: conflict-free */
            unsigned x, y, z, w;
            const unsigned addr = base + ((((unsigned)threadIdx.x + u * 256 + i * 32) & 2047u) << 4);
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(addr));
            acc ^= x ^ y ^ z ^ w;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

/* ------------------------------------------------------------------ issue rate */
__global__ void __launch_bounds__(256) issue_kernel(unsigned *out, int iters, unsigned k) {
    unsigned x0 = threadIdx.x, x1 = x0 ^ 1, x2 = x0 ^ 2, x3 = x0 ^ 3, x4 = x0 ^ 4, x5 = x0 ^ 5, x6 = x0 ^ 6, x7 = x0 ^ 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
            x0 = (x0 ^ k) + x1; x1 = (x1 ^ k) + x2; x2 = (x2 ^ k) + x3; x3 = (x3 ^ k) + x4;
            x4 = (x4 ^ k) + x5; x5 = (x5 ^ k) + x6; x6 = (x6 ^ k) + x7; x7 = (x7 ^ k) + x0;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;
}

/* ------------------------------------------------------------------ mma.sync u8 */
__global__ void __launch_bounds__(256) imma_kernel(int *out, int iters) {
    int c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0}, c2[4] = {0, 0, 0, 0}, c3[4] = {0, 0, 0, 0};
    unsigned a[4] = {threadIdx.x, threadIdx.x + 1u, threadIdx.x + 2u, threadIdx.x + 3u}, b0 = threadIdx.x * 7u, b1 = threadIdx.x * 13u;
    for (int i = 0; i < iters; i++) {
#define MMA(c) asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n" \
                            : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1))
        MMA(c0); MMA(c1); MMA(c2); MMA(c3);
#undef MMA
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c1[1] + c2[2] + c3[3] + c0[3] + c1[2] + c2[1] + c3[0];
}

/* ------------------------------------------------------------------ tcgen05 kind::i8, operands resident */
#define UK 128
__global__ void __launch_bounds__(128, 1) umma_kernel(int nmma_groups, long long *cycles, int *sink) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar_done;
    __shared__ uint32_t tmem_slot;
    unsigned char *sA = smem, *sB = smem + (size_t)128 * UK;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 384 * UK; i += blockDim.x) smem[i] = (unsigned char)(i & 1);
    fpt_fence_proxy_async();
    if (tid == 0) { fpt_mbar_init(&bar_done, 1); fpt_mbar_fence_init(); }
    if (warp == 0) fpt_tmem_alloc(&tmem_slot, 512);
    fpt_tc_fence_before();
    __syncthreads();
    fpt_tc_fence_after();
    const uint32_t tmem = tmem_slot;
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        const uint32_t idesc = fpt_umma_idesc_u8(128, 256);
        t0 = clock64();
        for (int g = 0; g < nmma_groups; g++) {
            const uint32_t acc = tmem + (uint32_t)(g & 1) * 256u;
            for (int ks = 0; ks < UK / 32; ks++) {
                const uint64_t ad = fpt_umma_desc(fpt_smem_u32(sA) + ks * 2 * 128 * 16, 128 * 16, 128);
                const uint64_t bd = fpt_umma_desc(fpt_smem_u32(sB) + ks * 2 * 256 * 16, 256 * 16, 128);
                fpt_umma_u8(acc, ad, bd, idesc, (g > 1 || ks > 0) ? 1u : 0u);
            }
        }
        fpt_umma_commit(&bar_done);
        fpt_mbar_wait(&bar_done, 0);
        t1 = clock64();
        cycles[blockIdx.x] = t1 - t0;
    }
    __syncthreads();
    fpt_tc_fence_after();
    uint32_t v[32];
    fpt_tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16), v);
    fpt_tmem_ld_wait();
    sink[blockIdx.x * blockDim.x + tid] = (int)v[tid & 31];
    fpt_tc_fence_before();
    __syncthreads();
    if (warp == 0) fpt_tmem_free(tmem, 512);
}

int main() {
    cudaDeviceProp pr;
    CK(cudaGetDeviceProperties(&pr, 0));
    const int sms = pr.multiProcessorCount;
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    void *scratch;
    CK(cudaMalloc(&scratch, (size_t)64 << 20));
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"sm_max_mhz\": %.0f", pr.name, sms, clk_khz / 1000.0);

    {   /* fp64 */
        const int iters = 1 << 14, grid = sms * 8, block = 256;
        const float ms = best_ms([&] { dfma_kernel<<<grid, block>>>((double *)scratch, iters, 1.0000001, 1e-9); });
        const double fma = (double)grid * block * 8.0 * iters;
        printf(", \"fp64\": {\"tflops\": %.3f, \"dfma_per_clk_per_sm\": %.2f, \"ms\": %.4f, \"how\": \"8 independent DFMA chains/thread, %d CTAs x %d threads x %d iterations\"}",
               2.0 * fma / (ms * 1e-3) / 1e12, fma / (ms * 1e-3) / (sms * (clk_khz * 1e3)), ms, grid, block, iters);
    }
    {   /* smem */
        const int iters = 1 << 12, grid = sms * 2, block = 1024;
        const float ms = best_ms([&] { smem_kernel<<<grid, block>>>((unsigned *)scratch, iters); });
        const double bytes = (double)grid * block * iters * 8.0 * 16.0;
        printf(", \"smem\": {\"tb_per_s\": %.3f, \"bytes_per_clk_per_sm\": %.2f, \"ms\": %.4f, \"how\": \"conflict-free LDS.128, %d CTAs x %d threads\"}",
               bytes / (ms * 1e-3) / 1e12, bytes / (ms * 1e-3) / (sms * (clk_khz * 1e3)), ms, grid, block);
    }
    {   /* issue */
        const int iters = 1 << 12, grid = sms * 8, block = 256;
        const float ms = best_ms([&] { issue_kernel<<<grid, block>>>((unsigned *)scratch, iters, 0x9e3779b9u); });
        const double winst = (double)grid * (block / 32) * iters * 4.0 * 8.0 * 2.0;     /* SASS: one LOP3 + one IMAD per statement */
        printf(", \"issue\": {\"gwarp_inst_per_s\": %.1f, \"per_clk_per_sm\": %.2f, \"ms\": %.4f, \"how\": \"integer ALU loop, LOP3 + IMAD per statement (checked in SASS), %d CTAs x %d threads\"}",
               winst / (ms * 1e-3) / 1e9, winst / (ms * 1e-3) / (sms * (clk_khz * 1e3)), ms, grid, block);
    }
    {   /* mma.sync u8 */
        const int iters = 1 << 13, grid = sms * 8, block = 256;
        const float ms = best_ms([&] { imma_kernel<<<grid, block>>>((int *)scratch, iters); });
        const double macs = (double)grid * (block / 32) * iters * 4.0 * 16.0 * 8.0 * 32.0;
        printf(", \"imma_sync_u8\": {\"tops\": %.1f, \"mac_per_clk_per_sm\": %.1f, \"ms\": %.4f, \"how\": \"mma.sync.m16n8k32 u8, 4 independent accumulators/warp, %d CTAs x %d threads\"}",
               2.0 * macs / (ms * 1e-3) / 1e12, macs / (ms * 1e-3) / (sms * (clk_khz * 1e3)), ms, grid, block);
    }
    {   /* tcgen05 kind::i8 */
        const int groups = 4096;                            /* x 4 MMAs of 128 x 256 x 32 */
        long long *cyc = (long long *)scratch;
        int *sink = (int *)((char *)scratch + (1 << 20));
        CK(cudaFuncSetAttribute(umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 384 * UK + 70 * 1024));
        const size_t smem = 384 * UK + 70 * 1024;           /* > half an SM: one CTA per SM */
        const float ms = best_ms([&] { umma_kernel<<<sms, 128, smem>>>(groups, cyc, sink); }, 3);
        std::vector<long long> h(sms);
        CK(cudaMemcpy(h.data(), cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost));
        long long mx = 0; double mean = 0;
        for (int i = 0; i < sms; i++) { mx = h[i] > mx ? h[i] : mx; mean += (double)h[i] / sms; }
        const double nmma = (double)groups * (UK / 32);
        const double macs = (double)sms * nmma * 128.0 * 256.0 * 32.0;
        printf(", \"umma_i8\": {\"tops\": %.1f, \"cycles_per_mma_128x256x32\": %.1f, \"mac_per_clk_per_sm\": %.1f, \"ms\": %.4f, "
               "\"how\": \"tcgen05.mma.cta_group::1.kind::i8, A 128x128 and B 256x128 u8 resident in shared memory, %d MMAs per SM back to back, 2 accumulators, clock64 around issue..commit\"}",
               2.0 * macs / (ms * 1e-3) / 1e12, mean / nmma, 128.0 * 256.0 * 32.0 / (mean / nmma), ms, (int)nmma);
        (void)mx;
    }
    printf("}\n");
    return 0;
}
