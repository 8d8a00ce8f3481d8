/* Stand-alone check of the tcgen05 kind::i8 plumbing in csrc/fpt_umma.cuh against a host product (test infrastructure).
   Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu ; run on a B200. */
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../fast-parallel-tools-for-genome-wide-analysis-of-genomic-divergence_b200/csrc/fpt_umma.cuh"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); return 2; } } while (0)

__global__ void __launch_bounds__(128, 1)
probe_kernel(const unsigned char *A, const unsigned char *B, int K, uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo,
             uint32_t a_kstep, uint32_t b_kstep, uint64_t desc_or, int *out) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar_full, bar_done;
    __shared__ uint32_t tmem_slot;
    unsigned char *sA = smem, *sB = smem + (size_t)128 * K;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { fpt_mbar_init(&bar_full, 1); fpt_mbar_init(&bar_done, 1); fpt_mbar_fence_init(); }
    if (warp == 0) fpt_tmem_alloc(&tmem_slot, 256);
    fpt_tc_fence_before();
    __syncthreads();
    fpt_tc_fence_after();
    const uint32_t tmem = tmem_slot;
    if (tid == 0) {
        fpt_mbar_expect_tx(&bar_full, (uint32_t)(384 * K));
        fpt_bulk_g2s(sA, A, (uint32_t)(128 * K), &bar_full);
        fpt_bulk_g2s(sB, B, (uint32_t)(256 * K), &bar_full);
        fpt_mbar_wait(&bar_full, 0);
        fpt_tc_fence_after();
        const uint32_t idesc = fpt_umma_idesc_u8(128, 256);
        for (int ks = 0; ks < K / 32; ks++) {
            const uint64_t ad = fpt_umma_desc(fpt_smem_u32(sA) + ks * a_kstep, a_lbo, a_sbo) | desc_or;
            const uint64_t bd = fpt_umma_desc(fpt_smem_u32(sB) + ks * b_kstep, b_lbo, b_sbo) | desc_or;
            fpt_umma_u8(tmem, ad, bd, idesc, ks > 0);
        }
        fpt_umma_commit(&bar_done);
    }
    __syncwarp();
    fpt_mbar_wait(&bar_done, 0);
    fpt_tc_fence_after();
    for (int c = 0; c < 256; c += 32) {
        uint32_t v[32];
        fpt_tmem_ld32(tmem + ((uint32_t)(32 * warp) << 16) + c, v);
        fpt_tmem_ld_wait();
        for (int i = 0; i < 32; i++) out[(size_t)tid * 256 + c + i] = (int)v[i];
    }
    fpt_tc_fence_before();
    __syncthreads();
    if (warp == 0) fpt_tmem_free(tmem, 256);
}

int main() {
    const int K = 128;
    std::vector<unsigned char> A(128 * K), B(256 * K), Ai(128 * K), Bi(256 * K);
    srand(7);
    for (int r = 0; r < 128; r++) for (int k = 0; k < K; k++) { A[r * K + k] = rand() & 1; Ai[fpt_umma_tile_off(128, r, k)] = A[r * K + k]; }
    for (int r = 0; r < 256; r++) for (int k = 0; k < K; k++) { B[r * K + k] = rand() & 255; Bi[fpt_umma_tile_off(256, r, k)] = B[r * K + k]; }
    std::vector<int> ref(128 * 256), got(128 * 256);
    for (int i = 0; i < 128; i++) for (int j = 0; j < 256; j++) { int s = 0; for (int k = 0; k < K; k++) s += A[i * K + k] * B[j * K + k]; ref[i * 256 + j] = s; }
    unsigned char *dA, *dB; int *dO;
    CK(cudaMalloc(&dA, A.size())); CK(cudaMalloc(&dB, B.size())); CK(cudaMalloc(&dO, got.size() * 4));
    CK(cudaMemcpy(dA, Ai.data(), A.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, Bi.data(), B.size(), cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 384 * K));
    int winner = -1;
    {
        /* the layout of fpt_umma.cuh: leading-dimension offset = rows * 16 (next core matrix along K), stride offset = 128 */
        const int variant = 0;
        CK(cudaMemset(dO, 0xff, got.size() * 4));
        probe_kernel<<<1, 128, 384 * K>>>(dA, dB, K, 128 * 16, 128, 256 * 16, 128, 2 * 128 * 16, 2 * 256 * 16, 0ULL, dO);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 3; }
        CK(cudaMemcpy(got.data(), dO, got.size() * 4, cudaMemcpyDeviceToHost));
        long bad = 0; for (size_t i = 0; i < got.size(); i++) bad += got[i] != ref[i];
        printf("%ld of %zu mismatches; got[0..3]=%d %d %d %d ref=%d %d %d %d\n", bad, got.size(),
               got[0], got[1], got[2], got[3], ref[0], ref[1], ref[2], ref[3]);
        if (!bad) winner = variant;
    }
    printf("UMMA_PROBE %s winner=%d\n", winner >= 0 ? "OK" : "FAIL", winner);
    return winner >= 0 ? 0 : 1;
}
