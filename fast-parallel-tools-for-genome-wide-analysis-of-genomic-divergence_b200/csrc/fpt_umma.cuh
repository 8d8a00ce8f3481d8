/*
 * fpt_umma.cuh — the sm_100a plumbing of the tensor-core permutation kernel (fpt_css_perm_umma.cuh): mbarriers, bulk
 * (TMA) copies global -> shared, tensor-memory allocation, shared-memory matrix descriptors, `tcgen05.mma kind::i8`
 * (u8 x u8 -> s32, accumulator in tensor memory), `tcgen05.commit`, `tcgen05.ld`. Inline PTX only, nothing from a library.
 *
 * Operand layout (both operands K-major, no swizzle): a tile of R rows x K bytes is stored as "core matrices" of
 * 8 rows x 16 bytes (128 contiguous bytes, row r of the core matrix at byte 16 r). Core matrix (rg, kg) — rows 8 rg..8 rg+7,
 * bytes 16 kg..16 kg+15 of K — lives at  kg * (R/8 * 128) + rg * 128.  So the descriptor's
 *   leading-dimension byte offset (next core matrix along K) = R * 16,
 *   stride-dimension byte offset  (next 8-row group)         = 128,
 * and one K = 32 instruction consumes two core matrices along K: the start address advances by 2 * R * 16 per step.
 */
#ifndef FPT_UMMA_CUH
#define FPT_UMMA_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#define FPT_UD __device__ __forceinline__

/* byte offset of element (row r, k) inside a canonical K-major tile of R rows */
__host__ __device__ __forceinline__ size_t fpt_umma_tile_off(int R, int r, int k) {
    return (size_t)(k >> 4) * ((size_t)R * 16) + (size_t)(r >> 3) * 128 + (size_t)(r & 7) * 16 + (size_t)(k & 15);
}

FPT_UD uint32_t fpt_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

FPT_UD void fpt_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(fpt_smem_u32(bar)), "r"(count) : "memory");
}
FPT_UD void fpt_mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
FPT_UD void fpt_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fpt_smem_u32(bar)), "r"(bytes) : "memory");
}
FPT_UD void fpt_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fpt_smem_u32(bar)) : "memory");
}
FPT_UD bool fpt_mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(fpt_smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
/* bounded wait: a protocol error must end in a trap (a CUDA error the host reports), never in a hung GPU */
FPT_UD void fpt_mbar_wait(uint64_t *bar, uint32_t parity) {
    for (uint32_t spins = 0; !fpt_mbar_try_wait(bar, parity); spins++) {
        if (spins > (1u << 26)) { __trap(); }
    }
}

/* bulk asynchronous copy global -> shared (the TMA engine, 1-D form), completion counted in bytes on an mbarrier */
FPT_UD void fpt_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(fpt_smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(fpt_smem_u32(bar)) : "memory");
}
/* writes of the generic proxy (ordinary stores) made visible to the asynchronous proxy (TMA, tensor cores) */
FPT_UD void fpt_fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

/* tensor memory: whole-warp allocation, column count a power of two >= 32 */
FPT_UD void fpt_tmem_alloc(uint32_t *slot_in_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(fpt_smem_u32(slot_in_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
FPT_UD void fpt_tmem_free(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
FPT_UD void fpt_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
FPT_UD void fpt_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

/* shared-memory matrix descriptor, no swizzle: start address, leading / stride byte offsets (all >> 4), version 1 (sm_100) */
FPT_UD uint64_t fpt_umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ULL << 46);
}
/* instruction descriptor of kind::i8: u8 x u8 -> s32, both operands K-major, M x N tile */
__host__ __device__ __forceinline__ uint32_t fpt_umma_idesc_u8(int M, int N) {
    return (2u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
/* D[tmem] (+)= A[smem] * B[smem]^T, one K = 32 step; issued by ONE thread for the CTA */
FPT_UD void fpt_umma_u8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
/* arrive on an mbarrier when every MMA issued so far by this thread has completed */
FPT_UD void fpt_umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(fpt_smem_u32(bar)) : "memory");
}
/* 32 consecutive accumulator columns of this thread's lane (warp w reads lanes 32 (w % 4) .. +31) */
FPT_UD void fpt_tmem_ld32(uint32_t taddr, uint32_t *v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
FPT_UD void fpt_tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

#endif
