// ptx_sm100.cuh -- inline-PTX wrappers for the sm_100a features contract_i8.cu uses: mbarriers, cp.async with
// mbarrier completion, tensor memory (tcgen05.alloc / ld) and the 5th-generation tensor-core MMA (tcgen05.mma).
#pragma once
#include <cstdint>

namespace scde {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// ---- cp.async (16 bytes, L2 only) with completion counted on an mbarrier ---------------------------------------------
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
// 4-byte copies (through L1) and group completion, for small staged headers (lp_table.cu)
__device__ __forceinline__ void cp_async4(uint32_t dst_smem, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void st_shared_u16(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"((uint16_t)v) : "memory");
}
__device__ __forceinline__ void st_shared_u8(uint32_t addr, uint32_t v) {
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v & 0xFFu) : "memory");
}
// the barrier receives one arrival once every cp.async this thread has issued so far has landed; .noinc: the arrival is
// part of the barrier's expected count (initialise the barrier with the number of threads that call this per phase)
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// generic-proxy writes (cp.async, st.shared) -> async-proxy reads (tcgen05.mma operands)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- tensor memory -------------------------------------------------------------------------------------------------
// whole warp; writes the base address (lane << 16 | column) of `cols` (power of two >= 32) columns to *dst_smem
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// 32 lanes x 8 consecutive 32-bit columns: thread t of the warp receives lane (taddr.lane + t), columns taddr.col .. +7
__device__ __forceinline__ void tmem_ld_32x32b_x8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
}
// 32 lanes x 2 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld_32x32b_x2(uint32_t taddr, uint32_t &r0, uint32_t &r1) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- tcgen05.mma ---------------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, no-swizzle ("interleave") canonical layout: the operand is a grid of 8 x 16-byte core
// matrices, each stored contiguously (128 bytes).  For an MN-major operand (16 consecutive M/N elements per 16-byte
// row, 8 consecutive K per core matrix) the stride between core matrices along M/N goes into the "stride byte offset"
// field and the stride between groups of 8 K into the "leading byte offset" field; `swap` exchanges them (probe).
__device__ __forceinline__ uint64_t umma_desc_nosw(uint32_t saddr, uint32_t stride_mn, uint32_t stride_k, bool swap) {
    const uint32_t lbo = swap ? stride_mn : stride_k, sbo = swap ? stride_k : stride_mn;
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);  // bits 46-47: descriptor version 1 (sm_100)
}
// 128-byte-swizzle canonical layout of an MN-major operand: runs of 128 consecutive M/N bytes, 8 consecutive K per
// 1024-byte atom (K stride 128 bytes, 16-byte pieces XOR-swizzled with K mod 8); "leading byte offset" = stride between
// 128-byte runs along M/N, "stride byte offset" = stride between groups of 8 K.  Atoms must be 1024-byte aligned.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr, uint32_t stride_mn_run, uint32_t stride_k_group) {
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((stride_mn_run >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((stride_k_group >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);  // layout type 2 = SWIZZLE_128B
}
// Instruction descriptor of kind::i8: signed 8-bit A and B, both MN-major, 32-bit integer accumulators, M x N tile
__device__ __forceinline__ uint32_t umma_idesc_s8_mn(int M, int N) {
    return (2u << 4)      // D format: S32
           | (1u << 7)    // A: signed 8-bit
           | (1u << 10)   // B: signed 8-bit
           | (1u << 15)   // A is MN-major
           | (1u << 16)   // B is MN-major
           | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]; one thread issues for the CTA
__device__ __forceinline__ void umma_s8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, bool accumulate) {
    const uint32_t acc = accumulate ? 1u : 0u, zero = 0u;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc), "r"(zero)
        : "memory");
}
// the barrier receives one arrival when every tcgen05.mma this thread has issued so far has completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

}  // namespace ptx
}  // namespace scde
