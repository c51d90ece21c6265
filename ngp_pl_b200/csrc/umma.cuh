// tcgen05 (5th-generation tensor core) building blocks for sm_100a, hand-written PTX: TMEM allocation, shared-memory matrix
// descriptors, single-thread MMA issue with the accumulator in TMEM, completion through an mbarrier, TMEM -> register loads.
// Used by the weight-gradient GEMMs of the MLP backward (network.cu: k_ngp_bwd3): dW = dOut^T * In over the staged rows of
// a block is a K = 256 GEMM with M, N <= 64 -- one elected thread issues it asynchronously while the sixteen warps carry on
// with the per-row dgrad chain, and the five accumulators (160 fp32 columns) live in TMEM for the CTA's whole lifetime
// instead of in registers.
//
// Operand layout (both operands are "MN-major": for every staged sample row k the M (or N) channel values are contiguous),
// no swizzle. In units of 16 bytes (T = 8 halves) the canonical layout the hardware expects is
//     ((8, m), (8, k)) : ((1 elem, SBO), (16 B, LBO))
// i.e. 8 channels x 8 rows form a 128-byte core matrix (row r of it at +16 r bytes, channel c at +2 c bytes); core matrices
// that are neighbours along the channels are SBO bytes apart, neighbours along the rows LBO bytes apart. A [rows][C]
// activation tile is stored with SBO = 128 B and LBO = C/8 * 128 B:
//     offset(row, ch) = (row / 8) * LBO + (ch / 8) * 128 + (row % 8) * 16 + (ch % 8) * 2      [bytes]
// One tcgen05.mma.kind::f16 consumes K = 16 rows (two core matrices along K); advancing K by 16 = start address + 2 LBO.
// Matrix-descriptor / instruction-descriptor bit layouts follow CUTLASS's cute/arch/mma_sm100_desc.hpp (SmemDescriptor,
// InstrDescriptor); the M = 64 accumulator layout in TMEM (row m -> lane (m % 16) + 32 (m / 16), column n) its
// cute/atom/mma_traits_sm100.hpp (tmem_frg, "half subpartitions").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

// half-element offset of (row, ch) in the canonical MN-major tile of C channels
__device__ __forceinline__ int umma_off(int row, int ch, int C) {
    return (row >> 3) * (C * 8) + (ch >> 3) * 64 + (row & 7) * 8 + (ch & 7);
}

__device__ __forceinline__ uint64_t umma_smem_desc(const void* smem, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(smem);
    uint64_t d = 0;
    d |= (uint64_t)((a >> 4) & 0x3fffu);              // start address, bits [0,14)
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;  // leading byte offset, bits [16,30)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;  // stride byte offset, bits [32,46)
    d |= (uint64_t)1 << 46;                             // descriptor version 1 (Blackwell), bits [46,48)
    return d;                                           // base offset 0, lbo mode 0, layout type 0 = no swizzle
}
// kind::f16, A and B fp16 (format 0), D fp32, both operands MN-major, M x N
__device__ __forceinline__ uint32_t umma_instr_desc_f16(int M, int N) {
    return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// make every MMA issued so far by this thread arrive on the mbarrier when it has completed
__device__ __forceinline__ void umma_commit(uint64_t* mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(mbar))
                 : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (the tensor core reads the operands through it)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t* mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(mbar)), "r"(count) : "memory");
}
// spin until the phase with the given parity has completed; traps after ~2 s (a malformed descriptor must fail the launch
// loudly, not hang the GPU)
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(mbar);
    const long long t0 = clock64();
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
        if (done) return;
        if (clock64() - t0 > 4000000000ll) asm volatile("trap;");
    }
}

// TMEM: `cols` (power of two >= 32) columns x 128 lanes x 32 bit. One warp allocates and later frees.
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(smem_result)),
                 "n"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_free(uint32_t tmem_addr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_addr), "n"(COLS) : "memory");
}
// this warp's 32 lanes x 16 consecutive columns -> 16 registers per thread (thread t = lane 32 (warp % 4) + t)
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
