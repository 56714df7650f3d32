// tcgen05 / TMEM building blocks for the tensor-core stages of the registration chain (sm_100a only).
//
// The registration path contracts the SHORT strip axis (n = 214 for 2048^2 tiles: 2 * 107, a prime that no radix
// FFT digests) with dense DFT matrices on the 5th-generation tensor cores: D[128 x N] (fp32, in TMEM) +=
// A[128 x 8] . B[8 x N] per tcgen05.mma.kind::tf32, operands in shared memory, issued by one thread.  float32
// accuracy comes from the 3-term split a = a_hi + a_lo, b = b_hi + b_lo (tf32 each):
//     a b ~= a_hi b_hi + a_lo b_hi + a_hi b_lo           (dropped: a_lo b_lo <= 2^-22 |a b|)
// accumulated in fp32 -- three MMAs per K step into the same accumulator.
//
// Shared-memory operand layout (both A and B; "K-major, no swizzle" canonical form of the matrix descriptor):
//     element (row r, k)  ->  byte  (k / 4) * LBO  +  (r / 8) * SBO  +  (r % 8) * 16  +  (k % 4) * 4
// i.e. 8 rows x 16 bytes core matrices; SBO = 128 makes the rows of one K chunk contiguous at a 16-byte pitch
// (conflict-free 128-bit stores, one row per thread); LBO = distance between K chunks of 4 elements.
#pragma once

#include <cuda_runtime.h>
#include <cstdint>

namespace umma {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// 64-bit shared-memory matrix descriptor: start address, leading (K-chunk) and stride (8-row group) byte offsets in
// 16-byte units, descriptor version 1 (Blackwell) at bit 46, no swizzle.
__device__ __forceinline__ uint64_t desc_kmajor(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46);
}

// 32-bit instruction descriptor of kind::tf32: D = fp32 (1 at bits [4,6)), A and B = TF32 (2 at [7,10) and [10,13)),
// both K-major (bits 15, 16 = 0), N >> 3 at [17,23), M >> 4 at [24,29); bits 13 / 14 negate A / B.
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, bool neg_a = false, bool neg_b = false) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((neg_a ? 1u : 0u) << 13) | ((neg_b ? 1u : 0u) << 14) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- TMEM allocation (one warp, all lanes; ncols a power of two >= 32)
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(smem_dst)), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- ordering
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes (st.shared) -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one lane of a CONVERGED warp (elect.sync): the single-thread tcgen05 / bulk-copy issue without leaving uniform control flow
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---- MMA issue (ONE thread): D[tmem] (+)= A[smem desc] . B[smem desc]
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the mbarrier when every MMA this thread issued so far has completed (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}

// ---- TMEM -> registers: 32 lanes x 32 bit, 16 / 32 consecutive columns per thread (the warp's own lane quadrant)
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- mbarrier (shared::cta), bounded wait: a malformed pipeline must end in an error, never in a hung GPU
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_addr(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// try_wait with a suspend-time hint: the thread may sleep up to `ns` and wakes when the phase completes
__device__ __forceinline__ bool mbar_try_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_addr(bar)), "r"(parity), "r"(ns)
        : "memory");
    return ok != 0;
}
// returns false after ~2^31 cycles (about a second) without completion.  The waiting threads SLEEP in try_wait (time hint)
// instead of polling: with plain try_wait + a clock read per iteration the spin loops of the loader / MMA / epilogue /
// converter warps issued 27 % of all instructions of the forward kernel (r2 ncu source counters) -- issue slots taken
// from the converter warps that bound it.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try(bar, parity)) return true;
    const long long t0 = clock64();
    for (uint32_t i = 1;; ++i) {
        if (mbar_try_hint(bar, parity, 100000u)) return true;
        if ((i & 15u) == 0u && clock64() - t0 > (1ll << 31)) return false;
    }
}
// 1-D bulk copy global -> shared with completion on an mbarrier (both addresses and the size multiples of 16 bytes)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

// ---- cp.async (LDGSTS), 8 bytes: global -> shared without staging registers; groups complete in order
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {      // .cg: L2 only, no L1 line allocated
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- tf32 split: hi = v rounded to tf32 (RNA), lo = v - hi (exactly representable in fp32; its own tf32 rounding is
// applied by the tensor core when it reads the operand, which truncates the 13 low mantissa bits)
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
    hi = __uint_as_float(h);
    lo = v - hi;
}

}  // namespace umma
