// tc_common.cuh -- tcgen05 / TMEM / mbarrier PTX wrappers shared by the tensor-core kernels (tc_gemm.cu, als_tc.cu).
// Operands are staged in shared memory in the canonical K-major no-swizzle UMMA layout; accumulators live in TMEM.
#pragma once
#include "common.cuh"

namespace cymf {
namespace tc {

constexpr int TILE_M = 128;       // rows of D (TMEM lanes)
constexpr int TC_GRAM_SLAB = 512; // rows of Y per Gram partial
constexpr int CHUNK_K = 32;       // reduction elements staged per step: 8 x 16-byte chunks, 4 MMA k-slices of 8

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {      // release.cta: this thread's earlier writes are published
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// One lane of a converged warp.  MMA-issuing code guarded by this (instead of `lane == 0`) lets the compiler keep the
// descriptors in uniform registers: with a lane test it cannot tell that a single thread is active and wraps every
// tcgen05.mma in an ELECT / R2UR.BROADCAST / BRA.U.ANY loop (~100 cycles per MMA, measured in als_ws.cu).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
// all previously issued MMAs of this thread arrive on the mbarrier when they have completed
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, kind::tf32, M = 128, N from the instruction descriptor, K = 8
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// shared-memory matrix descriptor, canonical K-major layout without swizzle: 8-row x 16-byte core matrices, rows of a
// core matrix 16 bytes apart, `sbo` bytes between 8-row groups, `lbo` bytes between the two 16-byte K chunks of a slice
__device__ __forceinline__ uint64_t smem_desc(const void *p, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((smem_u32(p) >> 4) & 0x3fffu) | ((uint64_t)((lbo >> 4) & 0x3fffu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3fffu) << 32) | (1ull << 46);
}
// instruction descriptor: D f32, A/B tf32, both K-major, dense, M = 128
__host__ __device__ constexpr uint32_t idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
}
// 32 consecutive f32 columns of this warp's 32 TMEM lanes: thread l of the warp receives lane (row) l
__device__ __forceinline__ void tmem_load32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int t = 0; t < 32; ++t) v[t] = __uint_as_float(r[t]);
}

__device__ __forceinline__ float4 tf32_hi(float4 a) {
    return make_float4(__uint_as_float(__float_as_uint(a.x) & 0xffffe000u), __uint_as_float(__float_as_uint(a.y) & 0xffffe000u),
                       __uint_as_float(__float_as_uint(a.z) & 0xffffe000u), __uint_as_float(__float_as_uint(a.w) & 0xffffe000u));
}
__device__ __forceinline__ float4 sub4(float4 a, float4 b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }

// Operand tile in shared memory: `groups` 8-row groups x 8 chunks of 16 bytes (32 reduction elements):
//   byte offset of (row m, chunk q) = q * groups * 128 + (m / 8) * 128 + (m % 8) * 16
__device__ __forceinline__ int tile_off(int m, int q, int groups) { return (q * groups * 128 + (m >> 3) * 128 + (m & 7) * 16) >> 2; }

// issue the 3 x 4 MMAs of one staged 32-element reduction chunk (thread 0 only)
__device__ __forceinline__ void issue_chunk(uint32_t d_tmem, const float *a_hi, const float *a_lo, const float *b_hi,
                                            const float *b_lo, int a_groups, int b_groups, uint32_t idesc, bool first) {
    const uint32_t lbo_a = a_groups * 128, lbo_b = b_groups * 128;
#pragma unroll
    for (int ks = 0; ks < CHUNK_K / 8; ++ks) {
        const uint64_t ah = smem_desc(a_hi + ks * 2 * (lbo_a >> 2), lbo_a, 128), al = smem_desc(a_lo + ks * 2 * (lbo_a >> 2), lbo_a, 128);
        const uint64_t bh = smem_desc(b_hi + ks * 2 * (lbo_b >> 2), lbo_b, 128), bl = smem_desc(b_lo + ks * 2 * (lbo_b >> 2), lbo_b, 128);
        mma_tf32(d_tmem, ah, bl, idesc, (first && ks == 0) ? 0u : 1u);     // small terms first
        mma_tf32(d_tmem, al, bh, idesc, 1u);
        mma_tf32(d_tmem, ah, bh, idesc, 1u);
    }
}


// 16 consecutive f32 columns of this warp's 32 TMEM lanes
__device__ __forceinline__ void tmem_load16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int t = 0; t < 16; ++t) v[t] = __uint_as_float(r[t]);
}

// the same without waiting: issue several, then tmem_wait_ld16 on each group before its values are used
__device__ __forceinline__ void tmem_load16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// tcgen05.wait::ld that the compiler must order before any use of the 16 loaded registers
__device__ __forceinline__ void tmem_wait_ld16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                   "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}

// Ampere-style asynchronous copies global -> shared (LDGSTS): no register is held while the data is in flight
__device__ __forceinline__ void cp_async4(void *smem, const void *gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// packed pairs of f32 (Blackwell FFMA2: two independent IEEE fused multiply-adds per instruction)
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void fma2(unsigned long long &acc, unsigned long long a, unsigned long long b) {
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

}  // namespace tc
}  // namespace cymf
