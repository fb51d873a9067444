// als_tc6.cu -- the one-pass WMF row solver of als_tc.cu with a fully asynchronous gather.
//
// Same mathematics and the same conjugate-gradient phase (see als_tc.cu); what changes is how the row's item vectors
// reach the tensor core.  als_tc.cu stages the operand tile K-major, which forces a transposing gather: 4-byte loads
// into registers, one chunk ahead at most, so a row of n entries pays ~n/32 exposed L2 round trips.  Here the tile is
// MN-MAJOR (swizzle 128B_BASE32B; T[m][i] = y~_i[m]: an item vector IS a contiguous run of the MN dimension), so
//   * one warp instruction (cp.async, 16 bytes per lane, LDGSTS.128) moves a whole 512-byte item vector from global
//     memory straight into its four swizzled 128-byte rows of the tile -- no registers, conflict-free, 4x fewer
//     instructions than the transposing gather;
//   * gathers run TWO chunks (64 entries) ahead of the conversion in a ring of four 16 KB "hi" slots, and the first two
//     chunks of the NEXT row are issued while the current row is still in its CG iterations, so a short row meets its
//     data already in shared memory;
//   * every thread converts exactly the 16-byte pieces it copied itself (hi = a with the low 13 mantissa bits cleared,
//     in place; lo = a - hi into a ring of two "lo" slots), so no cross-thread wait on the copies is needed:
//     cp.async.wait_group, then the mbarrier hand-over to the (rotating) MMA-issuing thread as in als_tc.cu.
// Shared memory per CTA: 4 x 16 KB + 2 x 16 KB of tiles + 7 KB of vectors; two CTAs per SM.
#include <stdlib.h>

#include "tc_common.cuh"

namespace cymf {
namespace tc {

constexpr int R6_THREADS = 256;
constexpr int R6_HI = 4;          // raw / hi slots (16 KB each): chunk g lives in slot g mod 4
constexpr int R6_LO = 2;          // lo slots: chunk g's lo tile lives in slot g mod 2
constexpr int R6_CHAIN = 16;      // 32-item chunks accumulated in TMEM before the chain is folded into registers
constexpr int R6_TILE_FLOATS = TILE_M * CHUNK_K;      // 4096 floats = 16 KB

struct Row6Args {
    const int64_t *indptr;
    const int32_t *indices;
    const int32_t *order;
    int32_t n_solve;
    float *X;
    const float *Y;
    int32_t max_iter;
    float weight, tol2;
    int32_t *queue;
    unsigned long long *stats;
};

struct Row6Info { int r; int nnz; long long lo; };

__device__ __forceinline__ float warp_sum6(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// Byte offset, inside a 16 KB MN-major tile of 128 (MN) x 32 (K) f32, of the 16-byte piece holding elements
// m = 4 g .. 4 g + 3 of reduction index k.  MN-major tf32 operands have ONE legal shared-memory layout,
// SWIZZLE_128B_BASE32B: atoms of 4 reduction rows x 128 bytes (32 MN elements) whose 32-byte chunks are XOR-ed with the
// row number (address bits [5,7) ^= bits [7,9)); here two such atoms are contiguous (SBO = 512 B: one K = 8 MMA slice
// is 1024 B), 32-element MN blocks are 1024 B apart (LBO) and the four K slices of a chunk 4096 B apart.
__device__ __forceinline__ int mn_off(int k, int g) {
    return (k >> 3) * 4096 + (g >> 3) * 1024 + (k & 7) * 128 + (((((g & 7) >> 1) ^ (k & 3))) << 5) + ((g & 1) << 4);
}
// shared-memory matrix descriptor of one K slice (8 reduction indices) of such a tile
__device__ __forceinline__ uint64_t mn_desc(const void *p) {
    return (uint64_t)((smem_u32(p) >> 4) & 0x3fffu) | ((uint64_t)(1024u >> 4) << 16) | ((uint64_t)(512u >> 4) << 32) |
           (1ull << 46) | (1ull << 61);
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_upto(int n) {       // at most n committed groups still in flight
    if (n <= 0) asm volatile("cp.async.wait_group 0;" ::: "memory");
    else if (n == 1) asm volatile("cp.async.wait_group 1;" ::: "memory");
    else asm volatile("cp.async.wait_group 2;" ::: "memory");
}

// mbarrier wait that sleeps in hardware for up to ~1 us per attempt instead of re-issuing the test every few cycles:
// waiting warps must not eat the issue slots of the warps that do the work
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAITS_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONES_%=;\n\t"
        "bra WAITS_%=;\n\t"
        "DONES_%=:\n\t"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(1000u)
        : "memory");
}
__device__ __forceinline__ unsigned long long sub2(unsigned long long a, unsigned long long b) {      // a - b, packed pair
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// RAWHI: the hi operand is the gathered data itself.  The tensor core reads tf32 operands from 32-bit containers and
// ignores the 13 low mantissa bits, i.e. it sees exactly hi = a & 0xffffe000, so only lo = a - hi has to be produced.
template <int LD, bool RAWHI>
__global__ void __launch_bounds__(R6_THREADS, 2) als_rows_tc6_kernel(const Row6Args a) {
    constexpr int HC = LD / 2;
    constexpr int IDX_BLOCK = R6_THREADS;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float *const hi_s = reinterpret_cast<float *>(smem_raw);                 // [4][4096]
    float *const lo_s = hi_s + R6_HI * R6_TILE_FLOATS;                       // [2][4096]
    float *const p_s = lo_s + R6_LO * R6_TILE_FLOATS;                        // [128]
    float *const part = p_s + 128;                                           // [2][2][128]
    float *const bpart = part + 512;                                         // [8][128] per-warp sums of the item vectors
    float *const x0_s = bpart + 1024;                                        // [128]
    int32_t *const idx_s = reinterpret_cast<int32_t *>(x0_s + 128);          // [3][256]
    __shared__ uint64_t mb_full[R6_HI];      // "all eight warps have converted their part of this chunk"
    __shared__ uint64_t mb_done[R6_HI];      // "the MMAs that read this chunk's tiles have completed"
    __shared__ uint64_t mb_acc[2];
    __shared__ uint32_t tmem_slot;
    __shared__ Row6Info rowq[3];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m = tid & 127, h = tid >> 7;   // row of S / column half (fold and CG phases)
    const bool l_on = 4 * lane < LD;         // this lane's 4-element slice (m-group `lane`) exists
    constexpr uint32_t tmem_cols = (2 * LD <= 32) ? 32 : (2 * LD <= 64) ? 64 : (2 * LD <= 128) ? 128 : 256;
    if (warp == 0) tmem_alloc(&tmem_slot, tmem_cols);

    int slotA = 0, rA = -1, rB = -1;
    long long loB = 0, hiB = 0;
    auto claim_now = [&](int &r, long long &lo, long long &hi) {
        const int slot = atomicAdd(a.queue, 1);
        r = -1; lo = 0; hi = 0;
        if (slot < a.n_solve) { r = a.order[slot]; lo = a.indptr[r]; hi = a.indptr[r + 1]; }
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < R6_HI; ++s) { mbar_init(&mb_done[s], 1); mbar_init(&mb_full[s], R6_THREADS / 32); }
        mbar_init(&mb_acc[0], 1);
        mbar_init(&mb_acc[1], 1);
        int r0; long long lo0, hi0;
        claim_now(r0, lo0, hi0);
        rowq[0] = Row6Info{r0, (int)(hi0 - lo0), lo0};
        claim_now(rB, loB, hiB);
    }
    // tiles start as zeros: operand rows m >= LD (ld < 128) are never written and must read as zero
    for (int t = tid; t < (R6_HI + R6_LO) * R6_TILE_FLOATS / 4; t += R6_THREADS)
        reinterpret_cast<float4 *>(hi_s)[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tid < 128) p_s[tid] = 0.f;
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t idesc = idesc_tf32(LD) | (1u << 15) | (1u << 16);          // A and B MN-major
    const float wm1 = a.weight - 1.f;

    uint32_t g = 0;                           // chunks processed so far by this CTA (all rows): slot = g mod 4 / g mod 2
    uint32_t pend = 0, ph_done = 0, ph_full = 0, ph_acc = 0;
    int acc = 0;

    auto wait_done = [&](uint32_t slot) {     // the MMAs that last read hi slot `slot` (and its lo slot) have finished
        const uint32_t bit = 1u << slot;
        if (pend & bit) {
            mbar_wait_sleep(&mb_done[slot], (ph_done >> slot) & 1u);
            ph_done ^= bit;
            pend &= ~bit;
        }
    };
    // gather chunk c of the row (column indices in idx_s) into hi slot `slot`: warp w copies items 4w .. 4w+3 of the
    // chunk, lane l the 16-byte piece of m-group l.  The four destinations are the same in every slot.
    const uint32_t off0 = (uint32_t)mn_off(4 * warp, lane), off1 = (uint32_t)mn_off(4 * warp + 1, lane),
                   off2 = (uint32_t)mn_off(4 * warp + 2, lane), off3 = (uint32_t)mn_off(4 * warp + 3, lane);
    const float *const ysrc = a.Y + 4 * lane;
    auto issue_raw = [&](int c, uint32_t slot, int nnz) {
        const int base = c * CHUNK_K + 4 * warp;
        if (l_on && base < nnz) {
            const int4 it = *reinterpret_cast<const int4 *>(idx_s + ((base / IDX_BLOCK) % 3) * IDX_BLOCK + (base & (IDX_BLOCK - 1)));
            unsigned char *dst = reinterpret_cast<unsigned char *>(hi_s + slot * R6_TILE_FLOATS);
            cp_async16(dst + off0, ysrc + (size_t)((uint64_t)(uint32_t)it.x * (uint32_t)LD));
            if (base + 3 < nnz) {
                cp_async16(dst + off1, ysrc + (size_t)((uint64_t)(uint32_t)it.y * (uint32_t)LD));
                cp_async16(dst + off2, ysrc + (size_t)((uint64_t)(uint32_t)it.z * (uint32_t)LD));
                cp_async16(dst + off3, ysrc + (size_t)((uint64_t)(uint32_t)it.w * (uint32_t)LD));
            } else {
                if (base + 1 < nnz) cp_async16(dst + off1, ysrc + (size_t)((uint64_t)(uint32_t)it.y * (uint32_t)LD));
                if (base + 2 < nnz) cp_async16(dst + off2, ysrc + (size_t)((uint64_t)(uint32_t)it.z * (uint32_t)LD));
            }
        }
    };
    // the row about to start: index block 0 and warm start into shared memory (one committed group)
    auto prefetch_row = [&](const Row6Info &ri) {
        if (ri.r >= 0 && ri.nnz > 0) {
            if (tid < ri.nnz) cp_async4(idx_s + tid, a.indices + ri.lo + tid);
            if (warp == 0 && l_on) cp_async16(x0_s + 4 * lane, a.X + (size_t)ri.r * LD + 4 * lane);
        }
        cp_async_commit();
    };
    // ... and, once that block is visible to every thread, the gathers of its first two chunks
    auto prefetch_chunks = [&](const Row6Info &ri) {      // always two committed groups (possibly empty)
        const bool on = ri.r >= 0 && ri.nnz > 0;
        if (on) { wait_done(g & 3u); issue_raw(0, g & 3u, ri.nnz); }
        cp_async_commit();
        if (on && ri.nnz > CHUNK_K) { wait_done((g + 1) & 3u); issue_raw(1, (g + 1) & 3u, ri.nnz); }
        cp_async_commit();
    };
    prefetch_row(rowq[0]);
    cp_async_wait_upto(0);
    __syncthreads();
    prefetch_chunks(rowq[0]);

    for (int row_i = 0;; ++row_i) {
        const Row6Info info = rowq[row_i % 3];
        if (info.r < 0) break;
        const int nnz = info.nnz;
        float *const xr = a.X + (size_t)info.r * LD;
        if (nnz == 0) {                                                        // wmf.pyx:154-156
            if (tid < LD) xr[tid] = 0.f;
            if (tid == 0) {
                rowq[(row_i + 1) % 3] = Row6Info{rB, (int)(hiB - loB), loB};
                claim_now(rB, loB, hiB);
            }
            __syncthreads();
            prefetch_row(rowq[(row_i + 1) % 3]);
            cp_async_wait_upto(0);
            __syncthreads();
            prefetch_chunks(rowq[(row_i + 1) % 3]);
            continue;
        }
        const int32_t *const idx = a.indices + info.lo;
        const int nchunks = (nnz + CHUNK_K - 1) / CHUNK_K;
        if (tid == 0) slotA = atomicAdd(a.queue, 1);               // row i+2, step 1
        const uint32_t d_tmem = tmem_slot;
        const uint32_t t_lane = d_tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(h * HC);

        unsigned long long S2[HC / 2];
        bool s_empty = true;
        auto fold_chain = [&](int which) {
            mbar_wait_sleep(&mb_acc[which], (ph_acc >> which) & 1u);
            ph_acc ^= 1u << which;
            fence_after_sync();
            const uint32_t t0 = t_lane + (uint32_t)(which * LD);
            if (s_empty) {
                uint32_t v[HC / 16][16];
#pragma unroll
                for (int q = 0; q < HC / 16; ++q) tmem_load16_nowait(t0 + 16 * q, v[q]);
#pragma unroll
                for (int q = 0; q < HC / 16; ++q) {
                    tmem_wait_ld16(v[q]);
#pragma unroll
                    for (int t = 0; t < 8; ++t)
                        S2[8 * q + t] = pack2(__uint_as_float(v[q][2 * t]), __uint_as_float(v[q][2 * t + 1]));
                }
                s_empty = false;
            } else {
#pragma unroll
                for (int q = 0; q < HC / 16; ++q) {
                    float v[16];
                    tmem_load16(t0 + 16 * q, v);
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        float lo, hi;
                        unpack2(S2[8 * q + t], lo, hi);
                        S2[8 * q + t] = pack2(lo + v[2 * t], hi + v[2 * t + 1]);
                    }
                }
            }
            fence_before_sync();
        };

        unsigned long long bs01 = 0ull, bs23 = 0ull;              // sum over this warp's items of elements 4 lane .. + 3
        int fold = -1;
        // chunks 0 and 1 were gathered during the previous row's CG iterations (two committed groups); from here on
        // every iteration commits exactly one group (the gather of chunk c + 2, or nothing), so "all but the two
        // newest groups have landed" always means "chunk c has landed"
        for (int c = 0; c < nchunks; ++c, ++g) {
            // index blocks of long rows (three buffers), fetched eight chunks ahead and published by a CTA barrier
            if ((c & 7) == 0 && (c / 8 + 1) * IDX_BLOCK < nnz) {
                const int t = (c / 8 + 1) * IDX_BLOCK + tid;
                if (t < nnz) cp_async4(idx_s + ((c / 8 + 1) % 3) * IDX_BLOCK + tid, idx + t);
            }
            if ((c & 7) == 6 && (c / 8 + 1) * IDX_BLOCK < nnz) { cp_async_wait_upto(2); __syncthreads(); }
            // slot (g + 2) mod 4 was last read by the MMAs of chunk g - 2, which also frees lo slot g mod 2
            wait_done((g + 2) & 3u);
            if (c + 2 < nchunks) issue_raw(c + 2, (g + 2) & 3u, nnz);
            cp_async_commit();
            cp_async_wait_upto(2);                                // chunk c has landed
            float *const t_hi = hi_s + (g & 3u) * R6_TILE_FLOATS, *const t_lo = lo_s + (g & 1u) * R6_TILE_FLOATS;
            if (l_on) {
                const int left = nnz - c * CHUNK_K - 4 * warp;    // items of this warp's quad that exist
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t o = (i == 0 ? off0 : i == 1 ? off1 : i == 2 ? off2 : off3) >> 2;
                    ulonglong2 v = make_ulonglong2(0ull, 0ull);
                    if (i < left) v = *reinterpret_cast<const ulonglong2 *>(t_hi + o);
                    const ulonglong2 hi2 = make_ulonglong2(v.x & 0xffffe000ffffe000ull, v.y & 0xffffe000ffffe000ull);
                    if (!RAWHI || i >= left) *reinterpret_cast<ulonglong2 *>(t_hi + o) = hi2;
                    *reinterpret_cast<ulonglong2 *>(t_lo + o) = make_ulonglong2(sub2(v.x, hi2.x), sub2(v.y, hi2.y));
                    bs01 = add2(bs01, v.x);                       // wmf.pyx:163
                    bs23 = add2(bs23, v.y);
                }
            }
            fence_async_smem();
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&mb_full[g & 3u]);
            const bool chain_first = (c % R6_CHAIN) == 0;
            const bool chain_last = (c % R6_CHAIN) == R6_CHAIN - 1 || c == nchunks - 1;
            const uint32_t sbit = 1u << (g & 3u);
            if (warp == (int)(g & 7u)) {                          // rotating issuer, see als_tc.cu
                mbar_wait_sleep(&mb_full[g & 3u], (ph_full >> (g & 3u)) & 1u);
                if (lane == 0) {
                    fence_after_sync();
                    const int items = nnz - c * CHUNK_K < CHUNK_K ? nnz - c * CHUNK_K : CHUNK_K;
                    const int slices = (items + 7) >> 3;
                    const uint32_t d = d_tmem + (uint32_t)(acc * LD);
                    for (int ks = 0; ks < slices; ++ks) {
                        const uint64_t dh = mn_desc(t_hi + ks * 1024), dl = mn_desc(t_lo + ks * 1024);
                        mma_tf32(d, dh, dl, idesc, (chain_first && ks == 0) ? 0u : 1u);
                        mma_tf32(d, dl, dh, idesc, 1u);
                        mma_tf32(d, dh, dh, idesc, 1u);
                    }
                    mma_commit(&mb_done[g & 3u]);
                    if (chain_last) mma_commit(&mb_acc[acc]);
                }
                __syncwarp();
            }
            ph_full ^= sbit;
            pend |= sbit;
            if (fold >= 0) { fold_chain(fold); fold = -1; }
            if (chain_last) { fold = acc; acc ^= 1; }
        }
        if (tid == 0) {
            rowq[(row_i + 1) % 3] = Row6Info{rB, (int)(hiB - loB), loB};
            rA = slotA < a.n_solve ? a.order[slotA] : -1;
        }
        if (l_on) *reinterpret_cast<ulonglong2 *>(bpart + warp * 128 + 4 * lane) = make_ulonglong2(bs01, bs23);
        fold_chain(fold);

        // ---- conjugate gradient on (I + (w-1) S) x = b, S in registers (as in als_tc.cu) -----------------------------
        auto matvec = [&](const float *vec) -> float {
            unsigned long long a0 = 0ull, a1 = 0ull, a2 = 0ull, a3 = 0ull;
            const uint32_t pv = smem_u32(vec + h * HC);
            constexpr int NL = HC / 4;
            ulonglong2 u[4];
            auto lds = [&](ulonglong2 &d, int t) {
                asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(d.x), "=l"(d.y) : "r"(pv + 16u * (uint32_t)t) : "memory");
            };
#pragma unroll
            for (int t = 0; t < 4 && t < NL; ++t) lds(u[t], t);
#pragma unroll
            for (int t = 0; t < NL; ++t) {
                const ulonglong2 w = u[t & 3];
                if (t + 4 < NL) lds(u[t & 3], t + 4);
                if (t & 1) { fma2(a2, S2[2 * t], w.x); fma2(a3, S2[2 * t + 1], w.y); }
                else { fma2(a0, S2[2 * t], w.x); fma2(a1, S2[2 * t + 1], w.y); }
            }
            float s0, s1, s2, s3, s4, s5, s6, s7;
            unpack2(a0, s0, s1); unpack2(a1, s2, s3); unpack2(a2, s4, s5); unpack2(a3, s6, s7);
            return ((s0 + s1) + (s2 + s3)) + ((s4 + s5) + (s6 + s7));
        };
        auto dot4 = [&](const float4 &u, const float4 &v) -> float {
            return warp_sum6(fmaf(u.x, v.x, fmaf(u.y, v.y, fmaf(u.z, v.z, u.w * v.w))));
        };
        auto apply = [&](const float4 &v, const float *pp) -> float4 {
            const float4 s0 = *reinterpret_cast<const float4 *>(pp + 4 * lane);
            const float4 s1 = *reinterpret_cast<const float4 *>(pp + 128 + 4 * lane);
            return make_float4(fmaf(wm1, s0.x + s1.x, v.x), fmaf(wm1, s0.y + s1.y, v.y), fmaf(wm1, s0.z + s1.z, v.z),
                               fmaf(wm1, s0.w + s1.w, v.w));
        };
        part[h * 128 + m] = matvec(x0_s);                         // S x0 for the warm start
        __syncthreads();                                          // bpart, S x0 and rowq[i+1] are visible
        float4 x4 = make_float4(0.f, 0.f, 0.f, 0.f), b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (l_on) {
            x4 = *reinterpret_cast<const float4 *>(x0_s + 4 * lane);
#pragma unroll
            for (int w = 0; w < R6_THREADS / 32; ++w) {
                const float4 u = *reinterpret_cast<const float4 *>(bpart + w * 128 + 4 * lane);
                b4.x += u.x; b4.y += u.y; b4.z += u.z; b4.w += u.w;
            }
            b4.x *= a.weight; b4.y *= a.weight; b4.z *= a.weight; b4.w *= a.weight;
        }
        const float bb = dot4(b4, b4);
        const float4 ax = apply(x4, part);
        __syncthreads();                                          // x0_s, idx_s and bpart may now be refilled
        const Row6Info next = rowq[(row_i + 1) % 3];
        prefetch_row(next);                                       // index block 0 + warm start of the next row
        bool chunks_prefetched = false;
        unsigned iters = 0;
        bool stalled = false;
        if (bb > 0.f) {
            float4 r4 = make_float4(b4.x - ax.x, b4.y - ax.y, b4.z - ax.z, b4.w - ax.w);
            float rs = dot4(r4, r4);
            float4 p4 = r4;
            const float stop = a.tol2 * bb;
            int buf = 1;
            while (rs > stop) {
                if ((int)iters >= a.max_iter) { stalled = true; break; }
                *reinterpret_cast<float4 *>(p_s + 4 * lane) = p4;
                __syncwarp();
                float *const pp = part + buf * 256;
                pp[h * 128 + m] = matvec(p_s);
                const bool pf = iters == 1 && !chunks_prefetched;             // the next row's index block has landed by now
                if (pf) cp_async_wait_upto(0);
                __syncthreads();
                if (pf) { prefetch_chunks(next); chunks_prefetched = true; }  // its first two gathers fly under the iterations
                const float4 ap = apply(p4, pp);
                const float pAp = dot4(p4, ap);
                if (!(pAp > 0.f)) { stalled = true; break; }
                const float alpha = rs * rcp_approx(pAp);
                x4 = make_float4(fmaf(alpha, p4.x, x4.x), fmaf(alpha, p4.y, x4.y), fmaf(alpha, p4.z, x4.z), fmaf(alpha, p4.w, x4.w));
                r4 = make_float4(fmaf(-alpha, ap.x, r4.x), fmaf(-alpha, ap.y, r4.y), fmaf(-alpha, ap.z, r4.z), fmaf(-alpha, ap.w, r4.w));
                const float rs_new = dot4(r4, r4);
                const float beta = rs_new * rcp_approx(rs);
                p4 = make_float4(fmaf(beta, p4.x, r4.x), fmaf(beta, p4.y, r4.y), fmaf(beta, p4.z, r4.z), fmaf(beta, p4.w, r4.w));
                rs = rs_new;
                ++iters;
                buf ^= 1;
            }
        } else {
            x4 = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (warp == 0 && l_on) *reinterpret_cast<float4 *>(xr + 4 * lane) = x4;
        if (!chunks_prefetched) {                                 // CG ended early: publish the index block now
            cp_async_wait_upto(0);
            __syncthreads();
            prefetch_chunks(next);
        }
        if (tid == 0) {
            rB = rA;
            loB = hiB = 0;
            if (rA >= 0) { loB = a.indptr[rA]; hiB = a.indptr[rA + 1]; }
            if (a.stats) {
                atomicAdd(a.stats, (unsigned long long)iters);
                if (stalled) atomicAdd(a.stats + 1, 1ull);
            }
        }
    }
    cp_async_wait_upto(0);
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_slot, tmem_cols);
}

template <int LD, bool RAWHI> static int launch_rows6(const Row6Args &a, cudaStream_t st) {
    const size_t smem = sizeof(float) * ((size_t)(R6_HI + R6_LO) * R6_TILE_FLOATS + 128 + 512 + 1024 + 128 + 3 * R6_THREADS);
    auto kern = als_rows_tc6_kernel<LD, RAWHI>;
    CYMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CYMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int64_t blocks = (int64_t)sm_count() * 2;
    if (blocks > a.n_solve) blocks = a.n_solve;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, R6_THREADS, smem, st>>>(a);
    CYMF_LAUNCHED();
    return 0;
}

}  // namespace tc

int tc_als_rows6(const int64_t *indptr, const int32_t *indices, const int32_t *order, int32_t n_solve, float *X,
                 const float *Y, int ld, float weight, float tol2, int32_t max_iter, int32_t *queue,
                 unsigned long long *stats, cudaStream_t st) {
    CYMF_CUDA(cudaMemsetAsync(queue, 0, sizeof(int32_t), st));
    tc::Row6Args a{indptr, indices, order, n_solve, X, Y, max_iter, weight, tol2, queue, stats};
    const char *v = getenv("CYMF_ALS_RAWHI");
    const bool rawhi = !(v && v[0] == '0');
    switch (ld) {
        case 32: return rawhi ? tc::launch_rows6<32, true>(a, st) : tc::launch_rows6<32, false>(a, st);
        case 64: return rawhi ? tc::launch_rows6<64, true>(a, st) : tc::launch_rows6<64, false>(a, st);
        case 96: return rawhi ? tc::launch_rows6<96, true>(a, st) : tc::launch_rows6<96, false>(a, st);
        case 128: return rawhi ? tc::launch_rows6<128, true>(a, st) : tc::launch_rows6<128, false>(a, st);
    }
    set_error("als rows (tensor cores): ld must be 32, 64, 96 or 128");
    return CYMF_EUNSUPPORTED;
}

}  // namespace cymf
