// als_ws.cu -- warp-specialised, persistent form of the one-pass WMF row solver (the prange body of WMF._als,
// cymf/wmf.pyx:150-168; f32 factors in the transformed coordinates of cymf_chol_transforms_dev, ld in {32,64,96,128}).
//
// Same mathematics as als_tc.cu: per row S = sum_{i in row} y~_i y~_i^T (the matrix the reference accumulates entry by
// entry, wmf.pyx:161-166) is built once on the tensor cores (3xTF32, TMEM accumulators, chains of 512 items folded in
// registers), then (I + (w-1) S) x~ = w sum y~_i (wmf.pyx:163,168) is solved by conjugate gradient out of registers.
// als_tc.cu runs those phases one after the other in a CTA (gather -> MMA -> fold -> ~8 CG iterations), two CTAs per
// SM: the tensor pipe is 27-41 % busy and every phase waits on the latency of the previous one.  Here ONE CTA per SM
// keeps all of them running at once on different rows:
//   * warps 13-14 ("copy"): stream the 32-item chunks of the CTA's rows, row after row, into a ring of ten 16 KB "hi"
//     slots as soon as a slot is free.  Either cp.async (LDGSTS.128: one warp instruction drops a 512-byte item vector
//     straight into the MN-major SWIZZLE_128B_BASE32B tile of als_tc6.cu -- an item vector IS a contiguous run of the MN
//     dimension -- and the completions arrive on the slot's mbarrier by themselves, cp.async.mbarrier.arrive.noinc), or
//     the TMA unit (cp.async.bulk.tensor tile::gather4 over a tensor map of Y whose swizzle mode is the operand's own
//     pattern: four item vectors per copy, complete_tx bytes on the same mbarrier).  These warps never wait for data and
//     never execute a proxy fence (which would wait for their copies in flight: measured, that alone cost the first
//     version of this kernel its look-ahead);
//   * warps 8-11 ("convert"): when a chunk has landed, produce lo = a - hi into one of three 16 KB "lo" slots (the raw
//     data is the hi operand: the tensor core ignores the 13 low mantissa bits), add the items into sum y~
//     (wmf.pyx:163), fence.proxy.async, hand the stage to the MMA warp;
//   * warp 12 ("mma"): one ELECTED lane (elect.sync -- with `lane == 0` the compiler wraps every tcgen05.mma in a
//     register-to-uniform broadcast loop) issues the twelve tcgen05.mma of a chunk as soon as its stage is full, chain
//     after chain into the FOUR 128-column TMEM accumulators, and hands the slots back through tcgen05.commit;
//   * warps 0-3 and 4-7 (two "solver" groups, thread = TMEM lane = row of S, 192 registers after setmaxnreg): take
//     the CTA's rows alternately; each group owns two accumulators and two sum-y~ slots (so that every mbarrier has ONE
//     waiter that sees each of its phases: a waiter that skipped a use would read the parity of the wrong phase); fold
//     the row's chains into registers as they complete (freeing the accumulator), then run CG (Chronopoulos-Gear form:
//     one reduction round per iteration) with a 128-thread named barrier.  While one group iterates, the other folds /
//     iterates on the next row and the copy / convert / MMA warps are several rows ahead.
// Rows are assigned to CTAs on the host (cymf_als_ws_schedule_host: longest rows first to the least loaded CTA; inside
// a CTA's list long and short rows alternate), so every role walks the same static list and no role ever has to tell
// another what comes next: all hand-overs are mbarrier phases whose parity follows from counters each role keeps for
// itself.  A hand-over that does not complete within ~2 s is recorded in `debug` and the kernel ends instead of hanging.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "tc_common.cuh"

namespace cymf {
namespace tc {

constexpr int WS_THREADS = 512;       // 4 warpgroups: solver 0, solver 1, convert, {mma, copy, copy, idle}
constexpr int WS_NHI = 10;            // raw / hi operand slots (16 KB each): chunk t lives in slot t mod 10
constexpr int WS_NLO = 3;             // lo operand slots (16 KB each): chunk t's lo tile lives in slot t mod 3
constexpr int WS_NI = 2;              // copy-issuing warps (16 items of a chunk each)
constexpr int WS_CHAIN = 16;          // 32-item chunks per accumulator chain
constexpr int WS_NACC = 4;            // TMEM accumulators (128 columns each)
constexpr int WS_NB = 4;              // row slots for sum y~ (gather -> solver)
constexpr int WS_NG = 4;              // converter warps (8 items of a chunk each)
constexpr int WS_TILE = TILE_M * CHUNK_K;     // floats per 16 KB tile

struct WsArgs {
    const int4 *rowinfo;        // per CTA, consecutive: (row, nnz, indptr low word, indptr high word)
    const int32_t *cta_ptr;     // [gridDim.x + 1]
    const int32_t *indices;
    float *X;                   // [rows, ld]  x~ rows, solved in place (warm start = current content)
    const float *Y;             // [n, ld]     y~ rows
    int32_t max_iter;
    float weight, tol2;
    unsigned long long *stats;  // [0] CG iterations summed over rows, [1] rows that hit max_iter (may be NULL)
    unsigned long long *debug;  // [8] first timed-out wait (may be NULL): count, CTA, site, parity, three counters, thread
};

// MN-major SWIZZLE_128B_BASE32B tile of 128 (MN) x 32 (K) f32, see als_tc6.cu: byte offset of the 16-byte piece holding
// elements m = 4 g .. 4 g + 3 of reduction index k
__device__ __forceinline__ int ws_off(int k, int g) {
    return (k >> 3) * 4096 + (g >> 3) * 1024 + (k & 7) * 128 + (((((g & 7) >> 1) ^ (k & 3))) << 5) + ((g & 1) << 4);
}
__device__ __forceinline__ uint64_t ws_desc(const void *p) {
    return (uint64_t)((smem_u32(p) >> 4) & 0x3fffu) | ((uint64_t)(1024u >> 4) << 16) | ((uint64_t)(512u >> 4) << 32) |
           (1ull << 46) | (1ull << 61);
}
__device__ __forceinline__ void ws_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void ws_wait_pending(int n) {      // at most n committed groups still in flight
    switch (n) {
        case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
        case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
        default: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    }
}
// Hand-over waits.  mbarrier.try_wait already suspends the thread for a short, hardware-chosen time; an explicit
// suspend-time hint compiles to NANOSLEEP loops whose wake-up latency (microseconds) is longer than a whole chunk and
// starves the pipeline (measured), so there is none.  A wait that does not complete within ~2 s is a protocol
// failure: it is recorded in a.debug (first failure only: CTA, wait site, the waiter's counters), the CTA-wide abort
// flag makes every later wait fall through, and the kernel ends with stats[1] poisoned instead of hanging the device.
struct WsWaitCtx {
    volatile int *abort_flag;
    unsigned long long *debug;
};
__device__ __forceinline__ bool ws_try(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ unsigned long long ws_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void ws_wait(const WsWaitCtx &cx, uint64_t *bar, uint32_t parity, int site, uint32_t c0 = 0, uint32_t c1 = 0,
                                        uint32_t c2 = 0) {
    uint32_t spins = 0;
    unsigned long long t0 = 0;
    while (!ws_try(bar, parity)) {
        if ((++spins & 1023u) == 0u) {
            const unsigned long long now = ws_now();
            if (t0 == 0) t0 = now;
            if (now - t0 > 2000000000ull || *cx.abort_flag) {
                *cx.abort_flag = 1;
                if (cx.debug && atomicAdd(cx.debug, 1ull) == 0ull) {
                    cx.debug[1] = blockIdx.x;
                    cx.debug[2] = (unsigned long long)site;
                    cx.debug[3] = parity;
                    cx.debug[4] = c0;
                    cx.debug[5] = c1;
                    cx.debug[6] = c2;
                    cx.debug[7] = threadIdx.x;
                }
                return;
            }
        }
    }
}
__device__ __forceinline__ unsigned long long ws_sub2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long ws_add2(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ float ws_warp_sum(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
__device__ __forceinline__ long long ws_lo64(const int4 &ri) {
    return (long long)(((unsigned long long)(uint32_t)ri.w << 32) | (unsigned long long)(uint32_t)ri.z);
}

struct WsShared {
    uint64_t landed[WS_NHI];      // copy -> convert: the chunk's item vectors are in the hi slot (cp.async completions)
    uint64_t full[WS_NHI];        // convert -> mma: all four converter warps have produced their part of the lo tile
    uint64_t done_hi[WS_NHI];     // mma -> copy: the MMAs that read the hi slot have completed
    uint64_t done_lo[WS_NLO];     // mma -> convert: the MMAs that read the lo slot have completed
    uint64_t acc_full[WS_NACC];   // mma -> solver: the chain in this accumulator has completed
    uint64_t acc_empty[WS_NACC];  // solver -> mma: the chain has been folded into registers
    uint64_t b_full[WS_NB];       // gather -> solver: sum y~ of the row is in its slot
    uint64_t b_free[WS_NB];       // solver -> gather: the slot has been read
    uint32_t tmem;
    int abort_flag;
};

// 4 rows x 128 bytes of the 2-D tensor `map` (rows r0..r3, columns c0 .. c0 + 31) -> 512 bytes of shared memory in the
// map's swizzle pattern; completion is counted in bytes on `bar` (TMA tile::gather4, sm_100)
__device__ __forceinline__ void tma_gather4(void *dst, const CUtensorMap *map, int32_t c0, int32_t r0, int32_t r1, int32_t r2,
                                            int32_t r3, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

template <int LD, bool TMA>
__global__ void __launch_bounds__(WS_THREADS, 1) als_rows_ws_kernel(const WsArgs a, const __grid_constant__ CUtensorMap ymap) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float *const hi_s = reinterpret_cast<float *>(smem_raw);                  // [WS_NHI][4096]
    float *const lo_s = hi_s + WS_NHI * WS_TILE;                              // [WS_NLO][4096]
    float *const bq = lo_s + WS_NLO * WS_TILE;                                // [WS_NB][WS_NG][128] partial sums of y~
    float *const p_s = bq + WS_NB * WS_NG * 128;                              // [2][128] CG direction, per solver group
    float *const red_s = p_s + 256;                                           // [2][2][4] reduction partials
    __shared__ WsShared sh;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int row0 = a.cta_ptr[blockIdx.x], n_rows = a.cta_ptr[blockIdx.x + 1] - row0;
    const int4 *const rows = a.rowinfo + row0;

    const WsWaitCtx cx{&sh.abort_flag, a.debug};
    if (tid == 0) {
        sh.abort_flag = 0;
        for (int s = 0; s < WS_NHI; ++s) {
            mbar_init(&sh.landed[s], TMA ? 1 : 32 * WS_NI);
            mbar_init(&sh.full[s], WS_NG);
            mbar_init(&sh.done_hi[s], 1);
        }
        for (int s = 0; s < WS_NLO; ++s) mbar_init(&sh.done_lo[s], 1);
        for (int s = 0; s < WS_NACC; ++s) { mbar_init(&sh.acc_full[s], 1); mbar_init(&sh.acc_empty[s], 4); }
        for (int s = 0; s < WS_NB; ++s) { mbar_init(&sh.b_full[s], WS_NG); mbar_init(&sh.b_free[s], 4); }
    }
    if (warp == 12) tmem_alloc(&sh.tmem, 512);
    // tiles start as zeros: operand rows m >= LD (ld < 128) are never written and must read as zero
    for (int t = tid; t < (WS_NHI + WS_NLO) * WS_TILE / 4; t += WS_THREADS)
        reinterpret_cast<float4 *>(hi_s)[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem0 = sh.tmem;
    // cycle accounting of the roles of one CTA (the middle one) into debug[8..24), see tools/als_ws_prof.py
    const bool prof = a.debug != nullptr && blockIdx.x == gridDim.x / 2;
    const bool dbg_on = prof && lane == 0;
    auto tick = [&]() -> uint32_t { return prof ? (uint32_t)clock() : 0u; };      // differences wrap correctly
#define WS_DBG(idx, val) do { if (dbg_on) a.debug[idx] = (unsigned long long)(val); } while (0)

    if (warp < 8) {
        // ================================ solver groups ======================================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 192;" ::: "memory");
        const int grp = warp >> 2, m = tid & 127, wq = warp & 3;
        const bool m_on = m < LD;
        const float wm1 = a.weight - 1.f;
        float *const pg = p_s + grp * 128;
        float *const rg = red_s + grp * 8;
        auto gsync = [&]() { asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory"); };
        // Each group owns two of the four accumulators and two of the four b slots (2 grp, 2 grp + 1) and uses them
        // alternately, so every mbarrier has ONE waiting role that sees each of its phases in turn (a waiter that
        // skipped a use would read the parity of the wrong phase).
        uint32_t chain = 0, nz = 0;                              // chains / non-empty rows of THIS GROUP so far
        uint32_t t_wait = 0, t_fold = 0, t_cg = 0, n_mine = 0, t_mv = 0, t_red = 0, t_pub = 0, n_it = 0;
        // the group's next row and its warm start are fetched one row ahead (the X rows come from DRAM)
        int4 ri_next = grp < n_rows ? rows[grp] : make_int4(0, 0, 0, 0);
        float x0_next = (grp < n_rows && m_on) ? a.X[(size_t)ri_next.x * LD + m] : 0.f;
        for (int i = grp; i < n_rows; i += 2) {
            const int4 ri = ri_next;
            const float x0 = x0_next;                            // warm start
            if (i + 2 < n_rows) {
                ri_next = rows[i + 2];
                x0_next = m_on ? a.X[(size_t)ri_next.x * LD + m] : 0.f;
            }
            const int nnz = ri.y;
            float *const xr = a.X + (size_t)ri.x * LD;
            if (nnz == 0) {                                      // wmf.pyx:154-156
                if (m_on) xr[m] = 0.f;
                continue;
            }
            const int nchains = (nnz + 32 * WS_CHAIN - 1) / (32 * WS_CHAIN);
            unsigned long long S2[LD / 2];                       // row m of S, packed pairs
            for (int c = 0; c < nchains; ++c, ++chain) {
                const uint32_t acc = 2u * grp + (chain & 1u), use = chain >> 1;
                const uint32_t k0 = tick();
                ws_wait(cx, &sh.acc_full[acc], use & 1u, 1, chain, (uint32_t)i, (uint32_t)c);
                const uint32_t k1 = tick();
                t_wait += k1 - k0;
                fence_after_sync();
                const uint32_t t0 = tmem0 + ((uint32_t)(wq * 32) << 16) + acc * 128u;
                if (c == 0) {
#pragma unroll
                    for (int q = 0; q < LD / 16; ++q) {
                        float v[16];
                        tmem_load16(t0 + 16 * q, v);
#pragma unroll
                        for (int t = 0; t < 8; ++t) S2[8 * q + t] = pack2(v[2 * t], v[2 * t + 1]);
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < LD / 16; ++q) {
                        float v[16];
                        tmem_load16(t0 + 16 * q, v);
#pragma unroll
                        for (int t = 0; t < 8; ++t) S2[8 * q + t] = ws_add2(S2[8 * q + t], pack2(v[2 * t], v[2 * t + 1]));
                    }
                }
                fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(&sh.acc_empty[acc]);  // this warp's quarter of the accumulator is in registers
                t_fold += tick() - k1;
            }
            // b = w sum y~ (four partial sums, one per gather warp, added in a fixed order)
            const uint32_t bs = 2u * grp + (nz & 1u), buse = nz >> 1;
            ++nz;
            const uint32_t k2 = tick();
            ws_wait(cx, &sh.b_full[bs], buse & 1u, 2, nz, (uint32_t)i, chain);
            float b = 0.f;
            if (m_on) {
                const float *bp = bq + bs * (WS_NG * 128) + m;
                b = a.weight * ((bp[0] + bp[128]) + (bp[256] + bp[384]));
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.b_free[bs]);

            // ---- conjugate gradient on (I + (w-1) S) x = b, row m of S in this thread's registers -----------------
            auto matvec = [&]() -> float {                       // (S p)_m, p in shared memory
                unsigned long long a0 = 0ull, a1 = 0ull, a2 = 0ull, a3 = 0ull;
                const uint32_t pv = smem_u32(pg);
                constexpr int NL = LD / 4;
                ulonglong2 u[4];                                 // four loads in flight (eight: no faster, measured, and spills)
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
            // Chronopoulos-Gear form of CG: the two inner products of an iteration, (r, r) and (r, A r), are taken
            // together, so an iteration costs ONE reduction round (two interleaved shuffle butterflies + one
            // 128-thread barrier) and one barrier for publishing r, instead of two rounds + one barrier: the rounds,
            // not the matvec, are what an iteration waits for (measured: 43 % of the loop on the shuffle chains).
            auto gsum2 = [&](float u, float v, float &su, float &sv) {       // sums over the group, the same in every thread
#pragma unroll
                for (int off = 16; off > 0; off >>= 1) {
                    u += __shfl_xor_sync(0xffffffffu, u, off);
                    v += __shfl_xor_sync(0xffffffffu, v, off);
                }
                if (lane == 0) *reinterpret_cast<float2 *>(rg + 2 * wq) = make_float2(u, v);
                gsync();
                const float4 q0 = *reinterpret_cast<const float4 *>(rg), q1 = *reinterpret_cast<const float4 *>(rg + 4);
                su = (q0.x + q0.z) + (q1.x + q1.z);
                sv = (q0.y + q0.w) + (q1.y + q1.w);
            };
            pg[m] = x0;
            gsync();
            float x = x0;
            float r = b - fmaf(wm1, matvec(), x0);               // r0 = b - A x0
            if (!m_on) r = 0.f;
            float bb, gam;
            gsum2(b * b, r * r, bb, gam);
            unsigned iters = 0;
            bool stalled = false;
            if (bb > 0.f) {
                const float stop = a.tol2 * bb;
                if (gam > stop) {
                    pg[m] = r;
                    gsync();
                    float w = fmaf(wm1, matvec(), r);            // w = A r
                    if (!m_on) w = 0.f;
                    float dlt, unused;
                    gsum2(r * w, 0.f, dlt, unused);
                    float p = r, sv = w;                         // s = A p
                    float alpha = gam * rcp_approx(dlt);
                    if (!(dlt > 0.f)) { stalled = true; alpha = 0.f; }
                    while (!stalled) {
                        x = fmaf(alpha, p, x);
                        r = fmaf(-alpha, sv, r);
                        ++iters;
                        const uint32_t q0 = tick();
                        pg[m] = r;
                        gsync();
                        const uint32_t q1 = tick();
                        w = fmaf(wm1, matvec(), r);
                        if (!m_on) w = 0.f;
                        const uint32_t q2 = tick();
                        float gam_new;
                        gsum2(r * r, r * w, gam_new, dlt);
                        t_pub += q1 - q0; t_mv += q2 - q1; t_red += tick() - q2; ++n_it;
                        if (!(gam_new > stop)) break;            // converged (or NaN: leave)
                        if ((int)iters >= a.max_iter) { stalled = true; break; }
                        const float beta = gam_new * rcp_approx(gam);
                        const float den = dlt - beta * gam_new * rcp_approx(alpha);
                        if (!(den > 0.f)) { stalled = true; break; }
                        alpha = gam_new * rcp_approx(den);
                        p = fmaf(beta, p, r);
                        sv = fmaf(beta, sv, w);
                        gam = gam_new;
                    }
                }
            } else {
                x = 0.f;                                         // b = 0  =>  x = 0
            }
            if (m_on) xr[m] = x;
            if (m == 0 && a.stats) {
                atomicAdd(a.stats, (unsigned long long)iters);
                if (stalled) atomicAdd(a.stats + 1, 1ull);
            }
            gsync();                                             // pg / rg are reused by the group's next row
            t_cg += tick() - k2;
            ++n_mine;
        }
        if (wq == 0) { WS_DBG(8 + 4 * grp, t_wait); WS_DBG(9 + 4 * grp, t_fold); WS_DBG(10 + 4 * grp, t_cg); WS_DBG(11 + 4 * grp, n_mine); }
        if (wq == 0 && grp == 0) { WS_DBG(26, t_pub); WS_DBG(27, t_mv); WS_DBG(28, t_red); WS_DBG(29, n_it); }
    } else {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;" ::: "memory");
        if (warp < 8 + WS_NG) {
            // ================================ converter warps ================================================
            // chunk t: wait until its vectors have landed in hi slot t mod 8 and lo slot t mod 3 is free, produce
            // lo = a - hi for this warp's eight items, add them into sum y~, hand the stage to the MMA warp.  These
            // warps never have a global load in flight, so the proxy fence (a full memory barrier) costs little.
            const int w = warp - 8;
            const bool l_on = 4 * lane < LD;
            // float offset of (item 8 w + j, this lane's 16-byte piece) = offq[j & 3] + 32 j: four registers instead of eight
            uint32_t offq[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) offq[q] = (uint32_t)ws_off(8 * w + q, lane) >> 2;
            uint32_t t = 0, nzg[2] = {0u, 0u};                   // chunks so far; non-empty rows so far per solver group
            unsigned long long bs01 = 0ull, bs23 = 0ull;         // sum over this warp's items of elements 4 lane .. + 3
            uint32_t t_lo = 0, t_land = 0, t_bfree = 0;
            const uint32_t g_start = tick();
            for (int i = 0; i < n_rows; ++i) {
                const int nnz = rows[i].y;
                const int nchunks = (nnz + 31) >> 5;
                for (int c = 0; c < nchunks; ++c, ++t) {
                    const uint32_t hs = t % WS_NHI, ls = t % WS_NLO;
                    const uint32_t k0 = tick();
                    if (t >= (uint32_t)WS_NLO) ws_wait(cx, &sh.done_lo[ls], (t / WS_NLO - 1) & 1u, 3, t, (uint32_t)i, (uint32_t)c);
                    const uint32_t k1 = tick();
                    ws_wait(cx, &sh.landed[hs], (t / WS_NHI) & 1u, 7, t, (uint32_t)i, (uint32_t)c);
                    t_lo += k1 - k0;
                    t_land += tick() - k1;
                    float *const t_hi = hi_s + hs * WS_TILE, *const t_lw = lo_s + ls * WS_TILE;
                    const int left = nnz - 32 * c - 8 * w;       // items of this warp's eight that exist
                    if (l_on) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {            // two batches of four items: four loads in flight
                            ulonglong2 v[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                v[q] = make_ulonglong2(0ull, 0ull);
                                if (4 * h + q < left) v[q] = *reinterpret_cast<const ulonglong2 *>(t_hi + offq[q] + 128 * h);
                            }
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                const uint32_t o = offq[q] + 128 * h;
                                if (4 * h + q >= left) *reinterpret_cast<ulonglong2 *>(t_hi + o) = v[q];     // past the end of the row: zeros
                                const ulonglong2 h2 = make_ulonglong2(v[q].x & 0xffffe000ffffe000ull, v[q].y & 0xffffe000ffffe000ull);
                                *reinterpret_cast<ulonglong2 *>(t_lw + o) = make_ulonglong2(ws_sub2(v[q].x, h2.x), ws_sub2(v[q].y, h2.y));
                                bs01 = ws_add2(bs01, v[q].x);    // wmf.pyx:163
                                bs23 = ws_add2(bs23, v[q].y);
                            }
                        }
                    }
                    if (c == nchunks - 1) {                      // the row's sum goes to its solver group
                        const int og = i & 1;
                        const uint32_t nz = nzg[og], bs = 2u * og + (nz & 1u);
                        const uint32_t k2 = tick();
                        if (nz >= 2u) ws_wait(cx, &sh.b_free[bs], ((nz >> 1) - 1) & 1u, 4, nz, t, (uint32_t)i);
                        t_bfree += tick() - k2;
                        if (l_on) *reinterpret_cast<ulonglong2 *>(bq + bs * (WS_NG * 128) + w * 128 + 4 * lane) = make_ulonglong2(bs01, bs23);
                        bs01 = bs23 = 0ull;
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&sh.b_full[bs]);
                        ++nzg[og];
                    }
                    fence_async_smem();                          // generic-proxy writes (cp.async, st.shared) -> tensor core
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&sh.full[hs]);
                }
            }
            if (w == 0) { WS_DBG(16, t_lo); WS_DBG(17, t_land); WS_DBG(18, t_bfree); WS_DBG(19, tick() - g_start); }
        } else if (warp >= 13 && warp < 13 + WS_NI) {
            // ================================ copy-issuing warps =============================================
            // chunk t: as soon as hi slot t mod 8 is free, one cp.async warp instruction per item drops the 512-byte
            // vector into its swizzled rows; the completions arrive on the slot's mbarrier by themselves
            // (cp.async.mbarrier.arrive.noinc), so these warps never wait for data: up to eight chunks are in flight.
            const int w = warp - 13;
            if constexpr (TMA) {
                // TMA form: the two warps take the chunks alternately; one elected lane issues, per group of four items
                // and 32-column block, ONE tile::gather4 copy (4 x 128 bytes, written in the operand's
                // SWIZZLE_128B_BASE32B pattern by the TMA unit itself).  Bulk copies do not pass through the L1's
                // miss tracking, whose capacity (~20 KB in flight per SM, measured) is what bounds the cp.async form.
                uint32_t t = 0;
                uint32_t t_free = 0;
                const uint32_t c_start = tick();
                int4 nxt = n_rows > 0 ? rows[0] : make_int4(0, 0, 0, 0);
                int32_t idx = 0;                                 // lane l: index of item l of the chunk about to be copied
                {
                    int i2 = 0;
                    int4 r2 = nxt;
                    while (i2 < n_rows && r2.y == 0) { ++i2; if (i2 < n_rows) r2 = rows[i2]; }
                    if (i2 < n_rows && lane < r2.y) idx = __ldg(a.indices + ws_lo64(r2) + lane);
                }
                for (int i = 0; i < n_rows; ++i) {
                    const int4 ri = nxt;
                    if (i + 1 < n_rows) nxt = rows[i + 1];
                    const int nnz = ri.y;
                    const long long lo = ws_lo64(ri);
                    const int nchunks = (nnz + 31) >> 5;
                    for (int c = 0; c < nchunks; ++c, ++t) {
                        if ((int)(t & 1u) == w) {
                            const uint32_t hs = t % WS_NHI;
                            const uint32_t k0 = tick();
                            if (t >= (uint32_t)WS_NHI) ws_wait(cx, &sh.done_hi[hs], (t / WS_NHI - 1) & 1u, 8, t, (uint32_t)i, (uint32_t)c);
                            t_free += tick() - k0;
                            const int items = nnz - 32 * c < 32 ? nnz - 32 * c : 32;
                            const int groups = (items + 3) >> 2;
                            unsigned char *dst = reinterpret_cast<unsigned char *>(hi_s + hs * WS_TILE);
                            const bool leader = elect_one();
                            if (leader) mbar_expect_tx(&sh.landed[hs], (uint32_t)(groups * (LD / 32) * 512));
#pragma unroll
                            for (int q = 0; q < 8; ++q) {
                                const int32_t r0 = __shfl_sync(0xffffffffu, idx, 4 * q);
                                int32_t r1 = __shfl_sync(0xffffffffu, idx, 4 * q + 1);
                                int32_t r2 = __shfl_sync(0xffffffffu, idx, 4 * q + 2);
                                int32_t r3 = __shfl_sync(0xffffffffu, idx, 4 * q + 3);
                                if (4 * q + 1 >= items) r1 = r0;          // past the end of the row: any valid row (the converter zeroes it)
                                if (4 * q + 2 >= items) r2 = r0;
                                if (4 * q + 3 >= items) r3 = r0;
                                if (leader && q < groups) {
#pragma unroll
                                    for (int mb = 0; mb < LD / 32; ++mb)
                                        tma_gather4(dst + (q >> 1) * 4096 + mb * 1024 + (q & 1) * 512, &ymap, 32 * mb, r0, r1, r2, r3,
                                                    &sh.landed[hs]);
                                }
                            }
                            __syncwarp();
                        }
                        // indices of the next chunk (of this row, or of the next non-empty row)
                        idx = 0;
                        if (c + 1 < nchunks) {
                            const int e = 32 * (c + 1) + lane;
                            if (e < nnz) idx = __ldg(a.indices + lo + e);
                        } else {
                            int i2 = i + 1;
                            int4 r2 = nxt;
                            while (i2 < n_rows && r2.y == 0) { ++i2; if (i2 < n_rows) r2 = rows[i2]; }
                            if (i2 < n_rows && lane < r2.y) idx = __ldg(a.indices + ws_lo64(r2) + lane);
                        }
                    }
                }
                if (w == 0) { WS_DBG(24, t_free); WS_DBG(25, tick() - c_start); }
            } else {
            const bool l_on = 4 * lane < LD;
            const float *const ysrc = a.Y + 4 * lane;
            uint32_t t = 0;
            uint32_t t_free = 0;
            const uint32_t c_start = tick();
            int4 nxt = n_rows > 0 ? rows[0] : make_int4(0, 0, 0, 0);
            int32_t idx = 0;                                     // lanes 0-15: this warp's item indices of the chunk about to be copied
            {
                int i2 = 0;
                int4 r2 = nxt;
                while (i2 < n_rows && r2.y == 0) { ++i2; if (i2 < n_rows) r2 = rows[i2]; }
                const int e = 16 * w + lane;
                if (i2 < n_rows && lane < 16 && e < r2.y) idx = __ldg(a.indices + ws_lo64(r2) + e);
            }
            for (int i = 0; i < n_rows; ++i) {
                const int4 ri = nxt;
                if (i + 1 < n_rows) nxt = rows[i + 1];
                const int nnz = ri.y;
                const long long lo = ws_lo64(ri);
                const int nchunks = (nnz + 31) >> 5;
                for (int c = 0; c < nchunks; ++c, ++t) {
                    const uint32_t hs = t % WS_NHI;
                    const uint32_t k0 = tick();
                    if (t >= (uint32_t)WS_NHI) ws_wait(cx, &sh.done_hi[hs], (t / WS_NHI - 1) & 1u, 8, t, (uint32_t)i, (uint32_t)c);
                    t_free += tick() - k0;
                    const int left = nnz - 32 * c - 16 * w;
                    unsigned char *dst = reinterpret_cast<unsigned char *>(hi_s + hs * WS_TILE);
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int32_t it = __shfl_sync(0xffffffffu, idx, j);
                        if (l_on && j < left)
                            cp_async16(dst + ws_off(16 * w + j, lane), ysrc + (size_t)((uint64_t)(uint32_t)it * (uint32_t)LD));
                    }
                    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&sh.landed[hs])) : "memory");
                    // indices of the next chunk (of this row, or of the next non-empty row)
                    idx = 0;
                    if (c + 1 < nchunks) {
                        const int e = 32 * (c + 1) + 16 * w + lane;
                        if (lane < 16 && e < nnz) idx = __ldg(a.indices + lo + e);
                    } else {
                        int i2 = i + 1;
                        int4 r2 = nxt;
                        while (i2 < n_rows && r2.y == 0) { ++i2; if (i2 < n_rows) r2 = rows[i2]; }
                        const int e = 16 * w + lane;
                        if (i2 < n_rows && lane < 16 && e < r2.y) idx = __ldg(a.indices + ws_lo64(r2) + e);
                    }
                }
            }
            if (w == 0) { WS_DBG(24, t_free); WS_DBG(25, tick() - c_start); }
            }
        } else if (warp == 12) {
            // ================================ MMA issuer ======================================================
            const uint32_t idesc = idesc_tf32(LD) | (1u << 15) | (1u << 16);          // A and B MN-major
            uint32_t t = 0, chg[2] = {0u, 0u};                   // chunks so far; chains so far per solver group
            uint32_t t_empty = 0, t_full = 0;
            const uint32_t m_start = tick();
            for (int i = 0; i < n_rows; ++i) {
                const int og = i & 1;
                const int nnz = rows[i].y;
                const int nchunks = (nnz + 31) >> 5;
                for (int c = 0; c < nchunks; ++c, ++t) {
                    const uint32_t chain = chg[og];
                    const uint32_t slot = t % WS_NHI, ls = t % WS_NLO, acc = 2u * og + (chain & 1u);
                    const bool chain_first = (c % WS_CHAIN) == 0;
                    const bool chain_last = (c % WS_CHAIN) == WS_CHAIN - 1 || c == nchunks - 1;
                    const uint32_t k0 = tick();
                    if (chain_first && chain >= 2u) ws_wait(cx, &sh.acc_empty[acc], ((chain >> 1) - 1) & 1u, 5, chain, t, (uint32_t)i);
                    const uint32_t k1 = tick();
                    ws_wait(cx, &sh.full[slot], (t / WS_NHI) & 1u, 6, t, chain, (uint32_t)i);
                    t_empty += k1 - k0;
                    t_full += tick() - k1;
                    if (elect_one()) {                           // (not `lane == 0`: see tc_common.cuh)
                        fence_after_sync();
                        const float *t_hi = hi_s + slot * WS_TILE, *t_lo = lo_s + ls * WS_TILE;
                        const int items = nnz - 32 * c < 32 ? nnz - 32 * c : 32;
                        const int slices = (items + 7) >> 3;
                        const uint32_t d = tmem0 + acc * 128u;
                        for (int ks = 0; ks < slices; ++ks) {
                            const uint64_t dh = ws_desc(t_hi + ks * 1024), dl = ws_desc(t_lo + ks * 1024);
                            mma_tf32(d, dh, dl, idesc, (chain_first && ks == 0) ? 0u : 1u);      // small terms first
                            mma_tf32(d, dl, dh, idesc, 1u);
                            mma_tf32(d, dh, dh, idesc, 1u);
                        }
                        mma_commit(&sh.done_hi[slot]);
                        mma_commit(&sh.done_lo[ls]);
                        if (chain_last) mma_commit(&sh.acc_full[acc]);
                    }
                    __syncwarp();
                    if (chain_last) ++chg[og];
                }
            }
            WS_DBG(20, t_empty); WS_DBG(21, t_full); WS_DBG(22, tick() - m_start); WS_DBG(23, t);
        }
    }
    __syncthreads();
    if (tid == 0 && sh.abort_flag && a.stats) atomicAdd(a.stats + 1, 1ull << 40);
    if (warp == 12) tmem_dealloc(tmem0, 512);
}

template <int LD, bool TMA> static int launch_ws(const WsArgs &a, const CUtensorMap &ymap, int n_ctas, cudaStream_t st) {
    const size_t smem = sizeof(float) * ((size_t)(WS_NHI + WS_NLO) * WS_TILE + WS_NB * WS_NG * 128 + 256 + 16) + 1024;
    auto kern = als_rows_ws_kernel<LD, TMA>;
    CYMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CYMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    kern<<<(unsigned)n_ctas, WS_THREADS, smem, st>>>(a, ymap);
    CYMF_LAUNCHED();
    return 0;
}

}  // namespace tc
}  // namespace cymf

using namespace cymf;

extern "C" int32_t cymf_als_ws_ctas(void) { return (int32_t)sm_count(); }

// Assignment of rows to CTAs, longest rows first, in rounds: the longest unassigned rows go to the CTAs that are still
// below the final mean load, in order of increasing load (the longest row to the least loaded CTA).  Same quality as one-row-at-a-time LPT for these
// length distributions (loads differ by about one row), at a third of its cost (3 ms instead of 12 for 138 k rows: the
// schedule is part of every `fit`).  lengths need not be sorted.
extern "C" int cymf_als_ws_schedule_host(const int64_t *indptr, const int32_t *rows, int32_t n, int32_t n_ctas,
                                         int32_t row_cost, int32_t *cta_ptr, int32_t *rowinfo) {
    CYMF_REQUIRE(indptr && (rows || n == 0) && cta_ptr && (rowinfo || n == 0), "null pointer");
    CYMF_REQUIRE(n >= 0 && n_ctas > 0 && row_cost >= 0, "bad argument");
    auto len = [&](int32_t j) { return indptr[rows[j] + 1] - indptr[rows[j]]; };
    std::vector<int32_t> by_len(n);
    for (int32_t j = 0; j < n; ++j) by_len[j] = j;
    bool sorted = true;
    for (int32_t j = 1; j < n && sorted; ++j) sorted = len(j - 1) >= len(j);
    if (!sorted) std::stable_sort(by_len.begin(), by_len.end(), [&](int32_t x, int32_t y) { return len(x) > len(y); });
    std::vector<int64_t> load(n_ctas, 0);
    std::vector<int32_t> bins(n_ctas), owner(n), count(n_ctas, 0);
    for (int32_t b = 0; b < n_ctas; ++b) bins[b] = b;
    int64_t total = 0;
    for (int32_t j = 0; j < n; ++j) total += len(j) + row_cost;
    const int64_t mean = total / n_ctas + 1;                       // a CTA that has reached the final mean load takes no more rows
    for (int32_t k0 = 0; k0 < n;) {
        if (k0) std::stable_sort(bins.begin(), bins.end(), [&](int32_t x, int32_t y) { return load[x] < load[y]; });
        int32_t k = k0;
        for (int32_t t = 0; t < n_ctas && k < n; ++t) {
            const int32_t b = bins[t];
            if (t > 0 && load[b] >= mean) break;                    // (the least loaded CTA always takes one: progress)
            const int32_t j = by_len[k++];
            owner[j] = b;
            ++count[b];
            load[b] += len(j) + row_cost;
        }
        k0 = k;
    }
    cta_ptr[0] = 0;
    for (int32_t b = 0; b < n_ctas; ++b) cta_ptr[b + 1] = cta_ptr[b] + count[b];
    // Order inside a CTA's list: longest, shortest, second longest, second shortest, ...  The solver groups take the
    // rows alternately, so one group folds the long rows while the other iterates on short ones, and the gather / MMA
    // warps always have a long row to stream while a CG runs: with the list simply sorted, the long rows at its head
    // leave the solvers idle and the short rows at its tail leave the tensor pipe idle (measured: 20 % of the solver
    // time spent waiting for chains while the MMA warp waits for accumulators).
    std::vector<int32_t> fill(n_ctas, 0);
    for (int32_t k = 0; k < n; ++k) {
        const int32_t j = by_len[k], b = owner[j], r = fill[b]++, cnt = count[b];
        // r-th longest of cnt rows -> position 2 r (first half, from the front) or 2 (cnt - 1 - r) + 1 (second half)
        const int32_t half = (cnt + 1) / 2;
        const int32_t pos = cta_ptr[b] + (r < half ? 2 * r : 2 * (cnt - 1 - r) + 1);
        const int64_t lo = indptr[rows[j]];
        rowinfo[4 * pos] = rows[j];
        rowinfo[4 * pos + 1] = (int32_t)len(j);
        rowinfo[4 * pos + 2] = (int32_t)(uint32_t)(lo & 0xffffffffll);
        rowinfo[4 * pos + 3] = (int32_t)(uint32_t)((uint64_t)lo >> 32);
    }
    return 0;
}

typedef CUresult (*WsEncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                    const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static WsEncodeTiledFn ws_encode_tiled() {
    static WsEncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (WsEncodeTiledFn)p;
    }();
    return fn;
}

extern "C" int cymf_als_rows_ws_dev(const int32_t *rowinfo, const int32_t *cta_ptr, int32_t n_ctas, const int32_t *indices,
                                    void *X, const void *Y, int64_t y_rows, int dtype, int32_t K, int32_t ld, double weight,
                                    double cg_tol, int32_t cg_max_iter, unsigned long long *stats, unsigned long long *debug,
                                    void *stream) {
    CYMF_REQUIRE(rowinfo && cta_ptr && indices && X && Y, "null pointer");
    CYMF_REQUIRE(K > 0 && ld >= K && cg_tol > 0 && cg_max_iter > 0 && n_ctas > 0, "bad argument");
    if (!(tc_shape_ok(dtype, ld) && tc_enabled())) {
        set_error("als rows (warp-specialised): needs f32 factors with ld in {32, 64, 96, 128} and tcgen05 enabled");
        return CYMF_EUNSUPPORTED;
    }
    tc::WsArgs a{reinterpret_cast<const int4 *>(rowinfo), cta_ptr, indices, (float *)X, (const float *)Y, cg_max_iter,
                 (float)weight, (float)(cg_tol * cg_tol), stats, debug};
    cudaStream_t st = (cudaStream_t)stream;
    // The gather goes through the TMA unit (tile::gather4: box = 32 columns x 1 row, four rows per copy, written in the
    // operand's SWIZZLE_128B_BASE32B pattern) when the driver can encode the map; cp.async otherwise (CYMF_ALS_WS_TMA=0).
    alignas(64) CUtensorMap ymap;
    memset(&ymap, 0, sizeof(ymap));
    bool tma = false;
    // Measured (B200): both gathers deliver ~10.5 bytes per clock and SM into the swizzled tile; cp.async is 3 % faster
    // while the fixed side lives in L2 (ml-20m shape: 9.65 vs 9.91 ms / epoch), the TMA form 1.3 % faster once it does not
    // (3 M x 300 k x 312 M nnz: 0.1666 vs 0.1688 s / epoch).  Default: TMA for fixed sides beyond 96 MB.
    const char *env = getenv("CYMF_ALS_WS_TMA");
    const bool want = env ? env[0] == '1' : (double)y_rows * ld * 4.0 > 96e6;
    if (want && y_rows > 0 && y_rows < (1ll << 31) && !((uintptr_t)Y & 15u)) {
        if (WsEncodeTiledFn enc = ws_encode_tiled()) {
            const cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)y_rows};
            const cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
            const cuuint32_t box[2] = {32, 1};
            const cuuint32_t estr[2] = {1, 1};
            tma = enc(&ymap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(Y), gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
        }
    }
    switch (ld) {
        case 32: return tma ? tc::launch_ws<32, true>(a, ymap, n_ctas, st) : tc::launch_ws<32, false>(a, ymap, n_ctas, st);
        case 64: return tma ? tc::launch_ws<64, true>(a, ymap, n_ctas, st) : tc::launch_ws<64, false>(a, ymap, n_ctas, st);
        case 96: return tma ? tc::launch_ws<96, true>(a, ymap, n_ctas, st) : tc::launch_ws<96, false>(a, ymap, n_ctas, st);
        case 128: return tma ? tc::launch_ws<128, true>(a, ymap, n_ctas, st) : tc::launch_ws<128, false>(a, ymap, n_ctas, st);
    }
    set_error("als rows (warp-specialised): ld must be 32, 64, 96 or 128");
    return CYMF_EUNSUPPORTED;
}
