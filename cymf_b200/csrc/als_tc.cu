// als_tc.cu -- WMF ALS row solver built around ONE pass over the row (replaces the prange body of WMF._als,
// cymf/wmf.pyx:150-168, for f32 factors with ld in {32, 64, 96, 128}).
//
// In the coordinates y~ = L^-1 y, x~ = L^T x (G = Y^T Y + wd I = L L^T, cymf_chol_transforms_dev) the reference's
// row system  (G + (w-1) sum_{i in row} y_i y_i^T) x = w sum_{i in row} y_i  (wmf.pyx:161-168)  reads
//       (I + (w-1) S) x~ = w sum y~_i ,        S = sum_{i in row} y~_i y~_i^T      (K x K).
// The streaming CG kernel (als.cu) never forms S and pays for it by re-reading the row's item vectors on every CG
// iteration (~7 passes: 2.6x the algorithmic DRAM traffic at C5 scale, issue-bound shuffles on short rows).
// Here S is built ONCE per row on the 5th-generation tensor cores, which is exactly the matrix the reference
// materialises entry by entry (wmf.pyx:161-166):
//   * one persistent 256-thread CTA per row (heaviest-first work queue), two CTAs per SM;
//   * the row's item vectors are gathered 32 at a time (coalesced 128-byte warp reads), split a = hi + lo and staged
//     in shared memory as the K-major UMMA operand tile  T[m][i] = y~_i[m]  (a ring of three 32 KB stages);
//   * one elected thread issues tcgen05.mma kind::tf32  D[128 x ld] += T T^T  as hi*lo + lo*hi + hi*hi (3xTF32)
//     into a TMEM accumulator; completion of a stage is signalled through an mbarrier (tcgen05.commit), so staging
//     of chunk c+1 / c+2 and the gathers of the chunk after run underneath the MMAs of chunk c;
//   * the tensor core truncates when it adds into D, so an accumulator only takes a chain of 16 chunks (512 items,
//     192 MMAs, <= ~6e-6 relative drift); chains alternate between two TMEM accumulators and are summed in registers
//     (round to nearest) while the next chain runs;
//   * every thread then owns half a row of S in registers (thread = TMEM lane m, column half h): conjugate gradient
//     on (I + (w-1) S) x~ = b runs entirely out of registers -- one 64-FMA matvec per thread and three barriers
//     per iteration -- warm-started from the current x~ and stopped at |r| <= tol |b| like the streaming kernel.
// The row is read once: N (K s + 4) bytes per half sweep, the algorithmic figure of SURVEY.md 8(d).
#include <stdlib.h>

#include "tc_common.cuh"

namespace cymf {
namespace tc {

constexpr int ROW_THREADS = 256;
constexpr int ROW_STAGES = 3;     // shared-memory operand stages (hi + lo tiles of 32 items each)
constexpr int ROW_CHAIN = 16;     // 32-item chunks accumulated in TMEM before the chain is folded into registers

struct RowSolveArgs {
    const int64_t *indptr;
    const int32_t *indices;
    const int32_t *order;       // rows to solve, heaviest first
    int32_t n_solve;
    float *X;                   // [rows, ld]  x~ rows, solved in place (warm start = current content)
    const float *Y;             // [n, ld]     y~ rows
    int32_t max_iter;
    float weight, tol2;
    int32_t *queue;             // work-queue head (zeroed before launch)
    unsigned long long *stats;  // [0] CG iterations summed over rows, [1] rows that hit max_iter (may be NULL)
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

struct RowInfo { int r; int nnz; long long lo; };      // r < 0: the queue is exhausted

template <int LD>
__global__ void __launch_bounds__(ROW_THREADS, 2) als_rows_tc_kernel(const RowSolveArgs a) {
    constexpr int HC = LD / 2;                                   // columns of S held by one thread
    constexpr int STAGE_FLOATS = 2 * TILE_M * CHUNK_K;           // hi tile + lo tile
    constexpr int IDX_BLOCK = ROW_THREADS;                       // indices staged per block: one per thread
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *const stages = reinterpret_cast<float *>(smem_raw);
    float *const p_s = stages + ROW_STAGES * STAGE_FLOATS;       // [128]       vector the matvec is applied to
    float *const part = p_s + 128;                               // [2][2][128] per-column-half partials of S p, double buffered
    float *const bpart = part + 512;                             // [2][128]    per-item-half partials of sum y~
    float *const x0_s = bpart + 256;                             // [128]       warm start of the row about to begin
    int32_t *const idx_s = reinterpret_cast<int32_t *>(x0_s + 128);    // [3][256] the row's column indices, block by block
    __shared__ uint64_t mb_full[ROW_STAGES];                     // "all eight warps have written their part of this stage"
    __shared__ uint64_t mb_stage[ROW_STAGES];                    // "the MMAs that read this stage have completed"
    __shared__ uint64_t mb_acc[2];                               // "the chain in this accumulator has completed"
    __shared__ uint32_t tmem_slot;
    __shared__ RowInfo rowq[3];                                  // this CTA's rows i, i+1 (claimed ahead of time)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m = tid & 127;                 // operand row while staging = TMEM lane = row of S
    const int h = tid >> 7;                  // half of the chunk's items while staging; half of S's columns afterwards
    const bool m_on = m < LD;
    const bool l_on = 4 * lane < LD;         // this lane's 4-element slice of x, r, p exists
    constexpr uint32_t tmem_cols = (2 * LD <= 32) ? 32 : (2 * LD <= 64) ? 64 : (2 * LD <= 128) ? 128 : 256;
    if (warp == 0) tmem_alloc(&tmem_slot, tmem_cols);

    // Work queue.  A row is claimed two rows before it is solved and the claim is resolved in steps spread over that
    // time (atomic -> row id -> row extent), each step consuming the previous one's result long after it was issued,
    // so that neither the claim nor the next row's first index block / warm start sits on the critical path.
    // Thread 0 carries the state:  slotA -> rA (row i+2, in progress),  (rB, loB, hiB) = row i+1, published below.
    int slotA = 0, rA = -1, rB = -1;
    long long loB = 0, hiB = 0;
    auto claim_now = [&](int &r, long long &lo, long long &hi) {           // synchronous form (prologue, empty rows)
        const int slot = atomicAdd(a.queue, 1);
        r = -1; lo = 0; hi = 0;
        if (slot < a.n_solve) { r = a.order[slot]; lo = a.indptr[r]; hi = a.indptr[r + 1]; }
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < ROW_STAGES; ++s) { mbar_init(&mb_stage[s], 1); mbar_init(&mb_full[s], ROW_THREADS / 32); }
        mbar_init(&mb_acc[0], 1);
        mbar_init(&mb_acc[1], 1);
        int r0; long long lo0, hi0;
        claim_now(r0, lo0, hi0);
        rowq[0] = RowInfo{r0, (int)(hi0 - lo0), lo0};
        claim_now(rB, loB, hiB);
    }
    if (tid < 128) p_s[tid] = 0.f;
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t idesc = idesc_tf32(LD);
    const float wm1 = a.weight - 1.f;
    const float *const Ym = a.Y + m;

    uint32_t stage = 0, pend = 0, ph_stage = 0, ph_acc = 0, ph_full = 0;      // uniform over the CTA
    int acc = 0;                                                  // accumulator of the running chain
    uint32_t issued = 0;                                          // chunks issued so far: chunk k is issued by warp k mod 8

    // the row about to start: first block of its column indices and its warm start, straight into shared memory
    auto prefetch_row = [&](const RowInfo &ri) {
        if (ri.r >= 0 && ri.nnz > 0) {
            if (tid < ri.nnz) cp_async4(idx_s + tid, a.indices + ri.lo + tid);
            if (warp == 0 && l_on) cp_async16(x0_s + 4 * lane, a.X + (size_t)ri.r * LD + 4 * lane);
        }
    };
    prefetch_row(rowq[0]);

    for (int row_i = 0;; ++row_i) {
        const RowInfo info = rowq[row_i % 3];
        if (info.r < 0) break;
        const int nnz = info.nnz;
        float *const xr = a.X + (size_t)info.r * LD;
        if (nnz == 0) {                                                        // wmf.pyx:154-156
            if (tid < LD) xr[tid] = 0.f;
            if (tid == 0) {
                rowq[(row_i + 1) % 3] = RowInfo{rB, (int)(hiB - loB), loB};
                claim_now(rB, loB, hiB);
            }
            __syncthreads();
            prefetch_row(rowq[(row_i + 1) % 3]);
            continue;
        }
        const int32_t *const idx = a.indices + info.lo;
        const int nchunks = (nnz + CHUNK_K - 1) / CHUNK_K;
        if (tid == 0) slotA = atomicAdd(a.queue, 1);               // row i+2, step 1
        cp_async_wait_all();                                       // this row's first index block and warm start have landed
        __syncthreads();
        const uint32_t d_tmem = tmem_slot;
        const uint32_t t_lane = d_tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(h * HC);

        unsigned long long S2[HC / 2];                             // this thread's half of row m of S, packed pairs

        // this thread's 16 items of chunk c: element m of y~_i for i = 32 c + 16 h + j
        auto gather16 = [&](float (&cur)[16], int c) {
            const int base = c * CHUNK_K + 16 * h;
            if (!m_on || base >= nnz) {
#pragma unroll
                for (int j = 0; j < 16; ++j) cur[j] = 0.f;
                return;
            }
            int32_t it[16];
            const int4 *src = reinterpret_cast<const int4 *>(idx_s + ((base / IDX_BLOCK) % 3) * IDX_BLOCK + (base & (IDX_BLOCK - 1)));
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int4 v = src[q];
                it[4 * q] = v.x; it[4 * q + 1] = v.y; it[4 * q + 2] = v.z; it[4 * q + 3] = v.w;
            }
            if (base + 16 <= nnz) {
#pragma unroll
                for (int j = 0; j < 16; ++j) cur[j] = __ldg(Ym + (size_t)((uint64_t)(uint32_t)it[j] * (uint32_t)LD));
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    cur[j] = base + j < nnz ? __ldg(Ym + (size_t)((uint64_t)(uint32_t)it[j] * (uint32_t)LD)) : 0.f;
            }
        };
        // S (+)= the finished chain in accumulator `which`; the first chain of a row is loaded straight into S
        bool s_empty = true;
        auto fold_chain = [&](int which) {
            mbar_wait(&mb_acc[which], (ph_acc >> which) & 1u);
            ph_acc ^= 1u << which;
            fence_after_sync();
            const uint32_t t0 = t_lane + (uint32_t)(which * LD);
            if (s_empty) {
                uint32_t v[HC / 16][16];
#pragma unroll
                for (int g = 0; g < HC / 16; ++g) tmem_load16_nowait(t0 + 16 * g, v[g]);
#pragma unroll
                for (int g = 0; g < HC / 16; ++g) {
                    tmem_wait_ld16(v[g]);
#pragma unroll
                    for (int t = 0; t < 8; ++t)
                        S2[8 * g + t] = pack2(__uint_as_float(v[g][2 * t]), __uint_as_float(v[g][2 * t + 1]));
                }
                s_empty = false;
            } else {
#pragma unroll
                for (int g = 0; g < HC / 16; ++g) {
                    float v[16];
                    tmem_load16(t0 + 16 * g, v);
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        float lo, hi;
                        unpack2(S2[8 * g + t], lo, hi);
                        S2[8 * g + t] = pack2(lo + v[2 * t], hi + v[2 * t + 1]);
                    }
                }
            }
            fence_before_sync();
        };

        float bs = 0.f;
        int fold = -1;                                            // chain that has ended and is not yet in S
        // one chunk: wait for its stage, split + store this thread's 16 values, hand the stage to the issuer
        auto process = [&](int c, float (&cur)[16], bool refill, bool may_fold) {
            // index blocks of long rows (three buffers): block b+1 is fetched while block b's first chunks are staged and
            // published by a CTA barrier two chunks before its first use (the gather of chunk 8b+8 is issued in
            // iteration 8b+7); warps drift apart by at most ROW_STAGES chunks in between
            if ((c & 7) == 0 && (c / 8 + 1) * IDX_BLOCK < nnz) {
                const int t = (c / 8 + 1) * IDX_BLOCK + tid;
                if (t < nnz) cp_async4(idx_s + ((c / 8 + 1) % 3) * IDX_BLOCK + tid, idx + t);
            }
            if ((c & 7) == 6 && (c / 8 + 1) * IDX_BLOCK < nnz) { cp_async_wait_all(); __syncthreads(); }
            const uint32_t sbit = 1u << stage;
            if (pend & sbit) {                                    // the MMAs that read this stage must have finished
                mbar_wait(&mb_stage[stage], (ph_stage >> stage) & 1u);
                ph_stage ^= sbit;
                pend &= ~sbit;
            }
            float *const t_hi = stages + stage * STAGE_FLOATS, *const t_lo = t_hi + TILE_M * CHUNK_K;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 v = make_float4(cur[4 * q], cur[4 * q + 1], cur[4 * q + 2], cur[4 * q + 3]);
                const float4 hi4 = tf32_hi(v);
                const int o = tile_off(m, 4 * h + q, TILE_M / 8);
                *reinterpret_cast<float4 *>(t_hi + o) = hi4;
                *reinterpret_cast<float4 *>(t_lo + o) = sub4(v, hi4);
                bs += (v.x + v.y) + (v.z + v.w);                  // wmf.pyx:163
            }
            if (refill && c + 1 < nchunks) gather16(cur, c + 1);  // in flight underneath the MMAs
            fence_async_smem();                                   // generic-proxy writes -> visible to the tensor core
            fence_before_sync();
            __syncwarp();
            if (lane == 0) mbar_arrive(&mb_full[stage]);          // this warp's part of the stage is written
            const bool chain_first = (c % ROW_CHAIN) == 0;
            const bool chain_last = (c % ROW_CHAIN) == ROW_CHAIN - 1 || c == nchunks - 1;
            // No CTA barrier per chunk: only the issuing warp waits for the other seven, and the issuer rotates (chunk
            // k: lane 0 of warp k mod 8), so the cost of issuing twelve MMAs is never on every warp's path and warps
            // run up to ROW_STAGES chunks apart.  Issue order = chunk order (the next issuer waits for this one's
            // arrival on the next stage) = execution order on the tensor core.
            if (warp == (int)(issued & 7u)) {
                mbar_wait(&mb_full[stage], (ph_full >> stage) & 1u);
                if (elect_one()) {                                // (not `lane == 0`: see elect_one in tc_common.cuh)
                    fence_after_sync();
                    const int items = nnz - c * CHUNK_K < CHUNK_K ? nnz - c * CHUNK_K : CHUNK_K;
                    const int slices = (items + 7) >> 3;
                    const uint32_t d = d_tmem + (uint32_t)(acc * LD);
                    constexpr uint32_t lbo = (TILE_M / 8) * 128;
                    for (int ks = 0; ks < slices; ++ks) {
                        const uint64_t dh = smem_desc(t_hi + ks * 2 * (lbo >> 2), lbo, 128);
                        const uint64_t dl = smem_desc(t_lo + ks * 2 * (lbo >> 2), lbo, 128);
                        mma_tf32(d, dh, dl, idesc, (chain_first && ks == 0) ? 0u : 1u);      // small terms first
                        mma_tf32(d, dl, dh, idesc, 1u);
                        mma_tf32(d, dh, dh, idesc, 1u);
                    }
                    mma_commit(&mb_stage[stage]);
                    if (chain_last) mma_commit(&mb_acc[acc]);
                }
                __syncwarp();
            }
            ph_full ^= sbit;
            pend |= sbit;
            stage = stage + 1 == ROW_STAGES ? 0 : stage + 1;
            ++issued;
            if (may_fold && fold >= 0) { fold_chain(fold); fold = -1; }       // previous chain, while this one's MMAs run
            if (chain_last) { fold = acc; acc ^= 1; }
        };
        {
            // the first two chunks (all of a row of <= 64 entries): both gathers are issued before the first value is
            // needed, one memory round trip for the two
            float c0[16], c1[16];
            gather16(c0, 0);
            if (nchunks > 1) gather16(c1, 1);
            process(0, c0, false, false);                         // (no chain can end before chunk ROW_CHAIN - 1)
            if (nchunks > 1) {
                if (nchunks > 2) gather16(c0, 2);
                process(1, c1, false, false);
                for (int c = 2; c < nchunks; ++c) process(c, c0, true, true);      // steady state: one chunk ahead
            }
        }
        if (tid == 0) {
            rowq[(row_i + 1) % 3] = RowInfo{rB, (int)(hiB - loB), loB};        // row i+1: visible after the next barrier
            rA = slotA < a.n_solve ? a.order[slotA] : -1;                       // row i+2, step 2
        }
        bpart[h * 128 + m] = bs;
        fold_chain(fold);

        // ---- conjugate gradient on (I + (w-1) S) x = b, S in registers --------------------------------------------
        // Every warp carries the WHOLE of x, r, p (4 elements per lane) and performs the same arithmetic on the same
        // values, so all warps take identical decisions and no vector has to be exchanged except S p itself: one
        // barrier per iteration.
        auto matvec = [&](const float *vec) -> float {           // this thread's half of row m of S times vec
            unsigned long long a0 = 0ull, a1 = 0ull, a2 = 0ull, a3 = 0ull;
            const uint32_t pv = smem_u32(vec + h * HC);
            constexpr int NL = HC / 4;                            // 16-byte pieces of this thread's half of vec
            ulonglong2 u[4];                                      // four loads in flight: one shared-memory latency per four
            auto lds = [&](ulonglong2 &d, int t) {                // volatile: issue order is the program order
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
        auto dot4 = [&](const float4 &u, const float4 &v) -> float {          // over the whole vector, same on all lanes
            return warp_sum(fmaf(u.x, v.x, fmaf(u.y, v.y, fmaf(u.z, v.z, u.w * v.w))));
        };
        auto apply = [&](const float4 &v, const float *pp) -> float4 {        // v + (w-1) (S v) from the two partials
            const float4 s0 = *reinterpret_cast<const float4 *>(pp + 4 * lane);
            const float4 s1 = *reinterpret_cast<const float4 *>(pp + 128 + 4 * lane);
            return make_float4(fmaf(wm1, s0.x + s1.x, v.x), fmaf(wm1, s0.y + s1.y, v.y), fmaf(wm1, s0.z + s1.z, v.z),
                               fmaf(wm1, s0.w + s1.w, v.w));
        };
        part[h * 128 + m] = matvec(x0_s);                         // S x0 for the warm start
        __syncthreads();                                          // bpart, S x0 and rowq[i+1] are visible
        float4 x4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (l_on) x4 = *reinterpret_cast<const float4 *>(x0_s + 4 * lane);
        float4 b4;
        {
            const float4 u = *reinterpret_cast<const float4 *>(bpart + 4 * lane);
            const float4 v = *reinterpret_cast<const float4 *>(bpart + 128 + 4 * lane);
            b4 = make_float4(a.weight * (u.x + v.x), a.weight * (u.y + v.y), a.weight * (u.z + v.z), a.weight * (u.w + v.w));
        }
        const float bb = dot4(b4, b4);
        const float4 ax = apply(x4, part);
        __syncthreads();                                          // x0_s and idx_s may now be refilled for the next row
        prefetch_row(rowq[(row_i + 1) % 3]);                      // lands underneath the CG iterations
        unsigned iters = 0;
        bool stalled = false;
        if (bb > 0.f) {
            float4 r4 = make_float4(b4.x - ax.x, b4.y - ax.y, b4.z - ax.z, b4.w - ax.w);   // r0 = b - A x0
            float rs = dot4(r4, r4);
            float4 p4 = r4;
            const float stop = a.tol2 * bb;
            int buf = 1;
            while (rs > stop) {
                if ((int)iters >= a.max_iter) { stalled = true; break; }
                *reinterpret_cast<float4 *>(p_s + 4 * lane) = p4;             // every warp writes the same values
                __syncwarp();
                float *const pp = part + buf * 256;
                pp[h * 128 + m] = matvec(p_s);
                __syncthreads();
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
            x4 = make_float4(0.f, 0.f, 0.f, 0.f);                 // b = 0  =>  x = 0
        }
        if (warp == 0 && l_on) *reinterpret_cast<float4 *>(xr + 4 * lane) = x4;
        if (tid == 0) {
            rB = rA;                                              // row i+2, step 3: its extent, stored one row later
            loB = hiB = 0;
            if (rA >= 0) { loB = a.indptr[rA]; hiB = a.indptr[rA + 1]; }
            if (a.stats) {
                atomicAdd(a.stats, (unsigned long long)iters);
                if (stalled) atomicAdd(a.stats + 1, 1ull);
            }
        }
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_slot, tmem_cols);
}

template <int LD> static int launch_rows(const RowSolveArgs &a, cudaStream_t st) {
    const size_t smem = sizeof(float) * ((size_t)ROW_STAGES * 2 * TILE_M * CHUNK_K + 128 + 512 + 256 + 128 + 3 * ROW_THREADS);
    auto kern = als_rows_tc_kernel<LD>;
    CYMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CYMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    // two CTAs per SM by construction: 2 x 101 KB of shared memory, 2 x 256 threads x 128 registers, 2 x 256 TMEM columns
    const int per_sm = 2;
    int64_t blocks = (int64_t)sm_count() * per_sm;
    if (blocks > a.n_solve) blocks = a.n_solve;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, ROW_THREADS, smem, st>>>(a);
    CYMF_LAUNCHED();
    return 0;
}

}  // namespace tc

int tc_als_rows(const int64_t *indptr, const int32_t *indices, const int32_t *order, int32_t n_solve, float *X,
                const float *Y, int ld, float weight, float tol2, int32_t max_iter, int32_t *queue,
                unsigned long long *stats, cudaStream_t st) {
    if (const char *v = getenv("CYMF_ALS_TC6"))              // A/B hook: the asynchronous-gather form of the kernel (als_tc6.cu)
        if (v[0] == '1') return tc_als_rows6(indptr, indices, order, n_solve, X, Y, ld, weight, tol2, max_iter, queue, stats, st);
    CYMF_CUDA(cudaMemsetAsync(queue, 0, sizeof(int32_t), st));
    tc::RowSolveArgs a{indptr, indices, order, n_solve, X, Y, max_iter, weight, tol2, queue, stats};
    switch (ld) {
        case 32: return tc::launch_rows<32>(a, st);
        case 64: return tc::launch_rows<64>(a, st);
        case 96: return tc::launch_rows<96>(a, st);
        case 128: return tc::launch_rows<128>(a, st);
    }
    set_error("als rows (tensor cores): ld must be 32, 64, 96 or 128");
    return CYMF_EUNSUPPORTED;
}

}  // namespace cymf

using namespace cymf;

extern "C" int cymf_als_rows_tc_dev(const int64_t *indptr, const int32_t *indices, const int32_t *order, int32_t n_solve,
                                    void *X, const void *Y, int dtype, int32_t K, int32_t ld, double weight,
                                    double cg_tol, int32_t cg_max_iter, int32_t *queue, unsigned long long *stats,
                                    void *stream) {
    CYMF_REQUIRE(indptr && indices && order && X && Y && queue, "null pointer");
    CYMF_REQUIRE(K > 0 && ld >= K && cg_tol > 0 && cg_max_iter > 0, "bad argument");
    if (!(tc_shape_ok(dtype, ld) && tc_enabled())) {
        set_error("als rows (tensor cores): needs f32 factors with ld in {32, 64, 96, 128} and tcgen05 enabled");
        return CYMF_EUNSUPPORTED;
    }
    if (n_solve <= 0) return 0;
    return tc_als_rows(indptr, indices, order, n_solve, (float *)X, (const float *)Y, ld, (float)weight,
                       (float)(cg_tol * cg_tol), cg_max_iter, queue, stats, (cudaStream_t)stream);
}
