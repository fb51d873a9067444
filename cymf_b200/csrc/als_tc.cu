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

template <int LD>
__global__ void __launch_bounds__(ROW_THREADS, 2) als_rows_tc_kernel(const RowSolveArgs a) {
    constexpr int HC = LD / 2;                                   // columns of S held by one thread
    constexpr int STAGE_FLOATS = 2 * TILE_M * CHUNK_K;           // hi tile + lo tile
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *const stages = reinterpret_cast<float *>(smem_raw);
    float *const p_s = stages + ROW_STAGES * STAGE_FLOATS;       // [128]    vector the matvec is applied to
    float *const part = p_s + 128;                               // [2][128] per-column-half partial of S p
    float *const bpart = part + 256;                             // [2][128] per-item-half partial of sum y~
    float *const red_a = bpart + 256;                            // [8]      block reductions (two arrays take turns)
    float *const red_b = red_a + 8;
    __shared__ uint64_t mb_stage[ROW_STAGES];                    // "the MMAs that read this stage have completed"
    __shared__ uint64_t mb_acc[2];                               // "the chain in this accumulator has completed"
    __shared__ uint32_t tmem_slot;
    __shared__ int row_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m = tid & 127;                 // operand row while staging = TMEM lane = row of S
    const int h = tid >> 7;                  // half of the chunk's items while staging; half of S's columns afterwards
    const bool m_on = m < LD;
    const bool own = tid < LD;               // thread k owns element k of x, r, p
    constexpr uint32_t tmem_cols = (2 * LD <= 32) ? 32 : (2 * LD <= 64) ? 64 : (2 * LD <= 128) ? 128 : 256;
    if (warp == 0) tmem_alloc(&tmem_slot, tmem_cols);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < ROW_STAGES; ++s) mbar_init(&mb_stage[s], 1);
        mbar_init(&mb_acc[0], 1);
        mbar_init(&mb_acc[1], 1);
    }
    if (tid < 128) p_s[tid] = 0.f;
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t d_tmem = tmem_slot;
    const uint32_t t_lane = d_tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(h * HC);
    const uint32_t idesc = idesc_tf32(LD);
    const float wm1 = a.weight - 1.f;
    const float *const Ym = a.Y + m;

    uint32_t stage = 0, pend = 0, ph_stage = 0, ph_acc = 0;      // uniform over the CTA
    int acc = 0;                                                  // accumulator of the running chain

    for (;;) {
        if (tid == 0) row_slot = atomicAdd(a.queue, 1);
        __syncthreads();
        const int slot = row_slot;
        __syncthreads();
        if (slot >= a.n_solve) break;
        const int r = a.order[slot];
        const int64_t lo = a.indptr[r];
        const int nnz = (int)(a.indptr[r + 1] - lo);
        float *const xr = a.X + (size_t)r * LD;
        if (nnz == 0) {                                                        // wmf.pyx:154-156
            if (own) xr[tid] = 0.f;
            continue;
        }
        const int32_t *const idx = a.indices + lo;
        const int nchunks = (nnz + CHUNK_K - 1) / CHUNK_K;

        float S[HC];
#pragma unroll
        for (int t = 0; t < HC; ++t) S[t] = 0.f;

        // this thread's 16 items of chunk c: element m of y~_i for i = 32 c + 16 h + j
        float cur[16];
        auto gather16 = [&](int c) {
            const int base = c * CHUNK_K + 16 * h;
            if (!m_on || base >= nnz) {
#pragma unroll
                for (int j = 0; j < 16; ++j) cur[j] = 0.f;
            } else if (base + 16 <= nnz) {
#pragma unroll
                for (int j = 0; j < 16; ++j) cur[j] = __ldg(Ym + (size_t)((uint64_t)(uint32_t)__ldg(idx + base + j) * (uint32_t)LD));
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    cur[j] = base + j < nnz ? __ldg(Ym + (size_t)((uint64_t)(uint32_t)__ldg(idx + base + j) * (uint32_t)LD)) : 0.f;
            }
        };
        // S += the finished chain in accumulator `which`
        auto fold_chain = [&](int which) {
            mbar_wait(&mb_acc[which], (ph_acc >> which) & 1u);
            ph_acc ^= 1u << which;
            fence_after_sync();
#pragma unroll
            for (int c0 = 0; c0 < HC; c0 += 16) {
                float v[16];
                tmem_load16(t_lane + (uint32_t)(which * LD + c0), v);
#pragma unroll
                for (int t = 0; t < 16; ++t) S[c0 + t] += v[t];
            }
            fence_before_sync();
        };

        gather16(0);
        float bs = 0.f;
        int fold = -1;                                            // chain that has ended and is not yet in S
        for (int c = 0; c < nchunks; ++c) {
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
            if (c + 1 < nchunks) gather16(c + 1);                 // in flight underneath the barrier and the MMAs
            fence_async_smem();                                   // generic-proxy writes -> visible to the tensor core
            fence_before_sync();
            __syncthreads();
            const bool chain_first = (c % ROW_CHAIN) == 0;
            const bool chain_last = (c % ROW_CHAIN) == ROW_CHAIN - 1 || c == nchunks - 1;
            if (tid == 0) {
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
            pend |= sbit;
            stage = stage + 1 == ROW_STAGES ? 0 : stage + 1;
            if (fold >= 0) { fold_chain(fold); fold = -1; }       // previous chain, while this one's MMAs run
            if (chain_last) { fold = acc; acc ^= 1; }
        }
        fold_chain(fold);

        // ---- conjugate gradient on (I + (w-1) S) x = b, S in registers --------------------------------------------
        auto matvec = [&]() -> float {                            // this thread's half of row m of S times p_s
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
            for (int t = 0; t < HC; t += 4) {
                const float4 pv = *reinterpret_cast<const float4 *>(p_s + h * HC + t);
                a0 = fmaf(S[t], pv.x, a0); a1 = fmaf(S[t + 1], pv.y, a1);
                a2 = fmaf(S[t + 2], pv.z, a2); a3 = fmaf(S[t + 3], pv.w, a3);
            }
            return (a0 + a1) + (a2 + a3);
        };
        auto block_sum = [&](float v, float *red) -> float {     // one barrier; callers alternate red_a / red_b
            v = warp_sum(v);
            if (lane == 0) red[warp] = v;
            __syncthreads();
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < ROW_THREADS / 32; ++w) s += red[w];
            return s;
        };

        bpart[h * 128 + m] = bs;
        float x = 0.f;
        if (own) { x = xr[tid]; p_s[tid] = x; }                   // warm start
        __syncthreads();
        float b = 0.f;
        if (own) b = a.weight * (bpart[tid] + bpart[128 + tid]);
        part[h * 128 + m] = matvec();
        const float bb = block_sum(b * b, red_a);
        unsigned iters = 0;
        bool stalled = false;
        if (bb > 0.f) {
            float res = 0.f;
            if (own) res = b - (x + wm1 * (part[tid] + part[128 + tid]));       // r0 = b - A x0
            float rs = block_sum(res * res, red_b);
            float p = res;
            while (rs > a.tol2 * bb) {
                if ((int)iters >= a.max_iter) { stalled = true; break; }
                if (own) p_s[tid] = p;
                __syncthreads();
                const float sp = matvec();
                part[h * 128 + m] = sp;
                const float pm = m_on ? p_s[m] : 0.f;
                // p . A p = sum_m p_m (p_m + (w-1) (S p)_m), every thread adds its half row's share
                const float pAp = block_sum(pm * (wm1 * sp) + (h == 0 ? pm * pm : 0.f), red_a);
                if (!(pAp > 0.f)) { stalled = true; break; }
                const float alpha = rs / pAp;
                if (own) {
                    const float Ap = p + wm1 * (part[tid] + part[128 + tid]);
                    x += alpha * p;
                    res -= alpha * Ap;
                }
                const float rs_new = block_sum(res * res, red_b);
                p = res + (rs_new / rs) * p;
                rs = rs_new;
                ++iters;
            }
        } else {
            x = 0.f;                                              // b = 0  =>  x = 0
        }
        if (own) xr[tid] = x;
        if (a.stats && tid == 0) {
            atomicAdd(a.stats, (unsigned long long)iters);
            if (stalled) atomicAdd(a.stats + 1, 1ull);
        }
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(d_tmem, tmem_cols);
}

template <int LD> static int launch_rows(const RowSolveArgs &a, cudaStream_t st) {
    const size_t smem = sizeof(float) * ((size_t)ROW_STAGES * 2 * TILE_M * CHUNK_K + 128 + 256 + 256 + 16);
    auto kern = als_rows_tc_kernel<LD>;
    CYMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CYMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, ROW_THREADS, smem));
    if (per_sm < 1) per_sm = 1;
    constexpr int cols = (2 * LD <= 32) ? 32 : (2 * LD <= 64) ? 64 : (2 * LD <= 128) ? 128 : 256;   // TMEM columns per CTA
    if (per_sm * cols > 512) per_sm = 512 / cols;
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
