// als_dual.cu -- WMF ALS row solver for SHORT rows (n <= 128 entries; AlsSession sends it the rows of <= 64): the prange body of WMF._als
// (cymf/wmf.pyx:150-168) solved in its dual (n x n) form, f32 factors with ld in {32, 64, 96, 128}.
//
// In the coordinates y~ = L^-1 y, x~ = L^T x (G = L L^T, cymf_chol_transforms_dev) the reference's row system
//       (G + (w-1) sum_{i in row} y_i y_i^T) x = w sum_{i in row} y_i                       (wmf.pyx:161-168)
// reads (I_K + c Y~^T Y~) x~ = w Y~^T 1 with c = w - 1 and Y~ the n x K matrix of the row's item vectors.  By the
// push-through identity (I_K + c Y~^T Y~)^-1 Y~^T = Y~^T (I_n + c Y~ Y~^T)^-1 the same x~ is
//       x~ = w Y~^T z ,      (I_n + c Y~ Y~^T) z = 1 ,
// an n x n system.  Most users have fewer entries than factors (ml-20m shape: 65 % of the rows have n <= 128 = K), and
// for those the dual is smaller in every respect: the Gram Y~ Y~^T is n^2 K instead of n K^2 products, a CG iteration
// n^2 instead of K^2, and -- since 0 <= Y~ Y~^T <= I -- the condition number is at most w whatever the row.
//
// Layout of the work: a TILE is 128 gathered item vectors = 128 rows of a K-major tcgen05 operand (a gathered vector
// IS a K-major operand row, no transposition); rows are packed by length class, BLK = 128 / 64 / 32 slots per row, so
// a tile carries 1 / 2 / 4 rows of X.  Per tile, one 128-thread CTA (three per SM, persistent, work queue):
//   1. gather: cp.async (LDGSTS.128) drops the vectors straight into the operand tile -- no register staging, all
//      of a tile's loads in flight at once;
//   2. D = T T^T on the tensor cores as hi*lo^T + lo*hi^T + hi*hi^T (3xTF32).  The hi operand is the gathered data
//      itself (the tensor core ignores the 13 low mantissa bits of a tf32 container); only lo = a - hi is produced,
//      8 reduction elements at a time into a two-stage ring, so splitting overlaps the MMAs;
//   3. thread i reads row i of its row's diagonal block of D from TMEM (the off-diagonal blocks of a packed tile are
//      never read) and keeps c D_ij in registers; conjugate gradient on (I + c D) z = 1 from z = 0 runs out of
//      registers inside a warp (BLK = 32), a warp pair (64) or the CTA (128);
//   4. x~ = w sum_j z_j y~_j from the operand tile still in shared memory (lane = 16-byte piece of the vector: the
//      tile's K-chunk stride is padded by 16 bytes so that this read is bank-conflict free too) -> 512-byte rows of X.
// Every item vector is read once: n (K s + 4) + K s + 8 bytes per row, the algorithmic figure of SURVEY.md 8(d).
#include <stdlib.h>

#include "tc_common.cuh"

namespace cymf {
namespace tc {

constexpr int DUAL_THREADS = 128;
constexpr int QSTRIDE = 2048 + 16;     // bytes between consecutive 16-byte K chunks of the 128-row operand tile (padded)
constexpr int DUAL_KS = 8;             // reduction elements per lo stage (one MMA k-slice)
constexpr int DUAL_NLO = 2;            // lo stages

struct DualArgs {
    const int64_t *indptr;
    const int32_t *indices;
    const int32_t *order;       // rows of this class, any order (longest first keeps the tiles balanced)
    int32_t n_rows;
    float *X;                   // [rows, ld]  x~ rows (overwritten; no warm start in the dual form)
    const float *Y;             // [n, ld]     y~ rows
    int32_t max_iter;
    float weight, tol2;
    int32_t *queue;             // work-queue head (zeroed before launch)
    unsigned long long *stats;  // [0] CG iterations summed over rows, [1] rows that hit max_iter (may be NULL)
};

__device__ __forceinline__ uint32_t gran_off(int m, int q) {          // byte offset of (operand row m, 16-byte K chunk q)
    return (uint32_t)(q * QSTRIDE + (m >> 3) * 128 + (m & 7) * 16);
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cp_wait_dyn(int n) {                   // n is a compile-time constant after unrolling
    switch (n) {
        case 0: cp_wait<0>(); break;
        case 1: cp_wait<1>(); break;
        case 2: cp_wait<2>(); break;
        case 3: cp_wait<3>(); break;
        case 4: cp_wait<4>(); break;
        case 5: cp_wait<5>(); break;
        case 6: cp_wait<6>(); break;
        default: cp_wait<7>(); break;
    }
}
__device__ __forceinline__ float warp_sum_d(float v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
// barrier over the BLK threads that share a row of X
template <int BLK> __device__ __forceinline__ void group_sync(int group) {
    if constexpr (BLK == 32) __syncwarp();
    else if constexpr (BLK == 128) __syncthreads();
    else asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(BLK) : "memory");
}
// sum over the row's BLK threads, identical on all of them (fixed order)
template <int BLK> __device__ __forceinline__ float group_sum(float v, float *red, int warp, int lane, int group, int which) {
    v = warp_sum_d(v);
    if constexpr (BLK == 32) {
        return v;
    } else {
        constexpr int WPR = BLK / 32;
        if (lane == 0) red[which * 4 + warp] = v;
        group_sync<BLK>(group);
        float s = red[which * 4 + group * WPR];
#pragma unroll
        for (int k = 1; k < WPR; ++k) s += red[which * 4 + group * WPR + k];
        return s;
    }
}

template <int LD, int BLK>
__global__ void __launch_bounds__(DUAL_THREADS, 3) als_rows_dual_kernel(const DualArgs a) {
    constexpr int RPT = 128 / BLK;                       // rows of X per tile
    constexpr int NQ = LD / 4;                           // 16-byte K chunks per vector
    constexpr int NI = LD / 16;                          // gather instructions per 8-row group (4 chunks each)
    constexpr int NS = LD / DUAL_KS;                     // lo stages per tile
    constexpr int QPS = DUAL_KS / 4;                     // K chunks per stage
    constexpr int LO_BYTES = QPS * QSTRIDE;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char *const raw = smem;                                     // operand tile: the gathered vectors
    unsigned char *const lo = raw + NQ * QSTRIDE;                        // [DUAL_NLO] lo stages
    float *const p_s = reinterpret_cast<float *>(lo + DUAL_NLO * LO_BYTES);      // [128]
    float *const z_s = p_s + 128;                                        // [128]
    float *const red = z_s + 128;                                        // [2][4]
    int32_t *const idx_s = reinterpret_cast<int32_t *>(red + 8);         // [2][128]
    float *const part_s = reinterpret_cast<float *>(lo);                 // [4][LD], after the MMAs (lo stages are free)
    __shared__ uint64_t mb_lo[DUAL_NLO];
    __shared__ uint64_t mb_acc;
    __shared__ uint32_t tmem_slot;
    __shared__ int tile_s[2];
    __shared__ int used_s[2];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int slot = tid / BLK, e = tid % BLK, slot_base = slot * BLK;
    const int phase = lane & 7, qg = lane >> 3;
    const int n_tiles = (a.n_rows + RPT - 1) / RPT;
    if (warp == 0) tmem_alloc(&tmem_slot, 128);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < DUAL_NLO; ++s) mbar_init(&mb_lo[s], 1);
        mbar_init(&mb_acc, 1);
        tile_s[0] = atomicAdd(a.queue, 1);
        used_s[0] = used_s[1] = 0;
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t d_tmem = tmem_slot;
    const uint32_t t_lane = d_tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)slot_base;
    const float wm1 = a.weight - 1.f;

    // this thread's slot of a tile: the row of X it belongs to and its extent
    auto slot_row = [&](int tile) -> int {
        const long long ri = (long long)tile * RPT + slot;
        return (tile < n_tiles && ri < a.n_rows) ? a.order[ri] : -1;
    };
    int tile = tile_s[0];
    int cur_r = slot_row(tile), cur_n = 0;
    {
        long long l0 = 0;
        if (cur_r >= 0) { l0 = a.indptr[cur_r]; cur_n = (int)(a.indptr[cur_r + 1] - l0); }
        idx_s[tid] = e < cur_n ? a.indices[l0 + e] : -1;
        if (e == 0 && cur_n > 0) atomicMax(&used_s[0], slot_base + cur_n);
    }
    uint32_t ph_lo = 0, lo_busy = 0, ph_acc = 0;

    for (int t_i = 0; tile < n_tiles; ++t_i) {
        const int buf = t_i & 1;
        __syncthreads();                                   // idx_s[buf], used_s[buf] visible; the previous tile is done with `raw`
        if (tid == 0) { tile_s[buf ^ 1] = atomicAdd(a.queue, 1); used_s[buf ^ 1] = 0; }
        const int n_used = used_s[buf];
        // ---- 1. gather: this warp's 32 operand rows, 8 rows x 4 K chunks per instruction --------------------------
        {
            int32_t rr[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) rr[g] = idx_s[buf * 128 + 32 * warp + 8 * g + phase];
#pragma unroll
            for (int i = 0; i < NI; ++i) {
#pragma unroll
                for (int g = 0; g < 4; ++g)
                    if (rr[g] >= 0)
                        cp_async16(raw + gran_off(32 * warp + 8 * g + phase, 4 * i + qg),
                                   a.Y + (size_t)((uint64_t)(uint32_t)rr[g] * (uint32_t)LD) + 4 * (4 * i + qg));
                cp_commit();                               // group i = K chunks 4i .. 4i+3 = lo stages 2i, 2i+1
            }
        }
        cp_wait_dyn(NI - 1);
        __syncthreads();                                   // K chunks 0..3 of every row have landed; tile_s[buf^1] visible
        // the next tile's rows, resolved step by step underneath this tile's work
        const int ntile = tile_s[buf ^ 1];
        const int nr = slot_row(ntile);
        long long nlo = 0;
        int nn = 0, nidx = -1;

        // ---- 2. D = T T^T, 3xTF32 ------------------------------------------------------------------------------
        const int n16 = n_used > 16 ? (n_used + 15) & ~15 : 16;
        const uint32_t idesc = idesc_tf32(n16);
#pragma unroll
        for (int s = 0; s < NS; ++s) {
            const int ls = s % DUAL_NLO;
            if (lo_busy & (1u << ls)) {                    // the MMAs that read this lo stage must have completed
                mbar_wait(&mb_lo[ls], (ph_lo >> ls) & 1u);
                ph_lo ^= 1u << ls;
            }
            unsigned char *const lo_st = lo + ls * LO_BYTES;
#pragma unroll
            for (int ql = 0; ql < QPS; ++ql) {
                const float4 v = *reinterpret_cast<const float4 *>(raw + gran_off(tid, s * QPS + ql));
                *reinterpret_cast<float4 *>(lo_st + gran_off(tid, ql)) = sub4(v, tf32_hi(v));
            }
            // the chunk group the NEXT stage reads must have landed before the barrier below releases it
            if (s + 1 < NS && ((s + 1) * DUAL_KS) % 16 == 0) cp_wait_dyn(NI - 1 - (s + 1) * DUAL_KS / 16);
            fence_async_smem();                            // generic-proxy writes (cp.async, st.shared) -> tensor core
            fence_before_sync();
            __syncthreads();
            if (warp == 0 && elect_one()) {                // (not `tid == 0`: see elect_one in tc_common.cuh)
                fence_after_sync();
#pragma unroll
                for (int ks = 0; ks < DUAL_KS / 8; ++ks) {
                    const uint64_t dh = smem_desc(raw + (s * QPS + 2 * ks) * QSTRIDE, QSTRIDE, 128);
                    const uint64_t dl = smem_desc(lo_st + (2 * ks) * QSTRIDE, QSTRIDE, 128);
                    mma_tf32(d_tmem, dh, dl, idesc, (s == 0 && ks == 0) ? 0u : 1u);      // small terms first
                    mma_tf32(d_tmem, dl, dh, idesc, 1u);
                    mma_tf32(d_tmem, dh, dh, idesc, 1u);
                }
                mma_commit(&mb_lo[ls]);
                if (s == NS - 1) mma_commit(&mb_acc);
            }
            lo_busy |= 1u << ls;
        }
        if (nr >= 0) { nlo = a.indptr[nr]; nn = (int)(a.indptr[nr + 1] - nlo); }      // next tile, step 2
        mbar_wait(&mb_acc, ph_acc);
        ph_acc ^= 1u;
        lo_busy = 0;                                       // everything issued so far has completed ...
#pragma unroll
        for (int s = 0; s < DUAL_NLO; ++s)                 // ... including each stage's last commit: consume its phase
            if (NS > s) { mbar_wait(&mb_lo[s], (ph_lo >> s) & 1u); ph_lo ^= 1u << s; }
        fence_after_sync();

        // ---- 3. row i of the diagonal block, CG on (I + c D) z = 1 -----------------------------------------------
        const bool valid = e < cur_n;
        unsigned long long M2[BLK / 2];
#pragma unroll
        for (int c = 0; c < BLK / 32; ++c) {
            float v[32];
            tmem_load32(t_lane + 32 * c, v);
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const float m0 = (valid && 32 * c + 2 * j < cur_n) ? wm1 * v[2 * j] : 0.f;
                const float m1 = (valid && 32 * c + 2 * j + 1 < cur_n) ? wm1 * v[2 * j + 1] : 0.f;
                M2[16 * c + j] = pack2(m0, m1);
            }
        }
        fence_before_sync();
        if (e < nn) nidx = a.indices[nlo + e];              // next tile, step 3

        float z = 0.f, r = valid ? 1.f : 0.f, p = r;
        float rs = (float)cur_n;
        const float stop = a.tol2 * rs;
        unsigned iters = 0;
        bool stalled = false;
        if (cur_n > 0) {
            while (rs > stop) {
                if ((int)iters >= a.max_iter) { stalled = true; break; }
                p_s[tid] = p;
                group_sync<BLK>(slot);
                unsigned long long a0 = 0ull, a1 = 0ull, a2 = 0ull, a3 = 0ull;
                const ulonglong2 *pp = reinterpret_cast<const ulonglong2 *>(p_s + slot_base);
#pragma unroll
                for (int t = 0; t < BLK / 4; ++t) {
                    const ulonglong2 w = pp[t];
                    if (t & 1) { fma2(a2, M2[2 * t], w.x); fma2(a3, M2[2 * t + 1], w.y); }
                    else { fma2(a0, M2[2 * t], w.x); fma2(a1, M2[2 * t + 1], w.y); }
                }
                float s0, s1, s2, s3, s4, s5, s6, s7;
                unpack2(a0, s0, s1); unpack2(a1, s2, s3); unpack2(a2, s4, s5); unpack2(a3, s6, s7);
                const float y = p + (((s0 + s1) + (s2 + s3)) + ((s4 + s5) + (s6 + s7)));
                const float pAp = group_sum<BLK>(p * y, red, warp, lane, slot, 0);
                if (!(pAp > 0.f)) { stalled = true; break; }
                const float alpha = rs / pAp;
                z = fmaf(alpha, p, z);
                r = fmaf(-alpha, y, r);
                const float rs_new = group_sum<BLK>(r * r, red, warp, lane, slot, 1);
                const float beta = rs_new / rs;
                p = fmaf(beta, p, r);
                rs = rs_new;
                ++iters;
            }
        }
        // ---- 4. x~ = w sum_j z_j y~_j from the operand tile ------------------------------------------------------
        z_s[tid] = valid ? a.weight * z : 0.f;
        group_sync<BLK>(slot);
        {
            int nv = cur_n - (32 * warp - slot_base);      // this warp's share of the row's entries
            nv = nv < 0 ? 0 : (nv > 32 ? 32 : nv);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            if (lane < NQ) {
                const unsigned char *const col = raw + lane * QSTRIDE;
                const float *const zz = z_s + 32 * warp;
#pragma unroll 4
                for (int j = 0; j < nv; ++j) {
                    const float4 v = *reinterpret_cast<const float4 *>(col + ((j >> 3) * 128 + (j & 7) * 16) + warp * 512);
                    const float zj = zz[j];
                    acc.x = fmaf(zj, v.x, acc.x); acc.y = fmaf(zj, v.y, acc.y);
                    acc.z = fmaf(zj, v.z, acc.z); acc.w = fmaf(zj, v.w, acc.w);
                }
            }
            if constexpr (BLK == 32) {
                if (cur_r >= 0 && lane < NQ) *reinterpret_cast<float4 *>(a.X + (size_t)cur_r * LD + 4 * lane) = acc;
            } else {
                constexpr int WPR = BLK / 32;
                if (lane < NQ) *reinterpret_cast<float4 *>(part_s + warp * LD + 4 * lane) = acc;
                group_sync<BLK>(slot);
                if (warp == slot * WPR && cur_r >= 0 && lane < NQ) {
                    float4 s = acc;
#pragma unroll
                    for (int k = 1; k < WPR; ++k) {
                        const float4 u = *reinterpret_cast<const float4 *>(part_s + (warp + k) * LD + 4 * lane);
                        s.x += u.x; s.y += u.y; s.z += u.z; s.w += u.w;
                    }
                    *reinterpret_cast<float4 *>(a.X + (size_t)cur_r * LD + 4 * lane) = s;
                }
            }
        }
        if (e == 0 && cur_r >= 0 && a.stats) {
            atomicAdd(a.stats, (unsigned long long)iters);
            if (stalled) atomicAdd(a.stats + 1, 1ull);
        }
        // ---- hand over to the next tile ---------------------------------------------------------------------------
        idx_s[(buf ^ 1) * 128 + tid] = nidx;
        if (e == 0 && nn > 0) atomicMax(&used_s[buf ^ 1], slot_base + nn);
        cur_r = nr;
        cur_n = nn;
        tile = ntile;
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_slot, 128);
}

template <int LD, int BLK> static int launch_dual(const DualArgs &a, cudaStream_t st) {
    constexpr int RPT = 128 / BLK;
    size_t smem = (size_t)(LD / 4) * QSTRIDE + (size_t)DUAL_NLO * (DUAL_KS / 4) * QSTRIDE + sizeof(float) * (128 + 128 + 8) +
                  sizeof(int32_t) * 2 * 128;
    // TMEM holds four 128-column accumulators per SM: never let more than three CTAs share one (58 KB each at least)
    if (smem < 58 * 1024) smem = 58 * 1024;
    auto kern = als_rows_dual_kernel<LD, BLK>;
    CYMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CYMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    const int64_t tiles = ((int64_t)a.n_rows + RPT - 1) / RPT;
    int64_t blocks = (int64_t)sm_count() * 3;
    if (blocks > tiles) blocks = tiles;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, DUAL_THREADS, smem, st>>>(a);
    CYMF_LAUNCHED();
    return 0;
}

template <int LD> static int launch_dual_classes(const int64_t *indptr, const int32_t *indices, const int32_t *order,
                                                 const int32_t counts[3], float *X, const float *Y, float weight, float tol2,
                                                 int32_t max_iter, int32_t *queue, unsigned long long *stats, cudaStream_t st) {
    int32_t first = 0;
    for (int c = 0; c < 3; ++c) {
        if (counts[c] > 0) {
            DualArgs a{indptr, indices, order + first, counts[c], X, Y, max_iter, weight, tol2, queue + c, stats};
            int rc = 0;
            if (c == 0) {
                if constexpr (LD >= 128) rc = launch_dual<LD, 128>(a, st);
                else { set_error("als rows (dual): rows of more than 64 entries need ld = 128"); return CYMF_EINVAL; }
            } else if (c == 1) {
                rc = launch_dual<LD, 64>(a, st);
            } else {
                rc = launch_dual<LD, 32>(a, st);
            }
            if (rc) return rc;
        }
        first += counts[c];
    }
    return 0;
}

}  // namespace tc
}  // namespace cymf

using namespace cymf;

extern "C" int cymf_als_rows_dual_dev(const int64_t *indptr, const int32_t *indices, const int32_t *order, int32_t n128,
                                      int32_t n64, int32_t n32, void *X, const void *Y, int dtype, int32_t K, int32_t ld,
                                      double weight, double cg_tol, int32_t cg_max_iter, int32_t *queue,
                                      unsigned long long *stats, void *stream) {
    CYMF_REQUIRE(indptr && indices && order && X && Y && queue, "null pointer");
    CYMF_REQUIRE(K > 0 && ld >= K && cg_tol > 0 && cg_max_iter > 0 && n128 >= 0 && n64 >= 0 && n32 >= 0, "bad argument");
    if (!(tc_shape_ok(dtype, ld) && tc_enabled())) {
        set_error("als rows (dual): needs f32 factors with ld in {32, 64, 96, 128} and tcgen05 enabled");
        return CYMF_EUNSUPPORTED;
    }
    if (n128 + n64 + n32 == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    CYMF_CUDA(cudaMemsetAsync(queue, 0, 3 * sizeof(int32_t), st));
    const int32_t counts[3] = {n128, n64, n32};
    const float tol2 = (float)(cg_tol * cg_tol);
    switch (ld) {
        case 32: return tc::launch_dual_classes<32>(indptr, indices, order, counts, (float *)X, (const float *)Y, (float)weight, tol2, cg_max_iter, queue, stats, st);
        case 64: return tc::launch_dual_classes<64>(indptr, indices, order, counts, (float *)X, (const float *)Y, (float)weight, tol2, cg_max_iter, queue, stats, st);
        case 96: return tc::launch_dual_classes<96>(indptr, indices, order, counts, (float *)X, (const float *)Y, (float)weight, tol2, cg_max_iter, queue, stats, st);
        case 128: return tc::launch_dual_classes<128>(indptr, indices, order, counts, (float *)X, (const float *)Y, (float)weight, tol2, cg_max_iter, queue, stats, st);
    }
    set_error("als rows (dual): ld must be 32, 64, 96 or 128");
    return CYMF_EUNSUPPORTED;
}
