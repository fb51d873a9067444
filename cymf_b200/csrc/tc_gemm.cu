// tc_gemm.cu -- the GEMM-shaped pieces of the WMF half sweep on the 5th-generation tensor cores (tcgen05 + TMEM).
//
//   tc_rows_times_matrix_kernel : out[r, :] = in[r, :] * B  for [rows, ld] f32 matrices, B [ld, ld]  (the three changes of
//       variables of the transformed ALS solver: Y~ = Y L^-T, X~ = X L, X = X~ L^-1).  One CTA per 128-row tile,
//       accumulator D[128 x ld] in tensor memory, operands staged in shared memory in the canonical K-major
//       no-swizzle UMMA layout.  The epilogue reads D back with tcgen05.ld and stores every row to ALL destinations
//       it is given -- with the peers' replicas of the factor matrix as destinations this one kernel is the GEMM and
//       the all-gather over NVLink peer memory.
//   tc_gram_partial_kernel      : per-slab partial of G = Y^T Y (cymf/wmf.pyx:142) with the same machinery; slabs are
//       summed in f64 in a fixed order by gram_finish_kernel (als.cu).  Its GATHER form builds, slab by slab, the
//       per-row matrix sum_{c in row} y_c y_c^T (wmf.pyx:161-166) of the few very long rows of a half sweep, which
//       are then solved directly (als_heavy_solve_kernel) instead of being streamed ~7 times by one CTA.
//
// Precision: kind::tf32 keeps 10 mantissa bits per operand, far too few for the 1e-4 parity bar, so every operand
// is split as a = hi + lo (hi = a with the low 13 mantissa bits cleared, lo = a - hi, both exactly representable) and
// three MMAs are issued per k-slice: hi*hi + hi*lo + lo*hi ("3xTF32"); the dropped lo*lo term is ~2^-22 relative.
// Accumulation is f32 in TMEM.
#include <stdlib.h>

#include "tc_common.cuh"

namespace cymf {
namespace tc {

constexpr int MAX_DESTS = 8;
struct MultiOutF { float *p[MAX_DESTS]; int n; };

__global__ void __launch_bounds__(128) tc_rows_times_matrix_kernel(const float *in /* may alias a destination */, const MultiOutF outs,
                                                                   const float *__restrict__ B, int64_t rows, int ld,
                                                                   uint32_t tmem_cols) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *a_hi = reinterpret_cast<float *>(smem_raw);            // [TILE_M x 32]
    float *a_lo = a_hi + TILE_M * CHUNK_K;
    float *b_hi = a_lo + TILE_M * CHUNK_K;                        // [ld x 32]
    float *b_lo = b_hi + ld * CHUNK_K;
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&tmem_slot, tmem_cols);
    if (tid == 0) mbar_init(&mbar, 1);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t d_tmem = tmem_slot;
    const uint32_t idesc = idesc_tf32(ld);
    const int a_groups = TILE_M / 8, b_groups = ld / 8;
    uint32_t phase = 0;
    for (int64_t tile = blockIdx.x; tile * TILE_M < rows; tile += gridDim.x) {
        const int64_t row0 = tile * TILE_M;
        for (int kc = 0; kc < ld / CHUNK_K; ++kc) {
            {   // A chunk: thread m stages row m, columns [32 kc, 32 kc + 32)
                const bool ok = row0 + tid < rows;
                const float4 *src = reinterpret_cast<const float4 *>(in + (size_t)(row0 + tid) * ld + kc * CHUNK_K);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 v = ok ? src[q] : make_float4(0.f, 0.f, 0.f, 0.f);   // coherent load: `in` may be `out`
                    const float4 h = tf32_hi(v);
                    const int o = tile_off(tid, q, a_groups);
                    *reinterpret_cast<float4 *>(a_hi + o) = h;
                    *reinterpret_cast<float4 *>(a_lo + o) = sub4(v, h);
                }
            }
            if (tid < ld) {   // B chunk: operand element (n, k) = B[32 kc + k][n]; thread n
                const float *src = B + (size_t)kc * CHUNK_K * ld + tid;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 v = make_float4(__ldg(src + (4 * q) * ld), __ldg(src + (4 * q + 1) * ld),
                                                 __ldg(src + (4 * q + 2) * ld), __ldg(src + (4 * q + 3) * ld));
                    const float4 h = tf32_hi(v);
                    const int o = tile_off(tid, q, b_groups);
                    *reinterpret_cast<float4 *>(b_hi + o) = h;
                    *reinterpret_cast<float4 *>(b_lo + o) = sub4(v, h);
                }
            }
            fence_async_smem();                      // generic-proxy writes -> visible to the tensor core (async proxy)
            __syncthreads();
            if (tid < 32 && elect_one()) {               // one lane of warp 0 (not `tid == 0`: see elect_one)
                fence_after_sync();
                // every 32-element reduction chunk gets its own accumulator (TMEM columns [kc ld, kc ld + ld)):
                // the tensor core truncates when it adds into D, so short chains (12 MMAs) keep the f32 result
                // within ~6e-7 of exact; the chunk sums are added in registers in the epilogue
                issue_chunk(d_tmem + (uint32_t)(kc * ld), a_hi, a_lo, b_hi, b_lo, a_groups, b_groups, idesc, true);
                mma_commit(&mbar);
            }
            mbar_wait(&mbar, phase);                 // MMAs of this chunk are done: operands may be overwritten
            phase ^= 1u;
        }
        fence_after_sync();
        const int64_t row = row0 + warp * 32 + lane;
        for (int c0 = 0; c0 < ld; c0 += 32) {        // epilogue: TMEM -> registers -> every destination
            float v[32];
            tmem_load32(d_tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
            for (int kc = 1; kc < ld / CHUNK_K; ++kc) {
                float u[32];
                tmem_load32(d_tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(kc * ld + c0), u);
#pragma unroll
                for (int t = 0; t < 32; ++t) v[t] += u[t];
            }
            if (row < rows)
                for (int d = 0; d < outs.n; ++d) {
                    float4 *o = reinterpret_cast<float4 *>(outs.p[d] + (size_t)row * ld + c0);
#pragma unroll
                    for (int q = 0; q < 8; ++q) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                }
        }
        fence_before_sync();
        __syncthreads();                             // D has been read: the next tile may overwrite it
        fence_after_sync();
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(d_tmem, tmem_cols);
}

// partial[slab] = Y[slab rows]^T Y[slab rows]  (dense [K, K] doubles, like gram_partial_kernel).
// Each 32-row chunk is multiplied on the tensor cores into a fresh TMEM accumulator (12 MMAs), read back and added
// into an f64 copy of the slab's result in shared memory, so the truncating f32 accumulation never runs long chains.
// GATHER = false: slab b = rows [512 b, 512 b + 512) of Y.
// GATHER = true : slab s belongs to heavy row h (first_slab[h] <= s < first_slab[h + 1]) of a CSR; its operand rows are
//                 Y[indices[lo + t]] for the row's entries t in [512 (s - first_slab[h]), ...): the partial of
//                 sum_{c in row} y_c y_c^T that the reference accumulates entry by entry (cymf/wmf.pyx:161-166);
//                 bsum[s][m] = sum of the slab's y_c[m] (wmf.pyx:163).
struct GatherPlan {
    const int64_t *indptr;
    const int32_t *indices;
    const int32_t *order;          // heavy row h = CSR row order[h]
    const int32_t *first_slab;     // [n_heavy + 1]
    int32_t n_heavy;
    double *bsum;                  // [n_slabs][ld]
};

template <bool GATHER>
__global__ void __launch_bounds__(128) tc_gram_partial_kernel(const float *__restrict__ Y, int64_t n, int64_t slab_rows,
                                                              int K, int ld, double *__restrict__ partial,
                                                              uint32_t tmem_cols, const GatherPlan plan) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *t_hi = reinterpret_cast<float *>(smem_raw);            // [TILE_M x 32]: operand rows = columns of Y (zero past ld)
    float *t_lo = t_hi + TILE_M * CHUNK_K;
    double *acc = reinterpret_cast<double *>(t_lo + TILE_M * CHUNK_K);   // [ld][TILE_M]: acc[c][m] = G[m][c] so far
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&tmem_slot, tmem_cols);
    if (tid == 0) mbar_init(&mbar, 1);
    for (int c = 0; c < ld; ++c) acc[c * TILE_M + tid] = 0.0;
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t d_tmem = tmem_slot;
    const uint32_t idesc = idesc_tf32(ld);
    const int groups = TILE_M / 8;
    int64_t r0, r1;
    const int32_t *gidx = nullptr;
    if (GATHER) {
        int lo = 0, hi = plan.n_heavy;                            // last h with first_slab[h] <= blockIdx.x
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (plan.first_slab[mid] <= (int)blockIdx.x) lo = mid; else hi = mid; }
        const int32_t row = plan.order[lo];
        const int64_t beg = plan.indptr[row], end = plan.indptr[row + 1];
        r0 = beg + (int64_t)((int)blockIdx.x - plan.first_slab[lo]) * TC_GRAM_SLAB;
        r1 = r0 + TC_GRAM_SLAB < end ? r0 + TC_GRAM_SLAB : end;
        gidx = plan.indices;
    } else {
        r0 = (int64_t)blockIdx.x * slab_rows;                    // a multiple of 32 rows per CTA
        r1 = r0 + slab_rows < n ? r0 + slab_rows : n;
    }
    double colsum = 0.0;
    uint32_t phase = 0;
    for (int64_t base = r0; base < r1; base += CHUNK_K) {
        // operand element (m, k) = Y[row(base + k)][m]; thread m (column of Y), zero rows for m >= ld or past the slab
        const float *src = Y + (size_t)base * ld + tid;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            float e[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const bool ok = tid < ld && base + 4 * q + t < r1;
                if (GATHER) e[t] = ok ? __ldg(Y + (size_t)__ldg(gidx + base + 4 * q + t) * ld + tid) : 0.f;
                else e[t] = ok ? __ldg(src + (size_t)(4 * q + t) * ld) : 0.f;
            }
            if (GATHER) colsum += ((double)e[0] + (double)e[1]) + ((double)e[2] + (double)e[3]);
            const float4 v = make_float4(e[0], e[1], e[2], e[3]);
            const float4 h = tf32_hi(v);
            const int o = tile_off(tid, q, groups);
            *reinterpret_cast<float4 *>(t_hi + o) = h;
            *reinterpret_cast<float4 *>(t_lo + o) = sub4(v, h);
        }
        fence_async_smem();
        __syncthreads();
        if (tid < 32 && elect_one()) {                   // one lane of warp 0 (not `tid == 0`: see elect_one)
            fence_after_sync();
            // the B operand is the same tile restricted to its first ld rows (same base address, same strides)
            issue_chunk(d_tmem, t_hi, t_lo, t_hi, t_lo, groups, groups, idesc, true);
            mma_commit(&mbar);
        }
        mbar_wait(&mbar, phase);
        phase ^= 1u;
        fence_after_sync();
        for (int c0 = 0; c0 < ld; c0 += 32) {                    // D row m = column m of Y -> this thread's f64 column
            float v[32];
            tmem_load32(d_tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
            for (int t = 0; t < 32; ++t) acc[(c0 + t) * TILE_M + tid] += (double)v[t];
        }
        fence_before_sync();
        __syncthreads();                                          // D was read: the next chunk may overwrite it
        fence_after_sync();
    }
    double *out = partial + (size_t)blockIdx.x * K * K;
    if (tid < K)
        for (int c = 0; c < K; ++c) out[(size_t)tid * K + c] = acc[c * TILE_M + tid];
    if (GATHER && tid < ld) plan.bsum[(size_t)blockIdx.x * ld + tid] = colsum;
    __syncthreads();
    if (warp == 0) tmem_dealloc(d_tmem, tmem_cols);
}

// power-of-two TMEM columns for `accs` accumulators of ld columns each
static uint32_t tmem_columns(int ld, int accs) {
    const int need = ld * accs;
    uint32_t c = 32;
    while ((int)c < need) c <<= 1;
    return c;
}

}  // namespace tc

// ---- entry points used by als.cu ----------------------------------------------------------------------------------
bool tc_shape_ok(int dtype, int ld) { return dtype == CYMF_F32 && ld % 32 == 0 && ld >= 32 && ld <= 128; }

int tc_rows_times_matrix(const float *in, float *const *outs, int n_outs, const float *B, int64_t rows, int ld,
                         cudaStream_t st) {
    static const bool use_tma = [] { const char *v = getenv("CYMF_GEMM_TMA"); return !(v && v[0] == '0'); }();
    if (use_tma) {                                           // TMA-fed, warp-specialised pipeline (tc_gemm_tma.cu)
        const int rc = tc_rows_times_matrix_tma(in, outs, n_outs, B, rows, ld, st);
        if (rc != CYMF_EUNSUPPORTED) return rc;
    }
    tc::MultiOutF mo{};
    mo.n = n_outs;
    for (int d = 0; d < n_outs; ++d) mo.p[d] = outs[d];
    const size_t smem = sizeof(float) * (size_t)(2 * tc::TILE_M * tc::CHUNK_K + 2 * ld * tc::CHUNK_K);
    // static (mbarrier, TMEM slot) + dynamic shared memory can exceed the 48 KB default already at ld = 64
    CYMF_CUDA(cudaFuncSetAttribute(tc::tc_rows_times_matrix_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t blocks = (rows + tc::TILE_M - 1) / tc::TILE_M;
    // one accumulator per reduction chunk: ld = 128 takes all 512 TMEM columns, so one CTA per SM there
    const int64_t cap = (int64_t)sm_count() * (ld >= 96 ? 1 : 3);
    if (blocks > cap) blocks = cap;
    tc::tc_rows_times_matrix_kernel<<<(unsigned)blocks, 128, smem, st>>>(in, mo, B, rows, ld,
                                                                         tc::tmem_columns(ld, ld / tc::CHUNK_K));
    CYMF_LAUNCHED();
    return 0;
}

// Rows of Y per Gram CTA: at least 512; for long matrices whatever spreads them over two CTAs per SM, so that the
// number of K x K partials (and the serial sum over them in gram_finish_kernel: 19.5 k partials = 2.5 GB at 10 M
// rows before) stays below ~300 whatever n is.  Depends on n and the SM count only: deterministic.
static int64_t tc_gram_slab_rows(int64_t n) {
    const int64_t ctas = 2 * (int64_t)sm_count();
    int64_t rows = ((n + ctas - 1) / ctas + 31) / 32 * 32;
    return rows < tc::TC_GRAM_SLAB ? tc::TC_GRAM_SLAB : rows;
}
int64_t tc_gram_slabs(int64_t n) { const int64_t r = tc_gram_slab_rows(n); return (n + r - 1) / r; }

int tc_gram_partial(const float *Y, int64_t n, int K, int ld, double *partial, cudaStream_t st) {
    const int64_t slabs = tc_gram_slabs(n);
    if (slabs == 0) return 0;
    const size_t smem = sizeof(float) * (size_t)(2 * tc::TILE_M * tc::CHUNK_K) + sizeof(double) * (size_t)ld * tc::TILE_M;
    CYMF_CUDA(cudaFuncSetAttribute(tc::tc_gram_partial_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc::tc_gram_partial_kernel<false><<<(unsigned)slabs, 128, smem, st>>>(Y, n, tc_gram_slab_rows(n), K, ld, partial,
                                                                         tc::tmem_columns(ld, 1), tc::GatherPlan{});
    CYMF_LAUNCHED();
    return 0;
}

// per-slab partials of sum_{c in row} y_c y_c^T (+ column sums) for the heavy rows order[0 .. n_heavy) of a CSR
int tc_gram_gather(const float *Y, const int64_t *indptr, const int32_t *indices, const int32_t *order,
                   const int32_t *first_slab, int n_heavy, int n_slabs, int K, int ld, double *partial, double *bsum,
                   cudaStream_t st) {
    if (n_slabs <= 0) return 0;
    const size_t smem = sizeof(float) * (size_t)(2 * tc::TILE_M * tc::CHUNK_K) + sizeof(double) * (size_t)ld * tc::TILE_M;
    CYMF_CUDA(cudaFuncSetAttribute(tc::tc_gram_partial_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc::GatherPlan plan{indptr, indices, order, first_slab, n_heavy, bsum};
    tc::tc_gram_partial_kernel<true><<<(unsigned)n_slabs, 128, smem, st>>>(Y, 0, 0, K, ld, partial, tc::tmem_columns(ld, 1), plan);
    CYMF_LAUNCHED();
    return 0;
}

}  // namespace cymf
