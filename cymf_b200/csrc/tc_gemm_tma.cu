// tc_gemm_tma.cu -- out[r, :] = in[r, :] * B on tcgen05 with TMA-fed, warp-specialised, multi-stage pipelining
// (the three changes of variables of the WMF half sweep: Y~ = Y L^-T, X~ = X L, X = X~ L^-1; f32, ld in {32,64,96,128}).
//
// tc_gemm.cu's first version of this kernel stages every operand chunk with CUDA cores, fences, issues 12 MMAs and
// waits for them before touching the next chunk: the tensor pipe sat at 12 % (profiles/r1_tcgen05_gemms_ncu_full.txt).
// Here one persistent CTA per SM runs four roles in parallel:
//   * warp 8 (one lane)  : TMA producer.  cp.async.bulk.tensor.2d brings the [128 rows x 32 columns] chunk of `in`
//                          into a 3-stage ring, already in the canonical K-major SWIZZLE_128B operand layout
//                          (mbarrier expect_tx / complete_tx; rows past the end are zero-filled by the TMA unit);
//   * warps 4-7          : lo conversion.  The tensor core reads tf32 operands from 32-bit containers and ignores the
//                          13 low mantissa bits, so the TMA-written tile IS the hi operand; only lo = a - hi is
//                          computed (element-wise, same swizzled positions) for the 3xTF32 scheme hi*lo + lo*hi + hi*hi;
//   * warp 9 (one lane)  : MMA issuer.  12 tcgen05.mma kind::tf32 per chunk against the B operand (B^T chunks,
//                          hi and lo, staged ONCE per CTA); tcgen05.commit frees the stage for the producer;
//   * warps 0-3          : epilogue.  tcgen05.ld of the finished tile (two accumulation chains per tile, added in
//                          registers) -> 128-bit stores to EVERY destination (own buffer and, over NVLink, every peer's:
//                          GEMM + all-gather in one kernel).  Accumulators are double buffered across tiles, so the
//                          epilogue of tile t overlaps the MMAs of tile t + 1.
#include <cuda.h>

#include "tc_common.cuh"

namespace cymf {
namespace tc {

constexpr int TG_STAGES = 3;
constexpr int TG_THREADS = 320;
constexpr int TG_MAX_DESTS = 8;
struct TgOuts { float *p[TG_MAX_DESTS]; int n; };

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int32_t c0, int32_t c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// K-major SWIZZLE_128B operand: rows of 128 bytes, 8-row groups 1024 bytes apart; start advanced by 32 bytes per K = 8 slice
__device__ __forceinline__ uint64_t sw128_desc(uint32_t saddr) {
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)1 << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(TG_THREADS, 1)
tc_rows_times_matrix_tma_kernel(const __grid_constant__ CUtensorMap in_map, const TgOuts outs, const float *__restrict__ B,
                                int64_t rows, int ld, uint32_t tmem_cols) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int nk = ld / CHUNK_K;                                  // reduction chunks per tile
    const int nacc = nk >= 2 ? 2 : 1;                             // accumulation chains (TMEM accumulators) per tile
    float *const a_raw = reinterpret_cast<float *>(smem_raw);     // [TG_STAGES][128 x 32]  TMA destination = hi operand
    float *const a_lo = a_raw + TG_STAGES * TILE_M * CHUNK_K;     // [TG_STAGES][128 x 32]
    float *const b_hi = a_lo + TG_STAGES * TILE_M * CHUNK_K;      // [nk][ld x 32]  B^T chunks, K-major no swizzle
    float *const b_lo = b_hi + nk * ld * CHUNK_K;                 // [nk][ld x 32]
    __shared__ uint64_t full_raw[TG_STAGES], conv[TG_STAGES], empty[TG_STAGES], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&tmem_slot, tmem_cols);
    if (tid == 0) {
        for (int s = 0; s < TG_STAGES; ++s) { mbar_init(&full_raw[s], 1); mbar_init(&conv[s], 128); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 4); }
    }
    // B operand, once per CTA: element (n, k) of chunk kc = B[32 kc + k][n]
    const int b_groups = ld / 8;
    for (int t = tid; t < nk * ld; t += TG_THREADS) {
        const int kc = t / ld, n = t - kc * ld;
        const float *src = B + (size_t)kc * CHUNK_K * ld + n;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const float4 v = make_float4(__ldg(src + (4 * q) * ld), __ldg(src + (4 * q + 1) * ld), __ldg(src + (4 * q + 2) * ld),
                                         __ldg(src + (4 * q + 3) * ld));
            const float4 h = tf32_hi(v);
            const int o = kc * ld * CHUNK_K + tile_off(n, q, b_groups);
            *reinterpret_cast<float4 *>(b_hi + o) = v;            // read as tf32: the low mantissa bits are ignored
            *reinterpret_cast<float4 *>(b_lo + o) = sub4(v, h);
        }
    }
    fence_async_smem();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t d_tmem = tmem_slot;
    const int64_t n_tiles = (rows + TILE_M - 1) / TILE_M;

    if (warp == 8) {                                              // ---- TMA producer
        if (elect_one()) {
            uint32_t cg = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                for (int kc = 0; kc < nk; ++kc, ++cg) {
                    const uint32_t s = cg % TG_STAGES, ph = (cg / TG_STAGES) & 1u;
                    mbar_wait(&empty[s], ph ^ 1u);                // (a fresh barrier passes: nothing to wait for yet)
                    mbar_expect_tx(&full_raw[s], TILE_M * CHUNK_K * 4);
                    tma_load_2d(a_raw + s * TILE_M * CHUNK_K, &in_map, kc * CHUNK_K, (int32_t)(tile * TILE_M), &full_raw[s]);
                }
        }
    } else if (warp == 9) {                                       // ---- MMA issuer
        if (elect_one()) {                                        // (not `lane == 0`: see elect_one in tc_common.cuh)
            const uint32_t idesc = idesc_tf32(ld);
            const uint32_t lbo_b = b_groups * 128;
            uint32_t cg = 0, lt = 0;
            for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
                const uint32_t ab = lt & 1u;
                mbar_wait(&acc_empty[ab], ((lt >> 1) & 1u) ^ 1u);  // the epilogue has read this accumulator pair
                fence_after_sync();
                for (int kc = 0; kc < nk; ++kc, ++cg) {
                    const uint32_t s = cg % TG_STAGES, ph = (cg / TG_STAGES) & 1u;
                    mbar_wait(&conv[s], ph);
                    fence_after_sync();
                    const int chain = kc * nacc / nk;             // chunks {0,1} -> chain 0, {2,3} -> chain 1 (ld = 128)
                    const bool first = kc == 0 || (kc * nacc / nk) != ((kc - 1) * nacc / nk);
                    const uint32_t d = d_tmem + (uint32_t)((ab * nacc + chain) * ld);
                    const uint32_t ah = smem_u32(a_raw + s * TILE_M * CHUNK_K), al = smem_u32(a_lo + s * TILE_M * CHUNK_K);
                    const float *bh = b_hi + kc * ld * CHUNK_K, *bl = b_lo + kc * ld * CHUNK_K;
#pragma unroll
                    for (int ks = 0; ks < CHUNK_K / 8; ++ks) {
                        const uint64_t dah = sw128_desc(ah + ks * 32), dal = sw128_desc(al + ks * 32);
                        const uint64_t dbh = smem_desc(bh + ks * 2 * (lbo_b >> 2), lbo_b, 128);
                        const uint64_t dbl = smem_desc(bl + ks * 2 * (lbo_b >> 2), lbo_b, 128);
                        mma_tf32(d, dah, dbl, idesc, (first && ks == 0) ? 0u : 1u);      // small terms first
                        mma_tf32(d, dal, dbh, idesc, 1u);
                        mma_tf32(d, dah, dbh, idesc, 1u);
                    }
                    mma_commit(&empty[s]);                        // the stage may be refilled once these MMAs are done
                }
                mma_commit(&acc_full[ab]);
            }
        }
    } else if (warp >= 4) {                                       // ---- lo conversion (warps 4-7)
        const int ct = tid - 128;
        uint32_t cg = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
            for (int kc = 0; kc < nk; ++kc, ++cg) {
                const uint32_t s = cg % TG_STAGES, ph = (cg / TG_STAGES) & 1u;
                mbar_wait(&full_raw[s], ph);
                const float4 *src = reinterpret_cast<const float4 *>(a_raw + s * TILE_M * CHUNK_K);
                float4 *dst = reinterpret_cast<float4 *>(a_lo + s * TILE_M * CHUNK_K);
#pragma unroll
                for (int j = 0; j < 8; ++j) {                     // element-wise: the swizzle is the same in both tiles
                    const float4 v = src[ct + 128 * j];
                    dst[ct + 128 * j] = sub4(v, tf32_hi(v));
                }
                fence_async_smem();
                mbar_arrive(&conv[s]);
            }
    } else {                                                      // ---- epilogue (warps 0-3 = TMEM lane quarters)
        uint32_t lt = 0;
        for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
            const uint32_t ab = lt & 1u;
            mbar_wait(&acc_full[ab], (lt >> 1) & 1u);
            fence_after_sync();
            const int64_t row = tile * TILE_M + tid;
            const uint32_t t0 = d_tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(ab * nacc * ld);
            for (int c0 = 0; c0 < ld; c0 += 32) {
                float v[32];
                tmem_load32(t0 + (uint32_t)c0, v);
                if (nacc == 2) {
                    float u[32];
                    tmem_load32(t0 + (uint32_t)(ld + c0), u);
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
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[ab]);
        }
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(d_tmem, tmem_cols);
}

}  // namespace tc

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// Returns CYMF_EUNSUPPORTED when the driver cannot encode the tensor map (the caller then uses the CUDA-core-staged kernel).
int tc_rows_times_matrix_tma(const float *in, float *const *outs, int n_outs, const float *B, int64_t rows, int ld,
                             cudaStream_t st) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc || ((uintptr_t)in & 15u)) return CYMF_EUNSUPPORTED;
    alignas(64) CUtensorMap map;
    const cuuint64_t gdim[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)tc::CHUNK_K, (cuuint32_t)tc::TILE_M};
    const cuuint32_t estr[2] = {1, 1};
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(in), gdim, gstride, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return CYMF_EUNSUPPORTED;
    tc::TgOuts mo{};
    mo.n = n_outs;
    for (int d = 0; d < n_outs; ++d) mo.p[d] = outs[d];
    const int nk = ld / tc::CHUNK_K, nacc = nk >= 2 ? 2 : 1;
    const size_t smem = sizeof(float) * ((size_t)2 * tc::TG_STAGES * tc::TILE_M * tc::CHUNK_K + (size_t)2 * nk * ld * tc::CHUNK_K);
    CYMF_CUDA(cudaFuncSetAttribute(tc::tc_rows_times_matrix_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    uint32_t cols = 32;
    while ((int)cols < 2 * nacc * ld) cols <<= 1;
    int64_t blocks = (rows + tc::TILE_M - 1) / tc::TILE_M;
    if (blocks > sm_count()) blocks = sm_count();
    tc::tc_rows_times_matrix_tma_kernel<<<(unsigned)blocks, tc::TG_THREADS, smem, st>>>(map, mo, B, rows, ld, cols);
    CYMF_LAUNCHED();
    return 0;
}

}  // namespace cymf
