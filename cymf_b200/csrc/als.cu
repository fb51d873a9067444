// als.cu -- WMF alternating-least-squares half sweep (replaces WMF._als, cymf/wmf.pyx:136-174, and the dgesv
// call behind cymf/linalg.pyx:144-163).
//
// Reference, per row r with item set S_r:   (G + (w-1) sum_{i in S_r} y_i y_i^T) x_r = w sum_{i in S_r} y_i,
// G = Y^T Y + wd I, solved by dense LU after materialising the K x K matrix (O(|S_r| K^2) scalar work).
// Here the matrix is never formed:
//   gram_partial_kernel / gram_finish_kernel : G (wmf.pyx:142-143).  Row slabs -> per-slab K x K partials (f32 or
//       f64 products, f64 cross-slab sum) -> deterministic reduction; in multi-GPU runs the partial of a rank's own
//       row block is all-reduced before wd*I is added.  G is handed to the solver as [ld, ld] with zero padding.
//   als_cg_kernel : one 128-thread CTA per row, conjugate gradient on A p = G p + (w-1) sum_i y_i (y_i . p).
//       The row's item vectors are staged once in shared memory (as many as fit the staging budget; the rest is
//       re-read through L2 each iteration).  Every lane owns a 1/2/4-element slice of the K-vectors (128-bit
//       shared/global accesses at K=128); the four warps split the items AND the rows of G, so one per-warp partial
//       vector carries both terms; dots of four items are reduced with one 10-shuffle butterfly.  The residual
//       recurrence runs until |r| <= tol |b| (warm start from the current x_r).  Rows come from a heaviest-first
//       work queue.
//   Algorithmic bytes per half sweep: N (K s + 4) + rows (K s + 8) + n K s.
#include <math.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace cymf {

// ---- Gram ------------------------------------------------------------------------------------------------------
constexpr int GRAM_ROWS = 32;        // rows staged per step
constexpr int GRAM_SLAB = 512;       // rows per CTA

template <typename T, int TK>        // each of 16x16 threads owns a TK x TK block of outputs (K <= 16*TK)
__global__ void __launch_bounds__(256) gram_partial_kernel(const T *__restrict__ Y, int64_t n, int K, int ld,
                                                           double *__restrict__ partial) {
    __shared__ T tile[GRAM_ROWS][16 * TK + 1];
    const int tj = threadIdx.x & 15, ti = threadIdx.x >> 4;
    double acc[TK][TK];
#pragma unroll
    for (int a = 0; a < TK; ++a)
#pragma unroll
        for (int b = 0; b < TK; ++b) acc[a][b] = 0.0;
    const int64_t r0 = (int64_t)blockIdx.x * GRAM_SLAB;
    const int64_t r1 = r0 + GRAM_SLAB < n ? r0 + GRAM_SLAB : n;
    for (int64_t base = r0; base < r1; base += GRAM_ROWS) {
        for (int t = threadIdx.x; t < GRAM_ROWS * K; t += 256) {
            const int rr = t / K, c = t - rr * K;
            tile[rr][c] = base + rr < r1 ? Y[(size_t)(base + rr) * ld + c] : T(0);
        }
        __syncthreads();
        if constexpr (sizeof(T) == 8) {              // f64: accumulate straight into the f64 sums
#pragma unroll 2
            for (int rr = 0; rr < GRAM_ROWS; ++rr) {
                T ya[TK], yb[TK];
#pragma unroll
                for (int a = 0; a < TK; ++a) { ya[a] = tile[rr][ti + 16 * a]; yb[a] = tile[rr][tj + 16 * a]; }
#pragma unroll
                for (int a = 0; a < TK; ++a)
#pragma unroll
                    for (int b = 0; b < TK; ++b) acc[a][b] += ya[a] * yb[b];
            }
        } else {                                     // f32: 32-row partial products in f32, summed in f64
            T part[TK][TK];
#pragma unroll
            for (int a = 0; a < TK; ++a)
#pragma unroll
                for (int b = 0; b < TK; ++b) part[a][b] = T(0);
#pragma unroll 4
            for (int rr = 0; rr < GRAM_ROWS; ++rr) {
                T ya[TK], yb[TK];
#pragma unroll
                for (int a = 0; a < TK; ++a) { ya[a] = tile[rr][ti + 16 * a]; yb[a] = tile[rr][tj + 16 * a]; }
#pragma unroll
                for (int a = 0; a < TK; ++a)
#pragma unroll
                    for (int b = 0; b < TK; ++b) part[a][b] += ya[a] * yb[b];
            }
#pragma unroll
            for (int a = 0; a < TK; ++a)
#pragma unroll
                for (int b = 0; b < TK; ++b) acc[a][b] += (double)part[a][b];
        }
        __syncthreads();
    }
    double *out = partial + (size_t)blockIdx.x * K * K;
#pragma unroll
    for (int a = 0; a < TK; ++a)
#pragma unroll
        for (int b = 0; b < TK; ++b) {
            const int i = ti + 16 * a, j = tj + 16 * b;
            if (i < K && j < K) out[i * K + j] = acc[a][b];
        }
}

// Sum the slab partials in slab order (deterministic), optionally add wd on the diagonal.  out64: dense [K, K] f64;
// outT: [ld, ld] of T with zero padding (the layout the CG kernel reads).
template <typename T>
__global__ void gram_finish_kernel(const double *__restrict__ partial, int slabs, int K, int ld, double wd,
                                   double *__restrict__ out64, T *__restrict__ outT) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ld * ld) return;
    const int i = t / ld, j = t - i * ld;
    double s = 0.0;
    if (i < K && j < K) {
        for (int b = 0; b < slabs; ++b) s += partial[(size_t)b * K * K + i * K + j];
        if (i == j) s += wd;
        if (out64) out64[i * K + j] = s;
    }
    if (outT) outT[t] = (T)s;
}

// ---- CG row solver ---------------------------------------------------------------------------------------------
template <typename T> struct AlsArgs {
    const int64_t *indptr;
    const int32_t *indices;
    const int32_t *order;       // rows to solve, heaviest first
    int32_t n_solve;
    T *X;                       // [rows, ld]  solved in place (warm start = current content)
    const T *Y;                 // [n, ld]
    const T *G;                 // [ld, ld]  Y^T Y + wd I, zero padded
    int32_t ld, stage_rows, max_iter;
    T weight, tol2;
    int32_t *queue;             // work-queue head (zeroed before launch)
    unsigned long long *stats;  // [0] CG iterations summed over rows, [1] rows that hit max_iter (may be NULL)
};

// VW contiguous elements, naturally aligned (VW * sizeof(T) up to 32 bytes)
template <int VW> __device__ __forceinline__ void ld_vec(const float *p, float (&v)[VW]) {
    if constexpr (VW == 4) { const float4 t = *reinterpret_cast<const float4 *>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else if constexpr (VW == 2) { const float2 t = *reinterpret_cast<const float2 *>(p); v[0] = t.x; v[1] = t.y; }
    else v[0] = *p;
}
template <int VW> __device__ __forceinline__ void ld_vec(const double *p, double (&v)[VW]) {
    if constexpr (VW == 4) {
        const double2 a = *reinterpret_cast<const double2 *>(p), b = *reinterpret_cast<const double2 *>(p + 2);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else if constexpr (VW == 2) { const double2 t = *reinterpret_cast<const double2 *>(p); v[0] = t.x; v[1] = t.y; }
    else v[0] = *p;
}
template <int VW> __device__ __forceinline__ void ldg_vec(const float *p, float (&v)[VW]) {
    if constexpr (VW == 4) { const float4 t = __ldg(reinterpret_cast<const float4 *>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else if constexpr (VW == 2) { const float2 t = __ldg(reinterpret_cast<const float2 *>(p)); v[0] = t.x; v[1] = t.y; }
    else v[0] = __ldg(p);
}
template <int VW> __device__ __forceinline__ void ldg_vec(const double *p, double (&v)[VW]) {
    if constexpr (VW == 4) {
        const double2 a = __ldg(reinterpret_cast<const double2 *>(p)), b = __ldg(reinterpret_cast<const double2 *>(p + 2));
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else if constexpr (VW == 2) { const double2 t = __ldg(reinterpret_cast<const double2 *>(p)); v[0] = t.x; v[1] = t.y; }
    else v[0] = __ldg(p);
}
template <int VW> __device__ __forceinline__ void st_vec(float *p, const float (&v)[VW]) {
    if constexpr (VW == 4) *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    else if constexpr (VW == 2) *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]);
    else *p = v[0];
}
template <int VW> __device__ __forceinline__ void st_vec(double *p, const double (&v)[VW]) {
    if constexpr (VW == 4) {
        *reinterpret_cast<double2 *>(p) = make_double2(v[0], v[1]);
        *reinterpret_cast<double2 *>(p + 2) = make_double2(v[2], v[3]);
    } else if constexpr (VW == 2) *reinterpret_cast<double2 *>(p) = make_double2(v[0], v[1]);
    else *p = v[0];
}

template <typename T> __device__ __forceinline__ T warp_allsum(T v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

// Sums of four per-lane partials over the warp, each returned to every lane: 10 shuffles instead of 20.
template <typename T>
__device__ __forceinline__ void warp_allsum4(T &d0, T &d1, T &d2, T &d3, int lane) {
    const bool hi16 = lane & 16, hi8 = lane & 8;
    T k0 = (hi16 ? d1 : d0) + __shfl_xor_sync(0xffffffffu, hi16 ? d0 : d1, 16);   // lanes<16: d0, lanes>=16: d1
    T k1 = (hi16 ? d3 : d2) + __shfl_xor_sync(0xffffffffu, hi16 ? d2 : d3, 16);   // lanes<16: d2, lanes>=16: d3
    T k = (hi8 ? k1 : k0) + __shfl_xor_sync(0xffffffffu, hi8 ? k0 : k1, 8);       // (bit4,bit3): 00 d0, 01 d2, 10 d1, 11 d3
    k += __shfl_xor_sync(0xffffffffu, k, 4);
    k += __shfl_xor_sync(0xffffffffu, k, 2);
    k += __shfl_xor_sync(0xffffffffu, k, 1);
    d0 = __shfl_sync(0xffffffffu, k, 0);
    d2 = __shfl_sync(0xffffffffu, k, 8);
    d1 = __shfl_sync(0xffffffffu, k, 16);
    d3 = __shfl_sync(0xffffffffu, k, 24);
}

constexpr int CG_VEC = 128;           // capacity of the shared K-vectors (ld <= 128)

template <typename T, int VW>         // VW = elements of a K-vector per lane (1: ld<=32, 2: ld<=64, 4: ld<=128)
__global__ void __launch_bounds__(128) als_cg_kernel(const AlsArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *p_s = reinterpret_cast<T *>(smem_raw);     // [CG_VEC]     search direction, readable by all warps
    T *part = p_s + CG_VEC;                       // [4][CG_VEC]  per-warp partial A p
    T *red = part + 4 * CG_VEC;                   // [2][4]       block reductions, double buffered
    T *Ys = red + 8;                              // [stage_rows][ld]
    __shared__ int row_slot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ld = a.ld;
    const int kq = lane * VW;                     // this lane's slice of every K-vector
    const bool lane_on = kq < ld;
    const bool own = tid < ld;                    // thread k owns element k of x, r, p
    const int jn = (ld + 3) >> 2, j0 = warp * jn, j1 = (j0 + jn < ld) ? j0 + jn : ld;   // this warp's rows of G
    int flip = 0;

    auto block_sum = [&](T v) -> T {
        v = warp_allsum(v);
        if (lane == 0) red[flip * 4 + warp] = v;
        __syncthreads();
        const T s = (red[flip * 4] + red[flip * 4 + 1]) + (red[flip * 4 + 2] + red[flip * 4 + 3]);
        flip ^= 1;
        return s;
    };

    for (;;) {
        if (tid == 0) row_slot = atomicAdd(a.queue, 1);
        __syncthreads();
        const int slot = row_slot;
        __syncthreads();
        if (slot >= a.n_solve) break;
        const int r = a.order[slot];
        const int64_t lo = a.indptr[r];
        const int nnz = (int)(a.indptr[r + 1] - lo);
        T *xr = a.X + (size_t)r * ld;
        if (nnz == 0) {                                                        // wmf.pyx:154-156
            if (own) xr[tid] = T(0);
            continue;
        }
        const int ns = nnz < a.stage_rows ? nnz : a.stage_rows;
        const int32_t *idx = a.indices + lo;

        auto load_item = [&](int i, T (&v)[VW]) {                             // item vector slice, zeros past the row
#pragma unroll
            for (int e = 0; e < VW; ++e) v[e] = T(0);
            if (i < nnz && lane_on) {
                if (i < ns) ld_vec<VW>(Ys + i * ld + kq, v);
                else ldg_vec<VW>(a.Y + (size_t)__ldg(idx + i) * ld + kq, v);
            }
        };

        // A v for the vector in p_s; returns element `tid` (0 for tid >= ld).  One barrier inside.
        auto apply = [&]() -> T {
            T ps[VW], acc[VW];
#pragma unroll
            for (int e = 0; e < VW; ++e) { ps[e] = T(0); acc[e] = T(0); }
            if (lane_on) ld_vec<VW>(p_s + kq, ps);
            for (int i = warp; i < nnz; i += 16) {                             // this warp: items warp, warp+4, ...
                T y0[VW], y1[VW], y2[VW], y3[VW];
                load_item(i, y0); load_item(i + 4, y1); load_item(i + 8, y2); load_item(i + 12, y3);
                T d0 = T(0), d1 = T(0), d2 = T(0), d3 = T(0);
#pragma unroll
                for (int e = 0; e < VW; ++e) { d0 += y0[e] * ps[e]; d1 += y1[e] * ps[e]; d2 += y2[e] * ps[e]; d3 += y3[e] * ps[e]; }
                warp_allsum4(d0, d1, d2, d3, lane);
#pragma unroll
                for (int e = 0; e < VW; ++e) acc[e] += (d0 * y0[e] + d1 * y1[e]) + (d2 * y2[e] + d3 * y3[e]);
            }
            const T wm1 = a.weight - T(1);
#pragma unroll
            for (int e = 0; e < VW; ++e) acc[e] *= wm1;
            if (lane_on) {
                const T *g = a.G + (size_t)j0 * ld + kq;
#pragma unroll 4
                for (int j = j0; j < j1; ++j, g += ld) {                       // + rows [j0, j1) of G p
                    T gv[VW];
                    ldg_vec<VW>(g, gv);
                    const T pj = p_s[j];
#pragma unroll
                    for (int e = 0; e < VW; ++e) acc[e] += gv[e] * pj;
                }
                st_vec<VW>(part + warp * CG_VEC + kq, acc);
            }
            __syncthreads();
            return own ? (part[tid] + part[CG_VEC + tid]) + (part[2 * CG_VEC + tid] + part[3 * CG_VEC + tid]) : T(0);
        };

        // stage the row's item vectors and accumulate b = w * sum y_i (wmf.pyx:163)
        {
            T bacc[VW];
#pragma unroll
            for (int e = 0; e < VW; ++e) bacc[e] = T(0);
            if (lane_on)
                for (int i = warp; i < nnz; i += 4) {
                    T v[VW];
                    ldg_vec<VW>(a.Y + (size_t)__ldg(idx + i) * ld + kq, v);
#pragma unroll
                    for (int e = 0; e < VW; ++e) bacc[e] += v[e];
                    if (i < ns) st_vec<VW>(Ys + i * ld + kq, v);
                }
            if (lane_on) st_vec<VW>(part + warp * CG_VEC + kq, bacc);
        }
        __syncthreads();
        T b = T(0), x = T(0);
        if (own) {
            b = a.weight * ((part[tid] + part[CG_VEC + tid]) + (part[2 * CG_VEC + tid] + part[3 * CG_VEC + tid]));
            x = xr[tid];                                                        // warm start
            p_s[tid] = x;
        }
        const T bb = block_sum(b * b);                                          // barrier: p_s, Ys and part are settled
        unsigned iters = 0;
        bool stalled = false;
        if (bb > T(0)) {
            T res = b - apply();                                                // r0 = b - A x0
            T p = res;
            T rs = block_sum(res * res);
            while (rs > a.tol2 * bb) {
                if ((int)iters >= a.max_iter) { stalled = true; break; }
                if (own) p_s[tid] = p;
                __syncthreads();
                const T Ap = apply();
                const T pAp = block_sum(p * Ap);
                if (!(pAp > T(0))) { stalled = true; break; }
                const T alpha = rs / pAp;
                x += alpha * p;
                res -= alpha * Ap;
                const T rs_new = block_sum(res * res);
                p = res + (rs_new / rs) * p;
                rs = rs_new;
                ++iters;
            }
        } else {
            x = T(0);                                                           // b = 0  =>  x = 0
        }
        if (own) xr[tid] = x;
        if (a.stats && tid == 0) {
            atomicAdd(a.stats, (unsigned long long)iters);
            if (stalled) atomicAdd(a.stats + 1, 1ull);
        }
    }
}

template <typename T> static int gram_impl(const T *Y, int64_t n, int K, int ld, double wd, int add_wd, double *partial,
                                           int64_t partial_capacity, double *out64, T *outT, cudaStream_t st) {
    const int slabs = (int)((n + GRAM_SLAB - 1) / GRAM_SLAB);
    if ((int64_t)slabs * K * K > partial_capacity) { set_error("gram: workspace too small"); return CYMF_EINVAL; }
    if (slabs > 0) {
        if (K <= 32) gram_partial_kernel<T, 2><<<slabs, 256, 0, st>>>(Y, n, K, ld, partial);
        else if (K <= 64) gram_partial_kernel<T, 4><<<slabs, 256, 0, st>>>(Y, n, K, ld, partial);
        else gram_partial_kernel<T, 8><<<slabs, 256, 0, st>>>(Y, n, K, ld, partial);
        CYMF_LAUNCHED();
    }
    gram_finish_kernel<T><<<(ld * ld + 255) / 256, 256, 0, st>>>(partial, slabs, K, ld, add_wd ? wd : 0.0, out64, outT);
    CYMF_LAUNCHED();
    return 0;
}

template <typename T, int VW> static int launch_cg(const AlsArgs<T> &a, cudaStream_t st) {
    const size_t smem = sizeof(T) * ((size_t)5 * CG_VEC + 8 + (size_t)a.stage_rows * a.ld);
    auto kern = als_cg_kernel<T, VW>;
    if (smem > 48 * 1024) CYMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CYMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem));
    if (per_sm < 1) per_sm = 1;
    int64_t blocks = (int64_t)sm_count() * per_sm;
    if (blocks > a.n_solve) blocks = a.n_solve;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, 128, smem, st>>>(a);
    CYMF_LAUNCHED();
    return 0;
}

template <typename T> static int cg_impl(AlsArgs<T> a, cudaStream_t st) {
    CYMF_CUDA(cudaMemsetAsync(a.queue, 0, sizeof(int32_t), st));
    if (a.ld <= 32) return launch_cg<T, 1>(a, st);
    if (a.ld <= 64) return launch_cg<T, 2>(a, st);
    return launch_cg<T, 4>(a, st);
}

}  // namespace cymf

using namespace cymf;

extern "C" int64_t cymf_gram_workspace_doubles(int64_t n, int32_t K) {
    return ((n + GRAM_SLAB - 1) / GRAM_SLAB) * (int64_t)K * K;
}

extern "C" int cymf_gram_dev(const void *Y, int dtype, int64_t n, int32_t K, int32_t ld, double weight_decay,
                             int add_weight_decay, double *workspace, int64_t workspace_doubles,
                             double *out_f64, void *out_native, void *stream) {
    CYMF_REQUIRE(Y && workspace && (out_f64 || out_native), "null pointer");
    CYMF_REQUIRE(n >= 0 && K > 0 && K <= 128 && ld >= K && ld % 4 == 0 && ld <= 128,
                 "bad shape (WMF supports num_components <= 128, ld a multiple of 4)");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CYMF_F32)
        return gram_impl<float>((const float *)Y, n, K, ld, weight_decay, add_weight_decay, workspace, workspace_doubles,
                                out_f64, (float *)out_native, st);
    if (dtype == CYMF_F64)
        return gram_impl<double>((const double *)Y, n, K, ld, weight_decay, add_weight_decay, workspace,
                                 workspace_doubles, out_f64, (double *)out_native, st);
    set_error("gram: unknown dtype %d", dtype);
    return CYMF_EINVAL;
}

// out_native = [ld, ld] zero-padded copy of (in_f64 [K, K] + wd on the diagonal): finishes a Gram matrix that was
// all-reduced in f64
extern "C" int cymf_gram_finalize_dev(const double *in_f64, int dtype, int32_t K, int32_t ld, double weight_decay,
                                      void *out_native, void *stream) {
    CYMF_REQUIRE(in_f64 && out_native && K > 0 && K <= 128 && ld >= K && ld <= 128, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CYMF_F32)
        gram_finish_kernel<float><<<(ld * ld + 255) / 256, 256, 0, st>>>(in_f64, 1, K, ld, weight_decay, nullptr,
                                                                         (float *)out_native);
    else
        gram_finish_kernel<double><<<(ld * ld + 255) / 256, 256, 0, st>>>(in_f64, 1, K, ld, weight_decay, nullptr,
                                                                          (double *)out_native);
    CYMF_LAUNCHED();
    return 0;
}

extern "C" int cymf_als_cg_dev(const int64_t *indptr, const int32_t *indices, const int32_t *order, int32_t n_solve,
                               void *X, const void *Y, const void *G, int dtype, int32_t K, int32_t ld,
                               double weight, double cg_tol, int32_t cg_max_iter, int32_t stage_rows,
                               int32_t *queue, unsigned long long *stats, void *stream) {
    CYMF_REQUIRE(indptr && indices && order && X && Y && G && queue, "null pointer");
    CYMF_REQUIRE(K > 0 && K <= 128 && ld >= K && ld % 4 == 0 && ld <= 128,
                 "bad shape (WMF supports num_components <= 128, ld a multiple of 4)");
    CYMF_REQUIRE(cg_tol > 0 && cg_max_iter > 0, "bad CG parameters");
    if (n_solve <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t es = dtype == CYMF_F32 ? 4 : 8;
    if (stage_rows <= 0) {                       // auto: ~32 KB of staged item vectors per CTA
        stage_rows = (int32_t)(32 * 1024 / (es * ld));
        if (stage_rows < 8) stage_rows = 8;
    }
    if (dtype == CYMF_F32) {
        AlsArgs<float> a{indptr, indices, order, n_solve, (float *)X, (const float *)Y, (const float *)G, ld,
                         stage_rows, cg_max_iter, (float)weight, (float)(cg_tol * cg_tol), queue, stats};
        return cg_impl<float>(a, st);
    }
    if (dtype == CYMF_F64) {
        AlsArgs<double> a{indptr, indices, order, n_solve, (double *)X, (const double *)Y, (const double *)G, ld,
                          stage_rows, cg_max_iter, weight, cg_tol * cg_tol, queue, stats};
        return cg_impl<double>(a, st);
    }
    set_error("als: unknown dtype %d", dtype);
    return CYMF_EINVAL;
}

extern "C" int cymf_als_half_host(const int32_t *indptr, const int32_t *indices, double *X, const double *Y,
                                  int64_t rows, int64_t n, int32_t K, double weight_decay, double weight,
                                  int dtype, double cg_tol, int32_t cg_max_iter, int64_t *cg_iterations_out) {
    CYMF_REQUIRE(indptr && indices && X && Y, "null pointer");
    CYMF_REQUIRE(rows > 0 && n > 0 && K > 0 && K <= 128, "bad shape (WMF supports num_components <= 128)");
    CYMF_REQUIRE(dtype == CYMF_F32 || dtype == CYMF_F64, "unknown dtype");
    int ndev = 0;
    CYMF_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev < 1) { set_error("no CUDA device: cymf_b200 has no CPU fallback"); return CYMF_EUNSUPPORTED; }
    if (cg_tol <= 0) cg_tol = dtype == CYMF_F32 ? 1e-6 : 1e-10;
    if (cg_max_iter <= 0) cg_max_iter = 2 * K;
    const size_t es = dtype == CYMF_F32 ? 4 : 8;
    const int32_t ld = (K + 3) / 4 * 4;
    const int64_t nnz = indptr[rows];
    std::vector<int64_t> ip64((size_t)rows + 1);
    std::vector<int32_t> order((size_t)rows);
    for (int64_t r = 0; r <= rows; ++r) ip64[(size_t)r] = indptr[r];
    for (int64_t r = 0; r < rows; ++r) order[(size_t)r] = (int32_t)r;
    std::stable_sort(order.begin(), order.end(), [&](int32_t p, int32_t q) {
        return indptr[p + 1] - indptr[p] > indptr[q + 1] - indptr[q];
    });
    DeviceArena mem;
    cudaStream_t st = nullptr;
    double *stage, *ws, *g64;
    void *dX, *dY, *dG;
    int64_t *d_ip;
    int32_t *d_ix, *d_order, *d_queue;
    unsigned long long *d_stats;
    const int64_t big = (rows > n ? rows : n) * K;
    const int64_t wsn = cymf_gram_workspace_doubles(n, K);
    CYMF_TRY(mem.get(&stage, (size_t)big * 8));
    CYMF_TRY(mem.get(&ws, (size_t)wsn * 8));
    CYMF_TRY(mem.get(&g64, (size_t)K * K * 8));
    CYMF_TRY(mem.get((char **)&dX, (size_t)rows * ld * es));
    CYMF_TRY(mem.get((char **)&dY, (size_t)n * ld * es));
    CYMF_TRY(mem.get((char **)&dG, (size_t)ld * ld * es));
    CYMF_TRY(mem.get(&d_ip, ((size_t)rows + 1) * 8));
    CYMF_TRY(mem.get(&d_ix, (size_t)nnz * 4));
    CYMF_TRY(mem.get(&d_order, (size_t)rows * 4));
    CYMF_TRY(mem.get(&d_queue, 4));
    CYMF_TRY(mem.get(&d_stats, 16));
    CYMF_CUDA(cudaMemcpyAsync(stage, Y, (size_t)n * K * 8, cudaMemcpyHostToDevice, st));
    CYMF_TRY(cymf_pack_rows_dev(stage, dY, dtype, n, K, ld, st));
    CYMF_CUDA(cudaMemcpyAsync(stage, X, (size_t)rows * K * 8, cudaMemcpyHostToDevice, st));
    CYMF_TRY(cymf_pack_rows_dev(stage, dX, dtype, rows, K, ld, st));
    CYMF_CUDA(cudaMemcpyAsync(d_ip, ip64.data(), ((size_t)rows + 1) * 8, cudaMemcpyHostToDevice, st));
    CYMF_CUDA(cudaMemcpyAsync(d_ix, indices, (size_t)nnz * 4, cudaMemcpyHostToDevice, st));
    CYMF_CUDA(cudaMemcpyAsync(d_order, order.data(), (size_t)rows * 4, cudaMemcpyHostToDevice, st));
    CYMF_CUDA(cudaMemsetAsync(d_stats, 0, 16, st));
    CYMF_TRY(cymf_gram_dev(dY, dtype, n, K, ld, weight_decay, 1, ws, wsn, g64, dG, st));
    CYMF_TRY(cymf_als_cg_dev(d_ip, d_ix, d_order, (int32_t)rows, dX, dY, dG, dtype, K, ld, weight, cg_tol,
                             cg_max_iter, 0, d_queue, d_stats, st));
    CYMF_TRY(cymf_unpack_rows_dev(dX, stage, dtype, rows, K, ld, st));
    CYMF_CUDA(cudaMemcpyAsync(X, stage, (size_t)rows * K * 8, cudaMemcpyDeviceToHost, st));
    unsigned long long stats[2] = {0, 0};
    CYMF_CUDA(cudaMemcpyAsync(stats, d_stats, 16, cudaMemcpyDeviceToHost, st));
    CYMF_CUDA(cudaStreamSynchronize(st));
    if (cg_iterations_out) *cg_iterations_out = (int64_t)stats[0];
    return 0;
}
