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

#include <stdlib.h>

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

// K > 128 (the reference has no limit on num_components, cymf/wmf.pyx:44): one CTA per (slab, 64 x 64 output block).
template <typename T>
__global__ void __launch_bounds__(256) gram_wide_kernel(const T *__restrict__ Y, int64_t n, int K, int ld,
                                                        double *__restrict__ partial) {
    __shared__ T ta[GRAM_ROWS][65], tb[GRAM_ROWS][65];
    const int nb = (K + 63) / 64, bi = blockIdx.y / nb, bj = blockIdx.y % nb;
    const int tj = threadIdx.x & 15, ti = threadIdx.x >> 4;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    const int64_t r0 = (int64_t)blockIdx.x * GRAM_SLAB;
    const int64_t r1 = r0 + GRAM_SLAB < n ? r0 + GRAM_SLAB : n;
    for (int64_t base = r0; base < r1; base += GRAM_ROWS) {
        for (int t = threadIdx.x; t < GRAM_ROWS * 64; t += 256) {
            const int rr = t >> 6, c = t & 63;
            const bool ok = base + rr < r1;
            ta[rr][c] = ok && bi * 64 + c < K ? Y[(size_t)(base + rr) * ld + bi * 64 + c] : T(0);
            tb[rr][c] = ok && bj * 64 + c < K ? Y[(size_t)(base + rr) * ld + bj * 64 + c] : T(0);
        }
        __syncthreads();
        T part[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) part[a][b] = T(0);
#pragma unroll 4
        for (int rr = 0; rr < GRAM_ROWS; ++rr) {
            T ya[4], yb[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) { ya[a] = ta[rr][ti + 16 * a]; yb[a] = tb[rr][tj + 16 * a]; }
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) part[a][b] += ya[a] * yb[b];
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] += (double)part[a][b];      // 32-row products in T, summed in f64
        __syncthreads();
    }
    double *out = partial + (size_t)blockIdx.x * K * K;
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int i = bi * 64 + ti + 16 * a, j = bj * 64 + tj + 16 * b;
            if (i < K && j < K) out[(size_t)i * K + j] = acc[a][b];
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
        // (P + P^T) / 2: exactly symmetric whatever the order in which the partials were accumulated (the 3xTF32
        // tensor-core path adds hi*lo and lo*hi in different MMAs, so P[i][j] and P[j][i] can differ in the last bit)
        for (int b = 0; b < slabs; ++b) {
            const double *p = partial + (size_t)b * K * K;
            s += 0.5 * (p[i * K + j] + p[j * K + i]);
        }
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
    const T *Ginv;              // [ld, ld]  inverse of G (preconditioner) or NULL
    int32_t ld, stage_rows, max_iter;
    T weight, tol2;
    int32_t *queue;             // work-queue head (zeroed before launch)
    unsigned long long *stats;  // [0] CG iterations summed over rows, [1] rows that hit max_iter (may be NULL)
};

// VW contiguous elements, naturally aligned (VW * sizeof(T) up to 32 bytes)
template <int VW> __device__ __forceinline__ void ld_vec(const float *p, float (&v)[VW]) {
    if constexpr (VW == 8) {
        const float4 s = *reinterpret_cast<const float4 *>(p), t = *reinterpret_cast<const float4 *>(p + 4);
        v[0] = s.x; v[1] = s.y; v[2] = s.z; v[3] = s.w; v[4] = t.x; v[5] = t.y; v[6] = t.z; v[7] = t.w;
    } else if constexpr (VW == 4) { const float4 t = *reinterpret_cast<const float4 *>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else if constexpr (VW == 2) { const float2 t = *reinterpret_cast<const float2 *>(p); v[0] = t.x; v[1] = t.y; }
    else v[0] = *p;
}
template <int VW> __device__ __forceinline__ void ld_vec(const double *p, double (&v)[VW]) {
    if constexpr (VW == 8) {
#pragma unroll
        for (int e = 0; e < 8; e += 2) { const double2 t = *reinterpret_cast<const double2 *>(p + e); v[e] = t.x; v[e + 1] = t.y; }
    } else if constexpr (VW == 4) {
        const double2 a = *reinterpret_cast<const double2 *>(p), b = *reinterpret_cast<const double2 *>(p + 2);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else if constexpr (VW == 2) { const double2 t = *reinterpret_cast<const double2 *>(p); v[0] = t.x; v[1] = t.y; }
    else v[0] = *p;
}
template <int VW> __device__ __forceinline__ void ldg_vec(const float *p, float (&v)[VW]) {
    if constexpr (VW == 8) {
        const float4 s = __ldg(reinterpret_cast<const float4 *>(p)), t = __ldg(reinterpret_cast<const float4 *>(p + 4));
        v[0] = s.x; v[1] = s.y; v[2] = s.z; v[3] = s.w; v[4] = t.x; v[5] = t.y; v[6] = t.z; v[7] = t.w;
    } else if constexpr (VW == 4) { const float4 t = __ldg(reinterpret_cast<const float4 *>(p)); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else if constexpr (VW == 2) { const float2 t = __ldg(reinterpret_cast<const float2 *>(p)); v[0] = t.x; v[1] = t.y; }
    else v[0] = __ldg(p);
}
template <int VW> __device__ __forceinline__ void ldg_vec(const double *p, double (&v)[VW]) {
    if constexpr (VW == 8) {
#pragma unroll
        for (int e = 0; e < 8; e += 2) { const double2 t = __ldg(reinterpret_cast<const double2 *>(p + e)); v[e] = t.x; v[e + 1] = t.y; }
    } else if constexpr (VW == 4) {
        const double2 a = __ldg(reinterpret_cast<const double2 *>(p)), b = __ldg(reinterpret_cast<const double2 *>(p + 2));
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else if constexpr (VW == 2) { const double2 t = __ldg(reinterpret_cast<const double2 *>(p)); v[0] = t.x; v[1] = t.y; }
    else v[0] = __ldg(p);
}
template <int VW> __device__ __forceinline__ void st_vec(float *p, const float (&v)[VW]) {
    if constexpr (VW == 8) {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4 *>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else if constexpr (VW == 4) *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    else if constexpr (VW == 2) *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]);
    else *p = v[0];
}
template <int VW> __device__ __forceinline__ void st_vec(double *p, const double (&v)[VW]) {
    if constexpr (VW == 8) {
#pragma unroll
        for (int e = 0; e < 8; e += 2) *reinterpret_cast<double2 *>(p + e) = make_double2(v[e], v[e + 1]);
    } else if constexpr (VW == 4) {
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
constexpr int CG_VEC = 256;           // capacity of the shared K-vectors (ld <= 256: 8 elements per lane)

// element offset of row r of an [n, ld] matrix; r >= 0, so the product is one unsigned wide multiply (IMAD.WIDE.U32)
// instead of the sign-extending 64-bit multiply chain (6 -> 3 instructions per gathered row in the CG loops)
__device__ __forceinline__ size_t yrow_off(int32_t r, int ld) { return (size_t)((uint64_t)(uint32_t)r * (uint32_t)ld); }

// acc + d0 y0 + d1 y1 + d2 y2 + d3 y3 as one chain of four fused multiply-adds (the CG kernel is issue-bound on
// short rows -- 73 % issue-active, profiles/r1_als_cg_c5_scale03_ncu_full.txt -- and the pairwise form cost six
// instructions per element instead of four)
#define CYMF_AXPY4(a, d0, y0, d1, y1, d2, y2, d3, y3) fma_t(d3, y3, fma_t(d2, y2, fma_t(d1, y1, fma_t(d0, y0, a))))
__device__ __forceinline__ float fma_t(float a, float b, float c) { return fmaf(a, b, c); }
__device__ __forceinline__ double fma_t(double a, double b, double c) { return fma(a, b, c); }

// (lo, hi) += (y_lo, y_hi) * d as ONE packed instruction (Blackwell FFMA2, PTX fma.rn.f32x2): two independent IEEE
// fused multiply-adds, so the values are those of the scalar chain.
__device__ __forceinline__ void fma2_acc(float &lo, float &hi, float y_lo, float y_hi, float d) {
    unsigned long long acc, y, dd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(acc) : "f"(lo), "f"(hi));
    asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(y_lo), "f"(y_hi));
    asm("mov.b64 %0, {%1, %1};" : "=l"(dd) : "f"(d));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(y), "l"(dd));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc));
}
// y . p over the lane's slice: packed products, one horizontal add at the end
template <typename T, int VW> __device__ __forceinline__ T dot_slice(const T (&y)[VW], const T (&p)[VW]) {
    if constexpr (sizeof(T) == 4 && VW == 4) {
        float lo = 0.f, hi = 0.f;
        unsigned long long acc, a, b;
        asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(y[0]), "f"(y[1]));
        asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(p[0]), "f"(p[1]));
        asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(acc) : "l"(a), "l"(b));
        asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(y[2]), "f"(y[3]));
        asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(p[2]), "f"(p[3]));
        asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc));
        return lo + hi;
    } else {
        T d = T(0);
#pragma unroll
        for (int e = 0; e < VW; ++e) d += y[e] * p[e];
        return d;
    }
}
// acc[:] += d0 y0[:] + d1 y1[:] + d2 y2[:] + d3 y3[:]
template <typename T, int VW>
__device__ __forceinline__ void axpy4(T (&acc)[VW], T d0, const T (&y0)[VW], T d1, const T (&y1)[VW], T d2,
                                      const T (&y2)[VW], T d3, const T (&y3)[VW]) {
    if constexpr (sizeof(T) == 4 && VW % 2 == 0) {
#pragma unroll
        for (int e = 0; e < VW; e += 2) {
            fma2_acc(acc[e], acc[e + 1], y0[e], y0[e + 1], d0);
            fma2_acc(acc[e], acc[e + 1], y1[e], y1[e + 1], d1);
            fma2_acc(acc[e], acc[e + 1], y2[e], y2[e + 1], d2);
            fma2_acc(acc[e], acc[e + 1], y3[e], y3[e + 1], d3);
        }
    } else {
#pragma unroll
        for (int e = 0; e < VW; ++e) acc[e] = CYMF_AXPY4(acc[e], d0, y0[e], d1, y1[e], d2, y2[e], d3, y3[e]);
    }
}

// NW warps cooperate on one row.  VW = elements of a K-vector per lane (1: ld<=32, 2: ld<=64, 4: ld<=128, 8: ld<=256,
// CTAs of at least 8 warps there: thread k owns element k of x, r, p).
template <typename T, int VW, int NW>
__global__ void __launch_bounds__(32 * NW) als_cg_kernel(const AlsArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T *const p_s = reinterpret_cast<T *>(smem_raw);     // [CG_VEC]      search direction, readable by all warps
    T *const part = p_s + CG_VEC;                       // [NW][CG_VEC]  per-warp partial A p
    T *const red = part + NW * CG_VEC;                  // [2][2 NW]     block reductions, double buffered
    T *const Ys = red + 4 * NW;                         // [stage_rows][ld]
    __shared__ int row_slot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ld = a.ld;
    const int kq = lane * VW;                           // this lane's slice of every K-vector
    const bool lane_on = kq < ld;
    const bool own = tid < ld;                          // thread k owns element k of x, r, p
    const int jn = (ld + NW - 1) / NW, j0 = warp * jn, j1 = (j0 + jn < ld) ? j0 + jn : ld;   // this warp's rows of G
    const T wm1 = a.weight - T(1);
    const bool dense = a.G != nullptr;                  // false: factors were transformed so that G = I
    const T *const Gw = dense ? a.G + (size_t)j0 * ld + kq : nullptr;
    const bool pre = a.Ginv != nullptr;                 // G^-1-preconditioned CG
    const T *const Giw = pre ? a.Ginv + (size_t)j0 * ld + kq : nullptr;
    const T *const Yq = a.Y + kq;
    const int stage_rows = a.stage_rows;
    int flip = 0;

    for (;;) {
        if (tid == 0) row_slot = atomicAdd(a.queue, 1);
        __syncthreads();
        const int slot = row_slot;
        __syncthreads();
        if (slot >= a.n_solve) break;
        const int r = a.order[slot];
        const int64_t lo = a.indptr[r];
        const int nnz = (int)(a.indptr[r + 1] - lo);
        T *const xr = a.X + (size_t)r * ld;
        if (nnz == 0) {                                                        // wmf.pyx:154-156
            if (own) xr[tid] = T(0);
            continue;
        }
        const int ns = nnz < stage_rows ? nnz : stage_rows;
        const int32_t *const idx = a.indices + lo;

#define CYMF_BLOCK_SUM(out, expr)                                                                  \
    {                                                                                              \
        T v_ = warp_allsum<T>(expr);                                                               \
        if (lane == 0) red[flip * 2 * NW + warp] = v_;                                             \
        __syncthreads();                                                                           \
        T s_ = T(0);                                                                               \
        _Pragma("unroll") for (int w_ = 0; w_ < NW; ++w_) s_ += red[flip * 2 * NW + w_];           \
        flip ^= 1;                                                                                 \
        out = s_;                                                                                  \
    }

#define CYMF_BLOCK_SUM2(out0, out1, expr0, expr1)                                                  \
    {                                                                                              \
        T v0_ = warp_allsum<T>(expr0), v1_ = warp_allsum<T>(expr1);                                \
        if (lane == 0) { red[flip * 2 * NW + warp] = v0_; red[flip * 2 * NW + NW + warp] = v1_; }  \
        __syncthreads();                                                                           \
        T s0_ = T(0), s1_ = T(0);                                                                  \
        _Pragma("unroll") for (int w_ = 0; w_ < NW; ++w_) {                                        \
            s0_ += red[flip * 2 * NW + w_]; s1_ += red[flip * 2 * NW + NW + w_];                   \
        }                                                                                          \
        flip ^= 1;                                                                                 \
        out0 = s0_; out1 = s1_;                                                                    \
    }

// z = Ginv * (vector in p_s) -> element `tid` in `out`.  Same row split as the G p term.  One barrier inside.
#define CYMF_PRECOND(out)                                                                          \
    {                                                                                              \
        if (lane_on) {                                                                             \
            T acc[VW];                                                                             \
            _Pragma("unroll") for (int e = 0; e < VW; ++e) acc[e] = T(0);                          \
            const T *g = Giw;                                                                      \
            _Pragma("unroll 4") for (int j = j0; j < j1; ++j, g += ld) {                           \
                T gv[VW];                                                                          \
                ldg_vec<VW>(g, gv);                                                                \
                const T pj = p_s[j];                                                               \
                _Pragma("unroll") for (int e = 0; e < VW; ++e) acc[e] += gv[e] * pj;               \
            }                                                                                      \
            st_vec<VW>(part + warp * CG_VEC + kq, acc);                                            \
        }                                                                                          \
        __syncthreads();                                                                           \
        T o_ = T(0);                                                                               \
        if (own) { _Pragma("unroll") for (int w_ = 0; w_ < NW; ++w_) o_ += part[w_ * CG_VEC + tid]; } \
        out = o_;                                                                                  \
    }

#define CYMF_STREAM_LOAD(v0, v1, v2, v3)                                                            \
                if (lane_on) {                                                                     \
                    ldg_vec<VW>(Yq + yrow_off(__ldg(idx + i), ld), v0);                            \
                    ldg_vec<VW>(Yq + yrow_off(__ldg(idx + i + NW), ld), v1);                       \
                    ldg_vec<VW>(Yq + yrow_off(__ldg(idx + i + 2 * NW), ld), v2);                   \
                    ldg_vec<VW>(Yq + yrow_off(__ldg(idx + i + 3 * NW), ld), v3);                   \
                }
#define CYMF_STREAM_GROUP(v0, v1, v2, v3)                                                           \
                {                                                                                  \
                    T d0 = dot_slice<T, VW>(v0, ps), d1 = dot_slice<T, VW>(v1, ps),                \
                      d2 = dot_slice<T, VW>(v2, ps), d3 = dot_slice<T, VW>(v3, ps);                \
                    warp_allsum4(d0, d1, d2, d3, lane);                                            \
                    axpy4<T, VW>(acc, d0, v0, d1, v1, d2, v2, d3, v3);                             \
                }

// A v for the vector in p_s -> element `tid` of the result in `out` (0 for tid >= ld).  One barrier inside.
// The warp's items are warp, warp+NW, ...; four at a time share one reduction butterfly.
#define CYMF_APPLY(out)                                                                            \
    {                                                                                              \
        T ps[VW], acc[VW];                                                                         \
        _Pragma("unroll") for (int e = 0; e < VW; ++e) { ps[e] = T(0); acc[e] = T(0); }            \
        if (lane_on) ld_vec<VW>(p_s + kq, ps);                                                     \
        int i = warp;                                                                              \
        for (; i + 3 * NW < ns; i += 4 * NW) {                  /* four staged items */           \
            T y0[VW], y1[VW], y2[VW], y3[VW];                                                      \
            _Pragma("unroll") for (int e = 0; e < VW; ++e) { y0[e] = y1[e] = y2[e] = y3[e] = T(0); } \
            if (lane_on) {                                                                         \
                const T *s = Ys + i * ld + kq;                                                     \
                ld_vec<VW>(s, y0); ld_vec<VW>(s + NW * ld, y1);                                    \
                ld_vec<VW>(s + 2 * NW * ld, y2); ld_vec<VW>(s + 3 * NW * ld, y3);                  \
            }                                                                                      \
            T d0 = dot_slice<T, VW>(y0, ps), d1 = dot_slice<T, VW>(y1, ps), d2 = dot_slice<T, VW>(y2, ps),      \
              d3 = dot_slice<T, VW>(y3, ps);                                                       \
            warp_allsum4(d0, d1, d2, d3, lane);                                                    \
            axpy4<T, VW>(acc, d0, y0, d1, y1, d2, y2, d3, y3);                                     \
        }                                                                                          \
        if (i < ns) { CYMF_MIXED_GROUP(); i += 4 * NW; }        /* the group that straddles the staging limit */ \
        if (i + 3 * NW < nnz) {                /* streamed groups: two register sets take turns, the other one's gathers in flight */ \
            T a0[VW], a1[VW], a2[VW], a3[VW], b0[VW], b1[VW], b2[VW], b3[VW];                      \
            _Pragma("unroll") for (int e = 0; e < VW; ++e) {                                       \
                a0[e] = a1[e] = a2[e] = a3[e] = T(0); b0[e] = b1[e] = b2[e] = b3[e] = T(0);        \
            }                                                                                      \
            CYMF_STREAM_LOAD(a0, a1, a2, a3)                                                       \
            for (;;) {                                                                             \
                i += 4 * NW;                                                                       \
                const bool more_b = i + 3 * NW < nnz;                                              \
                if (more_b) { CYMF_STREAM_LOAD(b0, b1, b2, b3) }                                   \
                CYMF_STREAM_GROUP(a0, a1, a2, a3)                                                  \
                if (!more_b) break;                                                                \
                i += 4 * NW;                                                                       \
                const bool more_a = i + 3 * NW < nnz;                                              \
                if (more_a) { CYMF_STREAM_LOAD(a0, a1, a2, a3) }                                   \
                CYMF_STREAM_GROUP(b0, b1, b2, b3)                                                  \
                if (!more_a) break;                                                                \
            }                                                                                      \
        }                                                                                          \
        for (; i < nnz; i += 4 * NW) { CYMF_MIXED_GROUP(); }    /* tail */                         \
        _Pragma("unroll") for (int e = 0; e < VW; ++e) acc[e] *= wm1;                              \
        CYMF_APPLY_TAIL(out)                                                                       \
    }

// one group of four items that may be staged, streamed or past the end of the row
#define CYMF_MIXED_GROUP()                                                                         \
        {                                                                                          \
            T y0[VW], y1[VW], y2[VW], y3[VW];                                                      \
            _Pragma("unroll") for (int e = 0; e < VW; ++e) { y0[e] = y1[e] = y2[e] = y3[e] = T(0); } \
            if (lane_on) {                                                                         \
                const int i1 = i + NW, i2 = i + 2 * NW, i3 = i + 3 * NW;                           \
                if (i < ns) ld_vec<VW>(Ys + i * ld + kq, y0);                                      \
                else ldg_vec<VW>(Yq + yrow_off(__ldg(idx + i), ld), y0);                            \
                if (i1 < ns) ld_vec<VW>(Ys + i1 * ld + kq, y1);                                    \
                else if (i1 < nnz) ldg_vec<VW>(Yq + yrow_off(__ldg(idx + i1), ld), y1);             \
                if (i2 < ns) ld_vec<VW>(Ys + i2 * ld + kq, y2);                                    \
                else if (i2 < nnz) ldg_vec<VW>(Yq + yrow_off(__ldg(idx + i2), ld), y2);             \
                if (i3 < ns) ld_vec<VW>(Ys + i3 * ld + kq, y3);                                    \
                else if (i3 < nnz) ldg_vec<VW>(Yq + yrow_off(__ldg(idx + i3), ld), y3);             \
            }                                                                                      \
            T d0 = dot_slice<T, VW>(y0, ps), d1 = dot_slice<T, VW>(y1, ps), d2 = dot_slice<T, VW>(y2, ps),      \
              d3 = dot_slice<T, VW>(y3, ps);                                                       \
            warp_allsum4(d0, d1, d2, d3, lane);                                                    \
            axpy4<T, VW>(acc, d0, y0, d1, y1, d2, y2, d3, y3);                                     \
        }

// + rows [j0, j1) of G p (untransformed solvers), publish the warp's partial, barrier, combine
#define CYMF_APPLY_TAIL(out)                                                                       \
        if (lane_on) {                                                                             \
            if (dense) {                                                                           \
                const T *g = Gw;                                                                   \
                _Pragma("unroll 4") for (int j = j0; j < j1; ++j, g += ld) {  /* + rows [j0, j1) of G p */ \
                    T gv[VW];                                                                      \
                    ldg_vec<VW>(g, gv);                                                            \
                    const T pj = p_s[j];                                                           \
                    _Pragma("unroll") for (int e = 0; e < VW; ++e) acc[e] += gv[e] * pj;           \
                }                                                                                  \
            }                                                                                      \
            st_vec<VW>(part + warp * CG_VEC + kq, acc);                                            \
        }                                                                                          \
        __syncthreads();                                                                           \
        T o_ = T(0);                                                                               \
        if (own) {                                                                                 \
            _Pragma("unroll") for (int w_ = 0; w_ < NW; ++w_) o_ += part[w_ * CG_VEC + tid];       \
            if (!dense) o_ += p_s[tid];                /* transformed space: G = I */             \
        }                                                                                          \
        out = o_;

        // Fused first pass over the row: stage its item vectors, accumulate b = w * sum y_i (wmf.pyx:163) AND the
        // (w-1) sum y_i (y_i . x0) term of A x0 for the warm start x0, so that r0 = b - A x0 costs no second pass
        // (one of the ~7 passes a row takes).  Groups of four items, the next group's gathers already in flight.
        if (own) p_s[tid] = xr[tid];                                            // warm start
        __syncthreads();
        T acc[VW];
        {
            T bacc[VW], ps[VW], c0[VW], c1[VW], c2[VW], c3[VW];
#pragma unroll
            for (int e = 0; e < VW; ++e) { bacc[e] = acc[e] = ps[e] = T(0); c0[e] = c1[e] = c2[e] = c3[e] = T(0); }
            if (lane_on) ld_vec<VW>(p_s + kq, ps);
#define CYMF_GATHER4(at, v0, v1, v2, v3)                                                           \
            if ((at) + 3 * NW >= nnz) {                      /* partial (or empty) group: clear the set first */ \
                _Pragma("unroll") for (int e = 0; e < VW; ++e) v0[e] = v1[e] = v2[e] = v3[e] = T(0); \
            }                                                                                      \
            if (lane_on) {                                                                         \
                if ((at) < nnz) ldg_vec<VW>(Yq + yrow_off(__ldg(idx + (at)), ld), v0);              \
                if ((at) + NW < nnz) ldg_vec<VW>(Yq + yrow_off(__ldg(idx + (at) + NW), ld), v1);    \
                if ((at) + 2 * NW < nnz) ldg_vec<VW>(Yq + yrow_off(__ldg(idx + (at) + 2 * NW), ld), v2); \
                if ((at) + 3 * NW < nnz) ldg_vec<VW>(Yq + yrow_off(__ldg(idx + (at) + 3 * NW), ld), v3); \
            }
#define CYMF_FIRST_GROUP(at, v0, v1, v2, v3)                                                       \
            {                                                                                      \
                if (lane_on) {                                                                     \
                    if ((at) < ns) st_vec<VW>(Ys + (at) * ld + kq, v0);                            \
                    if ((at) + NW < ns) st_vec<VW>(Ys + ((at) + NW) * ld + kq, v1);                \
                    if ((at) + 2 * NW < ns) st_vec<VW>(Ys + ((at) + 2 * NW) * ld + kq, v2);        \
                    if ((at) + 3 * NW < ns) st_vec<VW>(Ys + ((at) + 3 * NW) * ld + kq, v3);        \
                }                                                                                  \
                _Pragma("unroll") for (int e = 0; e < VW; ++e) bacc[e] += (v0[e] + v1[e]) + (v2[e] + v3[e]); \
                T d0 = dot_slice<T, VW>(v0, ps), d1 = dot_slice<T, VW>(v1, ps), d2 = dot_slice<T, VW>(v2, ps), \
                  d3 = dot_slice<T, VW>(v3, ps);                                                   \
                warp_allsum4(d0, d1, d2, d3, lane);                                                \
                axpy4<T, VW>(acc, d0, v0, d1, v1, d2, v2, d3, v3);                                 \
            }
            T n0[VW], n1[VW], n2[VW], n3[VW];
#pragma unroll
            for (int e = 0; e < VW; ++e) n0[e] = n1[e] = n2[e] = n3[e] = T(0);
            CYMF_GATHER4(warp, c0, c1, c2, c3)
            for (int i = warp; i < nnz; i += 8 * NW) {           // two register sets take turns (no copies)
                CYMF_GATHER4(i + 4 * NW, n0, n1, n2, n3)
                CYMF_FIRST_GROUP(i, c0, c1, c2, c3)
                if (i + 4 * NW >= nnz) break;
                CYMF_GATHER4(i + 8 * NW, c0, c1, c2, c3)
                CYMF_FIRST_GROUP(i + 4 * NW, n0, n1, n2, n3)
            }
#undef CYMF_FIRST_GROUP
#undef CYMF_GATHER4
            if (lane_on) st_vec<VW>(part + warp * CG_VEC + kq, bacc);
        }
        __syncthreads();
        T b = T(0), x = T(0);
        if (own) {
#pragma unroll
            for (int w = 0; w < NW; ++w) b += part[w * CG_VEC + tid];
            b *= a.weight;
            x = p_s[tid];
        }
        T bb;
        CYMF_BLOCK_SUM(bb, b * b);                                              // barrier: Ys is settled, part has been read
        unsigned iters = 0;
        bool stalled = false;
        if (bb > T(0)) {
            T Ax;
#pragma unroll
            for (int e = 0; e < VW; ++e) acc[e] *= wm1;
            CYMF_APPLY_TAIL(Ax)                                                 // + G x0 (or + x0), combine the warps
            T res = b - Ax;                                                     // r0 = b - A x0
            T rs, rz;
            T z = res;
            if (pre) {                                                          // z = G^-1 r (preconditioner)
                __syncthreads();
                if (own) p_s[tid] = res;
                __syncthreads();
                CYMF_PRECOND(z);
                CYMF_BLOCK_SUM2(rs, rz, res * res, res * z);
            } else {
                CYMF_BLOCK_SUM(rs, res * res);
                rz = rs;
            }
            T p = z;
            while (rs > a.tol2 * bb) {
                if ((int)iters >= a.max_iter) { stalled = true; break; }
                if (own) p_s[tid] = p;
                __syncthreads();
                T Ap;
                CYMF_APPLY(Ap);
                T pAp;
                CYMF_BLOCK_SUM(pAp, p * Ap);
                if (!(pAp > T(0))) { stalled = true; break; }
                const T alpha = rz / pAp;
                x += alpha * p;
                res -= alpha * Ap;
                T rs_new, rz_new;
                if (pre) {
                    if (own) p_s[tid] = res;                                    // all reads of p_s are behind a barrier
                    __syncthreads();
                    CYMF_PRECOND(z);
                    CYMF_BLOCK_SUM2(rs_new, rz_new, res * res, res * z);
                } else {
                    CYMF_BLOCK_SUM(rs_new, res * res);
                    rz_new = rs_new;
                    z = res;
                }
                p = z + (rz_new / rz) * p;
                rs = rs_new;
                rz = rz_new;
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
#undef CYMF_APPLY
#undef CYMF_STREAM_LOAD
#undef CYMF_STREAM_GROUP
#undef CYMF_BLOCK_SUM
#undef CYMF_BLOCK_SUM2
#undef CYMF_MIXED_GROUP
#undef CYMF_APPLY_TAIL
#undef CYMF_PRECOND
    }
}

// ---- heavy rows: direct solve ---------------------------------------------------------------------------------
// A row with hundreds of thousands of entries keeps ONE CTA of the CG kernel busy for tens of milliseconds (its
// item vectors are streamed ~7 times at a single SM's bandwidth) while the rest of the GPU has drained -- the tail of
// every half sweep at C5, and the same at any GPU count.  For those rows the reference's own formulation is used
// instead: the K x K matrix A = G + (w-1) sum y y^T (wmf.pyx:161-166) is accumulated ONCE, by many CTAs, on the
// tensor cores (tc_gram_partial_kernel<GATHER>, 512 entries per slab), and solved directly (wmf.pyx:168) here:
// one CTA per row sums the slab partials in slab order (deterministic), factors A = L D L^T in f64 in shared memory
// and substitutes.  G == NULL means the identity (transformed coordinates).
template <typename T>
__global__ void __launch_bounds__(256) als_heavy_solve_kernel(const double *__restrict__ partial,
                                                              const double *__restrict__ bsum,
                                                              const int32_t *__restrict__ first_slab,
                                                              const int32_t *__restrict__ order, T *__restrict__ X,
                                                              const double *__restrict__ G, double add_diag, int K, int ld,
                                                              double weight) {
    extern __shared__ double sm[];
    const int S = K + 1;                                    // padded stride: column walks spread over the banks
    double *A = sm, *b = sm + (size_t)K * S;
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int h = blockIdx.x, s0 = first_slab[h], s1 = first_slab[h + 1];
    for (int t = tid; t < K * K; t += blockDim.x) {
        double acc = 0.0;
        for (int s = s0; s < s1; ++s) acc += partial[(size_t)s * K * K + t];
        const int i = t / K, j = t - i * K;
        const double g = G ? G[t] + (i == j ? add_diag : 0.0) : (i == j ? 1.0 : 0.0);
        A[i * S + j] = g + (weight - 1.0) * acc;
    }
    for (int t = tid; t < K; t += blockDim.x) {
        double acc = 0.0;
        for (int s = s0; s < s1; ++s) acc += bsum[(size_t)s * ld + t];
        b[t] = weight * acc;
    }
    __syncthreads();
    for (int k = 0; k < K; ++k) {                           // L D L^T, column scaling deferred: one barrier per step
        const double inv_d = 1.0 / A[k * S + k];
        for (int i = k + 1 + ty; i < K; i += 8) {
            const double lik = A[i * S + k] * inv_d;
            for (int j = k + 1 + tx; j <= i; j += 32) A[i * S + j] -= lik * A[j * S + k];
        }
        __syncthreads();
    }
    for (int k = 0; k < K; ++k) {                           // forward: unit lower factor L_ik = A[i][k] / d_k
        const double zk = b[k] / A[k * S + k];
        __syncthreads();
        for (int i = k + 1 + tid; i < K; i += blockDim.x) b[i] -= A[i * S + k] * zk;
        __syncthreads();
    }
    for (int k = tid; k < K; k += blockDim.x) b[k] /= A[k * S + k];       // D^-1
    __syncthreads();
    for (int k = K - 1; k > 0; --k) {                       // backward: L^T x = y
        const double xk = b[k];
        for (int i = tid; i < k; i += blockDim.x) b[i] -= A[k * S + i] / A[i * S + i] * xk;
        __syncthreads();
    }
    T *xr = X + (size_t)order[h] * ld;
    for (int t = tid; t < ld; t += blockDim.x) xr[t] = t < K ? (T)b[t] : T(0);
}

template <typename T> static int gram_impl(const T *Y, int64_t n, int K, int ld, double wd, int add_wd, double *partial,
                                           int64_t partial_capacity, double *out64, T *outT, cudaStream_t st) {
    if constexpr (sizeof(T) == 4) {
        if (tc_enabled() && tc_shape_ok(CYMF_F32, ld)) {          // tensor cores: 3xTF32 tcgen05 partials per 256-row slab
            const int tslabs = (int)tc_gram_slabs(n);
            if ((int64_t)tslabs * K * K > partial_capacity) { set_error("gram: workspace too small"); return CYMF_EINVAL; }
            CYMF_TRY(tc_gram_partial((const float *)Y, n, K, ld, partial, st));
            gram_finish_kernel<T><<<(ld * ld + 255) / 256, 256, 0, st>>>(partial, tslabs, K, ld, add_wd ? wd : 0.0, out64, outT);
            CYMF_LAUNCHED();
            return 0;
        }
    }
    const int slabs = (int)((n + GRAM_SLAB - 1) / GRAM_SLAB);
    if ((int64_t)slabs * K * K > partial_capacity) { set_error("gram: workspace too small"); return CYMF_EINVAL; }
    if (slabs > 0) {
        if (K <= 32) gram_partial_kernel<T, 2><<<slabs, 256, 0, st>>>(Y, n, K, ld, partial);
        else if (K <= 64) gram_partial_kernel<T, 4><<<slabs, 256, 0, st>>>(Y, n, K, ld, partial);
        else if (K <= 128) gram_partial_kernel<T, 8><<<slabs, 256, 0, st>>>(Y, n, K, ld, partial);
        else {
            const int nb = (K + 63) / 64;
            gram_wide_kernel<T><<<dim3((unsigned)slabs, (unsigned)(nb * nb)), 256, 0, st>>>(Y, n, K, ld, partial);
        }
        CYMF_LAUNCHED();
    }
    gram_finish_kernel<T><<<(ld * ld + 255) / 256, 256, 0, st>>>(partial, slabs, K, ld, add_wd ? wd : 0.0, out64, outT);
    CYMF_LAUNCHED();
    return 0;
}


template <typename T, int VW, int NW> static int launch_cg(AlsArgs<T> a, int32_t stage_rows, cudaStream_t st) {
    const size_t fixed = sizeof(T) * ((size_t)(1 + NW) * CG_VEC + 4 * NW);
    if (stage_rows <= 0) {   // auto: 4 KB of staged item vectors per warp -- measured (tools/als_tune.py): resident warps
                             // matter more than staging; 16/32/64 KB per CTA keeps >= 36 warps per SM
        stage_rows = (int32_t)((size_t)NW * 4 * 1024 / (sizeof(T) * a.ld));
        if (stage_rows < 8) stage_rows = 8;
    }
    a.stage_rows = stage_rows;
    const size_t smem = fixed + sizeof(T) * (size_t)stage_rows * a.ld;
    if (smem > 220 * 1024) { set_error("als: stage_rows=%d does not fit shared memory", stage_rows); return CYMF_EINVAL; }
    auto kern = als_cg_kernel<T, VW, NW>;
    if (smem > 48 * 1024) CYMF_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CYMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 32 * NW, smem));
    if (per_sm < 1) per_sm = 1;
    if (const char *cap = getenv(NW == 4 ? "CYMF_ALS_CTAS4" : (NW == 8 ? "CYMF_ALS_CTAS8" : "CYMF_ALS_CTAS16"))) {
        const int c = atoi(cap);                       // tuning hook: resident CTAs per SM for this row class
        if (c >= 1 && c < per_sm) per_sm = c;
    }
    int64_t blocks = (int64_t)sm_count() * per_sm;
    if (blocks > a.n_solve) blocks = a.n_solve;
    if (blocks < 1) blocks = 1;
    kern<<<(unsigned)blocks, 32 * NW, smem, st>>>(a);
    CYMF_LAUNCHED();
    return 0;
}

template <typename T, int NW> static int cg_by_width(const AlsArgs<T> &a, int32_t stage_rows, cudaStream_t st) {
    if (a.ld <= 32) return launch_cg<T, 1, NW>(a, stage_rows, st);
    if (a.ld <= 64) return launch_cg<T, 2, NW>(a, stage_rows, st);
    if (a.ld <= 128) return launch_cg<T, 4, NW>(a, stage_rows, st);
    if constexpr (NW >= 8) return launch_cg<T, 8, NW>(a, stage_rows, st);
    set_error("als: ld > 128 needs at least 8 warps per row");
    return CYMF_EINVAL;
}

template <typename T> static int cg_impl(const AlsArgs<T> &a, int32_t warps_per_row, int32_t stage_rows, cudaStream_t st) {
    CYMF_CUDA(cudaMemsetAsync(a.queue, 0, sizeof(int32_t), st));
    if (warps_per_row >= 16) return cg_by_width<T, 16>(a, stage_rows, st);
    if (warps_per_row >= 8 || a.ld > 128) return cg_by_width<T, 8>(a, stage_rows, st);
    return cg_by_width<T, 4>(a, stage_rows, st);
}

}  // namespace cymf

using namespace cymf;

extern "C" int64_t cymf_gram_workspace_doubles(int64_t n, int32_t K) {
    const int64_t a = (n + GRAM_SLAB - 1) / GRAM_SLAB, b = tc_gram_slabs(n);     // FFMA / tcgen05 slab counts
    return (a > b ? a : b) * (int64_t)K * K;
}

extern "C" int cymf_gram_dev(const void *Y, int dtype, int64_t n, int32_t K, int32_t ld, double weight_decay,
                             int add_weight_decay, double *workspace, int64_t workspace_doubles,
                             double *out_f64, void *out_native, void *stream) {
    CYMF_REQUIRE(Y && workspace && (out_f64 || out_native), "null pointer");
    CYMF_REQUIRE(n >= 0 && K > 0 && K <= 256 && ld >= K && ld % 4 == 0 && ld <= 256,
                 "bad shape (WMF supports num_components <= 256, ld a multiple of 4)");
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
    CYMF_REQUIRE(in_f64 && out_native && K > 0 && K <= 256 && ld >= K && ld <= 256, "bad argument");
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

// Cross-rank sum of Gram partials: out64 = sum_r parts[r] (+ wd on the diagonal), ranks in index order on every rank
// (so all ranks hold bit-identical G).  parts[r] may be NVLink peer memory (symmetric buffers): this is the Gram
// all-reduce of the sharded half sweep done with plain peer loads instead of a latency-bound NCCL call.
struct GramParts { const double *p[8]; int n; };
__global__ void gram_sum_kernel(const GramParts parts, int KK, int K, double wd, double *__restrict__ out64) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= KK) return;
    double s = 0.0;
    for (int r = 0; r < parts.n; ++r) s += parts.p[r][t];
    if (t / K == t % K) s += wd;
    out64[t] = s;
}

extern "C" int cymf_gram_sum_dev(const double *const *parts, int32_t n_parts, int32_t K, double weight_decay,
                                 double *out_f64, void *stream) {
    CYMF_REQUIRE(parts && out_f64 && n_parts >= 1 && n_parts <= 8 && K > 0 && K <= 256, "bad argument");
    GramParts gp{};
    gp.n = n_parts;
    for (int r = 0; r < n_parts; ++r) { CYMF_REQUIRE(parts[r] != nullptr, "null partial"); gp.p[r] = parts[r]; }
    gram_sum_kernel<<<(K * K + 255) / 256, 256, 0, (cudaStream_t)stream>>>(gp, K * K, K, weight_decay, out_f64);
    CYMF_LAUNCHED();
    return 0;
}

// In-place Gauss-Jordan inversion of a symmetric positive definite K x K matrix (f64, one CTA, matrix in shared
// memory; no pivoting is needed for SPD input).  Output [ld, ld] of T, zero padded.
template <typename T>
__global__ void __launch_bounds__(1024) spd_inverse_kernel(const double *__restrict__ A, int K, int ld, double add_diag,
                                                           T *__restrict__ out) {
    extern __shared__ double sm[];
    double *M = sm, *rowk = sm + K * K, *colk = rowk + K;
    for (int t = threadIdx.x; t < K * K; t += blockDim.x) M[t] = A[t] + ((t / K == t % K) ? add_diag : 0.0);
    __syncthreads();
    for (int k = 0; k < K; ++k) {
        const double piv = M[k * K + k];
        for (int t = threadIdx.x; t < K; t += blockDim.x) {
            rowk[t] = (t == k ? 1.0 : M[k * K + t]) / piv;
            colk[t] = M[t * K + k];
        }
        __syncthreads();
        for (int t = threadIdx.x; t < K * K; t += blockDim.x) {
            const int i = t / K, j = t - i * K;
            if (i == k) M[t] = rowk[j];
            else M[t] = (j == k ? 0.0 : M[t]) - colk[i] * rowk[j];
        }
        __syncthreads();
    }
    for (int t = threadIdx.x; t < ld * ld; t += blockDim.x) {
        const int i = t / ld, j = t - i * ld;
        out[t] = (i < K && j < K) ? (T)(0.5 * (M[i * K + j] + M[j * K + i])) : T(0);
    }
}

// Cholesky G = L L^T (f64, one CTA, shared memory) and L^-1, written out as the three [ld, ld] (zero padded)
// right-hand matrices of the change of variables  y~ = L^-1 y,  x~ = L^T x  under which the row systems become
// (I + (w-1) sum y~ y~^T) x~ = w sum y~ :   By = L^-T (Y~ = Y By),   Bfwd = L (X~ = X Bfwd),   Bbwd = L^-1 (X = X~ Bbwd).
template <typename T>
__global__ void __launch_bounds__(1024) chol_transforms_kernel(const double *__restrict__ A, int K, int ld, double add_diag,
                                                               T *__restrict__ By, T *__restrict__ Bfwd, T *__restrict__ Bbwd,
                                                               int *__restrict__ info) {
    extern __shared__ double sm[];
    const int S = K + 1;                     // padded row stride: column accesses L[j][k] spread over the banks
    double *L = sm;                          // [K][S]
    double *part = sm + K * S;               // [2][8][128] partial sums of the forward substitution
    double *rdiag = part + 2048;             // [K] 1 / L[k][k];  [K .. K+1]: reciprocal of the next pivot (two slots)
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    for (int t = tid; t < K * K; t += blockDim.x) {
        const int i = t / K, j = t - i * K;
        L[i * S + j] = A[t] + (i == j ? add_diag : 0.0);
    }
    for (int t = tid; t < ld * ld; t += blockDim.x) { By[t] = T(0); Bfwd[t] = T(0); Bbwd[t] = T(0); }
    __syncthreads();
    // Right-looking factorisation with the column scaling deferred (L D L^T form): step k only subtracts
    // l_ik l_jk / d_k from the trailing block, so one barrier per step suffices; columns are divided by sqrt(d_k)
    // in a single pass at the end.  The reciprocal of pivot k+1 is computed by the one thread that finishes
    // L[k+1][k+1] during step k, so the ~200-cycle f64 division is off the other 1023 threads' critical path.
    if (tid == 0) {
        const double d0 = L[0];
        if (!(d0 > 0.0) && info) *info = 1;
        rdiag[K] = 1.0 / (d0 > 0.0 ? d0 : 1.0);
    }
    __syncthreads();
    for (int k = 0; k < K; ++k) {
        const double inv_d = rdiag[K + (k & 1)];
        for (int i = k + 1 + ty; i < K; i += 32) {                 // 32 x 32 thread tile over the trailing block
            const double lik = L[i * S + k] * inv_d;
            for (int j = k + 1 + tx; j <= i; j += 32) L[i * S + j] -= lik * L[j * S + k];
        }
        if (tid == 0 && k + 1 < K) {                               // thread 0 has just finished L[k+1][k+1]
            const double dn = L[(k + 1) * S + k + 1];
            if (!(dn > 0.0) && info) *info = k + 2;
            rdiag[K + ((k + 1) & 1)] = 1.0 / (dn > 0.0 ? dn : 1.0);
        }
        __syncthreads();
    }
    for (int t = tid; t < K * K; t += blockDim.x) {
        const int i = t / K, j = t - i * K;
        if (i > j) { const double dj = L[j * S + j]; L[i * S + j] = L[i * S + j] / sqrt(dj > 0.0 ? dj : 1.0); }
    }
    __syncthreads();
    for (int k = tid; k < K; k += blockDim.x) {
        const double dk = L[k * S + k], lk = sqrt(dk > 0.0 ? dk : 1.0);
        L[k * S + k] = lk;
        rdiag[k] = 1.0 / lk;
    }
    __syncthreads();
    for (int t = tid; t < K * K; t += blockDim.x) {
        const int i = t / K, j = t - i * K;
        if (i >= j) Bfwd[i * ld + j] = (T)L[i * S + j];            // L     (lower triangular)
    }
    // L^-1 by forward substitution, column by column: column c is x with L x = e_c,
    //     x_i = ((i == c) - sum_{c <= j < i} L[i][j] x_j) / L[i][i] ,   i = c .. K-1 .
    // Columns are independent, so no block barrier is needed (round 1 ran all columns in lock step with one
    // __syncthreads per row: ~110 us of the kernel's 250 at K = 128): an octet of lanes owns a column -- lane s
    // keeps x_j for j = s (mod 8) in registers and adds its share of the sum, three shuffle steps combine the eight
    // shares -- and a warp carries four columns, 32 warps all 128.
    {
        const int c = 4 * ty + (tx >> 3), s8 = tx & 7;
        double mine[16];                                           // x_j for j = 8 q + s8
#pragma unroll
        for (int q = 0; q < 16; ++q) mine[q] = 0.0;
        for (int i = 4 * ty; i < K; ++i) {                         // (warp-uniform bounds: the shuffles need every lane)
            double acc = 0.0;
            if (c < K && c <= i) {
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const int j = 8 * q + s8;
                    if (j >= c && j < i) acc += L[i * S + j] * mine[q];
                }
            }
            acc += __shfl_xor_sync(0xffffffffu, acc, 1);
            acc += __shfl_xor_sync(0xffffffffu, acc, 2);
            acc += __shfl_xor_sync(0xffffffffu, acc, 4);
            if (c < K && c <= i) {
                const double v = ((i == c ? 1.0 : 0.0) - acc) * rdiag[i];
                if (s8 == (i & 7)) {
#pragma unroll
                    for (int q = 0; q < 16; ++q)
                        if (q == (i >> 3)) mine[q] = v;
                    Bbwd[i * ld + c] = (T)v;                       // L^-1  (lower triangular)
                    By[c * ld + i] = (T)v;                         // L^-T  (upper triangular): By[k][c'] = Linv[c'][k]
                }
            }
        }
    }
}

// out[r, :] = in[r, :] * B   (B is [ld, ld] in shared memory; 64-row tiles; in place allowed).
// The result tile is written to every destination in `outs`: with one destination this is a plain skinny GEMM;
// with the peers' replicas of the factor matrix as destinations (NVLink peer memory) the kernel is the GEMM
// AND the all-gather of the solved block -- each rank's rows land in all replicas straight from the epilogue.
constexpr int MAX_DESTS = 8;
template <typename T> struct MultiOut { T *p[MAX_DESTS]; int n; };

template <typename T>
__global__ void __launch_bounds__(256) rows_times_matrix_kernel(const T *in /* may alias a destination */, const MultiOut<T> outs,
                                                                const T *__restrict__ B, int64_t rows, int ld) {
    extern __shared__ __align__(16) unsigned char smem_raw2[];
    T *Bs = reinterpret_cast<T *>(smem_raw2);          // [ld][ld]
    T *tile = Bs + ld * ld;                            // [64][ld + 1]
    const int ts = ld + 1;
    for (int t = threadIdx.x; t < ld * ld; t += 256) Bs[t] = B[t];
    const int cgroups = ld / 4;                        // thread -> 4 consecutive output columns, several rows
    const int cq = (threadIdx.x % cgroups) * 4, rsub = threadIdx.x / cgroups, rstep = 256 / cgroups;
    for (int64_t base = (int64_t)blockIdx.x * 64; base < rows; base += (int64_t)gridDim.x * 64) {
        __syncthreads();
        for (int t = threadIdx.x; t < 64 * ld; t += 256) {
            const int rr = t / ld, c = t - rr * ld;
            tile[rr * ts + c] = base + rr < rows ? in[(size_t)(base + rr) * ld + c] : T(0);
        }
        __syncthreads();
        if (threadIdx.x < cgroups * rstep)
            for (int r0 = rsub; r0 < 64; r0 += 4 * rstep) {        // 4 rows x 4 columns per thread and pass
                T acc[4][4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[q][e] = T(0);
                const T *x = tile + r0 * ts;
#pragma unroll 2
                for (int k = 0; k < ld; ++k) {
                    const T *b = Bs + k * ld + cq;
                    const T b0 = b[0], b1 = b[1], b2 = b[2], b3 = b[3];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int rr = r0 + q * rstep;
                        const T xv = rr < 64 ? x[q * rstep * ts + k] : T(0);
                        acc[q][0] += xv * b0; acc[q][1] += xv * b1; acc[q][2] += xv * b2; acc[q][3] += xv * b3;
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int rr = r0 + q * rstep;
                    if (rr < 64 && base + rr < rows) {
                        const size_t off = (size_t)(base + rr) * ld + cq;
                        for (int d = 0; d < outs.n; ++d) st_vec<4>(outs.p[d] + off, acc[q]);
                    }
                }
            }
    }
}

// Blocked form of the same computation (G = L L^T, outputs L, L^-1, L^-T), 256 threads: thread (bi, bj) owns the 8 x 8
// block of rows 8 bi .., columns 8 bj .. IN REGISTERS.  The unblocked kernel above spends ~1.5 us on each of its 128
// barrier-separated column steps whatever the work; here a step is a block column:
//   (a) the diagonal thread factorises its 8 x 8 block and inverts the triangular factor (the only sqrt / divide chain),
//       hidden behind the other warps' trailing update of the previous step;
//   (b) the panel threads form L_ik = A_ik L_kk^-T; (c) the trailing threads subtract L_ik L_jk^T -- two barriers per step.
// L^-1 follows right-looking as well: once block row t of X = L^-1 is complete, every block below adds L_it X_tj to its
// accumulator, and X_ij = -L_ii^-1 acc_ij when its own row comes up.  Rows / columns past K are padded with the identity.
template <typename T>
__global__ void __launch_bounds__(256) chol_transforms_blocked_kernel(const double *__restrict__ A, int K, int ld, double add_diag,
                                                                      T *__restrict__ By, T *__restrict__ Bfwd,
                                                                      T *__restrict__ Bbwd, int *__restrict__ info) {
    extern __shared__ double sm[];
    double *Lb = sm;                         // [16][16][64] blocks of L (lower block triangle used)
    double *Wall = Lb + 16 * 16 * 64;        // [16][64] inverses of the diagonal blocks of L
    double *XR = Wall + 16 * 64;             // [16][64] the block row of L^-1 that has just been completed
    const int tid = threadIdx.x, bi = tid >> 4, bj = tid & 15;
    const int NB = (K + 7) >> 3;
    const bool lower = bi >= bj && bi < NB;
    for (int t = tid; t < ld * ld; t += 256) { By[t] = T(0); Bfwd[t] = T(0); Bbwd[t] = T(0); }
    double a[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            const int i = 8 * bi + r, j = 8 * bj + c;
            a[r][c] = (lower && i < K && j < K) ? A[(size_t)i * K + j] + (i == j ? add_diag : 0.0) : (i == j ? 1.0 : 0.0);
        }
    // (a): factor this thread's (diagonal) block in place -> lower triangle of a = L_kk, w = L_kk^-1, both published
    auto factor_diag = [&](int kb) {
        double w[8][8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            double d = a[c][c];
#pragma unroll
            for (int t = 0; t < c; ++t) d -= a[c][t] * a[c][t];
            if (!(d > 0.0)) { if (info) *info = 8 * kb + c + 1; d = 1.0; }
            const double l = sqrt(d), rl = 1.0 / l;
            a[c][c] = l;
#pragma unroll
            for (int r = c + 1; r < 8; ++r) {
                double v = a[r][c];
#pragma unroll
                for (int t = 0; t < c; ++t) v -= a[r][t] * a[c][t];
                a[r][c] = v * rl;
            }
#pragma unroll
            for (int r = 0; r < c; ++r) a[r][c] = 0.0;
            // row c of the inverse: w[c][c] = 1 / l, w[c][j] = -(sum_{t=j}^{c-1} l[c][t] w[t][j]) / l for j < c
            w[c][c] = rl;
#pragma unroll
            for (int j = 0; j < c; ++j) {
                double v = 0.0;
#pragma unroll
                for (int t = j; t < c; ++t) v += a[c][t] * w[t][j];
                w[c][j] = -v * rl;
            }
#pragma unroll
            for (int j = c + 1; j < 8; ++j) w[c][j] = 0.0;
        }
        double *lw = Lb + (kb * 16 + kb) * 64, *ww = Wall + kb * 64;
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) { lw[8 * r + c] = a[r][c]; ww[8 * r + c] = w[r][c]; }
    };
    if (bi == 0 && bj == 0) factor_diag(0);
    __syncthreads();
    for (int kb = 0; kb < NB; ++kb) {
        if (bj == kb && bi > kb && bi < NB) {                      // (b) panel: L_ik = A_ik W_k^T
            const double *ww = Wall + kb * 64;
            double l[8][8];
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                double wc[8];
#pragma unroll
                for (int t = 0; t <= c; ++t) wc[t] = ww[8 * c + t];
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    double v = 0.0;
#pragma unroll
                    for (int t = 0; t <= c; ++t) v += a[r][t] * wc[t];
                    l[r][c] = v;
                }
            }
            double *lw = Lb + (bi * 16 + kb) * 64;
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) lw[8 * r + c] = l[r][c];
        }
        __syncthreads();
        if (bj > kb && lower) {                                    // (c) trailing update: A_ij -= L_ik L_jk^T
            const double *li = Lb + (bi * 16 + kb) * 64, *lj = Lb + (bj * 16 + kb) * 64;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                double x[8], y[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) { x[r] = li[8 * r + t]; y[r] = lj[8 * r + t]; }
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = 0; c < 8; ++c) a[r][c] -= x[r] * y[c];
            }
            if (bi == kb + 1 && bj == kb + 1) factor_diag(kb + 1);   // next step's (a), under the other warps' updates
        }
        __syncthreads();
    }
    // L (lower triangular) -> Bfwd
    if (lower) {
        const double *lw = Lb + (bi * 16 + bj) * 64;
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const int i = 8 * bi + r, j = 8 * bj + c;
                if (i < K && j <= i) Bfwd[i * ld + j] = (T)lw[8 * r + c];
            }
    }
    // X = L^-1, block row by block row; a[][] now accumulates sum_t L_it X_tj for this thread's block
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) a[r][c] = 0.0;
    for (int t = 0; t < NB; ++t) {
        if (bi == t && bj <= t) {
            const double *ww = Wall + t * 64;
            double x[8][8];
            if (bj == t) {
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = 0; c < 8; ++c) x[r][c] = ww[8 * r + c];
            } else {
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    double wr[8];
#pragma unroll
                    for (int q = 0; q <= r; ++q) wr[q] = ww[8 * r + q];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        double v = 0.0;
#pragma unroll
                        for (int q = 0; q <= r; ++q) v += wr[q] * a[q][c];
                        x[r][c] = -v;
                    }
                }
            }
            double *xw = XR + bj * 64;
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    xw[8 * r + c] = x[r][c];
                    const int i = 8 * t + r, j = 8 * bj + c;
                    if (i < K && j <= i) {
                        Bbwd[i * ld + j] = (T)x[r][c];              // L^-1  (lower triangular)
                        By[j * ld + i] = (T)x[r][c];                // L^-T  (upper triangular)
                    }
                }
        }
        __syncthreads();
        if (bi > t && bi < NB && bj <= t) {                         // acc_ij += L_it X_tj
            const double *li = Lb + (bi * 16 + t) * 64, *xj = XR + bj * 64;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                double x[8], y[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) { x[r] = li[8 * r + q]; y[r] = xj[8 * q + r]; }
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int c = 0; c < 8; ++c) a[r][c] += x[r] * y[c];
            }
        }
        __syncthreads();
    }
}

extern "C" int cymf_chol_transforms_dev(const double *A, int32_t K, int32_t ld, double add_diag, int dtype,
                                        void *By, void *Bfwd, void *Bbwd, int32_t *info, void *stream) {
    CYMF_REQUIRE(A && By && Bfwd && Bbwd && K > 0 && K <= 128 && ld >= K && ld <= 128, "bad argument");
    // (A blocked variant -- 32-column panels, in-warp diagonal blocks, a dozen barriers -- was measured SLOWER: 289 us vs
    // 235 us at K = 128, 129 vs 84 at K = 64: the serial f64 sqrt / divide chain of the pivots dominates either way.)
    cudaStream_t st = (cudaStream_t)stream;
    // ... until the blocks were kept in REGISTERS (chol_transforms_blocked_kernel); CYMF_CHOL_BLOCKED=0 selects the unblocked kernel
    const char *env = getenv("CYMF_CHOL_BLOCKED");
    if (!(env && env[0] == '0')) {
        const size_t bsm = sizeof(double) * (16 * 16 * 64 + 2 * 16 * 64);
        if (dtype == CYMF_F32) {
            CYMF_CUDA(cudaFuncSetAttribute(chol_transforms_blocked_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsm));
            chol_transforms_blocked_kernel<float><<<1, 256, bsm, st>>>(A, K, ld, add_diag, (float *)By, (float *)Bfwd, (float *)Bbwd, info);
        } else {
            CYMF_CUDA(cudaFuncSetAttribute(chol_transforms_blocked_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsm));
            chol_transforms_blocked_kernel<double><<<1, 256, bsm, st>>>(A, K, ld, add_diag, (double *)By, (double *)Bfwd, (double *)Bbwd, info);
        }
        CYMF_LAUNCHED();
        return 0;
    }
    const size_t smem = sizeof(double) * ((size_t)K * (K + 1) + 2048 + K + 2);
    if (dtype == CYMF_F32) {
        if (smem > 48 * 1024)
            CYMF_CUDA(cudaFuncSetAttribute(chol_transforms_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        chol_transforms_kernel<float><<<1, 1024, smem, st>>>(A, K, ld, add_diag, (float *)By, (float *)Bfwd, (float *)Bbwd, info);
    } else {
        if (smem > 48 * 1024)
            CYMF_CUDA(cudaFuncSetAttribute(chol_transforms_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        chol_transforms_kernel<double><<<1, 1024, smem, st>>>(A, K, ld, add_diag, (double *)By, (double *)Bfwd, (double *)Bbwd, info);
    }
    CYMF_LAUNCHED();
    return 0;
}

extern "C" int cymf_rows_times_matrix_multi_dev(const void *in, void *const *outs, int32_t n_outs, const void *B,
                                                int dtype, int64_t rows, int32_t ld, void *stream);

extern "C" int cymf_rows_times_matrix_dev(const void *in, void *out, const void *B, int dtype, int64_t rows, int32_t ld,
                                          void *stream) {
    void *outs[1] = {out};
    return cymf_rows_times_matrix_multi_dev(in, outs, 1, B, dtype, rows, ld, stream);
}

extern "C" int cymf_rows_times_matrix_multi_dev(const void *in, void *const *outs, int32_t n_outs, const void *B,
                                                int dtype, int64_t rows, int32_t ld, void *stream) {
    CYMF_REQUIRE(in && outs && B && rows >= 0 && ld > 0 && ld % 4 == 0 && ld <= 128, "bad argument");
    CYMF_REQUIRE(n_outs >= 1 && n_outs <= MAX_DESTS, "1..8 destinations");
    for (int d = 0; d < n_outs; ++d) CYMF_REQUIRE(outs[d] != nullptr, "null destination");
    if (rows == 0) return 0;
    if (tc_enabled() && tc_shape_ok(dtype, ld))                  // tensor cores (tcgen05, 3xTF32), tc_gemm.cu
        return tc_rows_times_matrix((const float *)in, (float *const *)outs, n_outs, (const float *)B, rows, ld,
                                    (cudaStream_t)stream);
    const size_t es = dtype == CYMF_F32 ? 4 : 8;
    const size_t smem = es * ((size_t)ld * ld + (size_t)64 * (ld + 1));
    cudaStream_t st = (cudaStream_t)stream;
    int64_t blocks = (rows + 63) / 64;
    const int64_t cap = (int64_t)sm_count() * 2;
    if (blocks > cap) blocks = cap;
    if (dtype == CYMF_F32) {
        if (smem > 48 * 1024)
            CYMF_CUDA(cudaFuncSetAttribute(rows_times_matrix_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MultiOut<float> mo{};
        mo.n = n_outs;
        for (int d = 0; d < n_outs; ++d) mo.p[d] = (float *)outs[d];
        rows_times_matrix_kernel<float><<<(unsigned)blocks, 256, smem, st>>>((const float *)in, mo, (const float *)B, rows, ld);
    } else {
        if (smem > 48 * 1024)
            CYMF_CUDA(cudaFuncSetAttribute(rows_times_matrix_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MultiOut<double> mo{};
        mo.n = n_outs;
        for (int d = 0; d < n_outs; ++d) mo.p[d] = (double *)outs[d];
        rows_times_matrix_kernel<double><<<(unsigned)blocks, 256, smem, st>>>((const double *)in, mo, (const double *)B, rows, ld);
    }
    CYMF_LAUNCHED();
    return 0;
}

extern "C" int cymf_spd_inverse_dev(const double *A, int32_t K, int32_t ld, double add_diag, int dtype,
                                    void *out_native, void *stream) {
    CYMF_REQUIRE(A && out_native && K > 0 && K <= 128 && ld >= K && ld <= 128, "bad argument");
    const size_t smem = sizeof(double) * ((size_t)K * K + 2 * K);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CYMF_F32) {
        if (smem > 48 * 1024)
            CYMF_CUDA(cudaFuncSetAttribute(spd_inverse_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        spd_inverse_kernel<float><<<1, 1024, smem, st>>>(A, K, ld, add_diag, (float *)out_native);
    } else {
        if (smem > 48 * 1024)
            CYMF_CUDA(cudaFuncSetAttribute(spd_inverse_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        spd_inverse_kernel<double><<<1, 1024, smem, st>>>(A, K, ld, add_diag, (double *)out_native);
    }
    CYMF_LAUNCHED();
    return 0;
}

extern "C" int cymf_als_cg_dev(const int64_t *indptr, const int32_t *indices, const int32_t *order, int32_t n_solve,
                               void *X, const void *Y, const void *G, const void *Ginv, int dtype, int32_t K, int32_t ld,
                               double weight, double cg_tol, int32_t cg_max_iter, int32_t warps_per_row,
                               int32_t stage_rows, int32_t *queue, unsigned long long *stats, void *stream) {
    CYMF_REQUIRE(indptr && indices && order && X && Y && queue, "null pointer");
    CYMF_REQUIRE(G || !Ginv, "Ginv without G");
    CYMF_REQUIRE(K > 0 && K <= 256 && ld >= K && ld % 4 == 0 && ld <= 256,
                 "bad shape (WMF supports num_components <= 256, ld a multiple of 4)");
    CYMF_REQUIRE(cg_tol > 0 && cg_max_iter > 0, "bad CG parameters");
    if (n_solve <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CYMF_F32) {
        AlsArgs<float> a{indptr, indices, order, n_solve, (float *)X, (const float *)Y, (const float *)G,
                         (const float *)Ginv, ld, 0, cg_max_iter, (float)weight, (float)(cg_tol * cg_tol), queue, stats};
        return cg_impl<float>(a, warps_per_row, stage_rows, st);
    }
    if (dtype == CYMF_F64) {
        AlsArgs<double> a{indptr, indices, order, n_solve, (double *)X, (const double *)Y, (const double *)G,
                          (const double *)Ginv, ld, 0, cg_max_iter, weight, cg_tol * cg_tol, queue, stats};
        return cg_impl<double>(a, warps_per_row, stage_rows, st);
    }
    set_error("als: unknown dtype %d", dtype);
    return CYMF_EINVAL;
}

extern "C" int64_t cymf_als_heavy_workspace_doubles(int64_t n_slabs, int32_t K, int32_t ld) {
    return n_slabs * ((int64_t)K * K + ld);
}

// Direct solve of the heavy rows order[0 .. n_heavy) of the block (see als_heavy_solve_kernel).  first_slab: device
// int32[n_heavy + 1], first_slab[h + 1] - first_slab[h] = ceil(len_h / 512), first_slab[n_heavy] = n_slabs.
// G64: dense [K, K] doubles (+ add_diag on the diagonal) or NULL for the identity (Y in transformed coordinates).
// f32 factors with ld in {32, 64, 96, 128} only (tensor-core path); otherwise CYMF_EUNSUPPORTED.
extern "C" int cymf_als_heavy_rows_dev(const int64_t *indptr, const int32_t *indices, const int32_t *order,
                                       int32_t n_heavy, const int32_t *first_slab, int32_t n_slabs, void *X, const void *Y,
                                       const double *G64, double add_diag, int dtype, int32_t K, int32_t ld, double weight,
                                       double *workspace, int64_t workspace_doubles, void *stream) {
    CYMF_REQUIRE(indptr && indices && order && first_slab && X && Y && workspace, "null pointer");
    CYMF_REQUIRE(n_heavy >= 0 && n_slabs >= n_heavy && K > 0 && ld >= K, "bad shape");
    if (!(tc_shape_ok(dtype, ld) && tc_enabled())) {
        set_error("als heavy rows: needs f32 factors with ld in {32, 64, 96, 128} and tcgen05 enabled");
        return CYMF_EUNSUPPORTED;
    }
    if (n_heavy == 0) return 0;
    CYMF_REQUIRE(workspace_doubles >= cymf_als_heavy_workspace_doubles(n_slabs, K, ld), "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    double *partial = workspace, *bsum = workspace + (size_t)n_slabs * K * K;
    CYMF_TRY(tc_gram_gather((const float *)Y, indptr, indices, order, first_slab, n_heavy, n_slabs, K, ld, partial, bsum, st));
    const size_t smem = sizeof(double) * ((size_t)K * (K + 1) + K);
    CYMF_CUDA(cudaFuncSetAttribute(als_heavy_solve_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    als_heavy_solve_kernel<float><<<(unsigned)n_heavy, 256, smem, st>>>(partial, bsum, first_slab, order, (float *)X, G64,
                                                                      add_diag, K, ld, weight);
    CYMF_LAUNCHED();
    return 0;
}

// Rows sorted by decreasing length split into three classes by how many item vectors fit the staging area of a
// 4-, 8- or 16-warp CTA (8 KB per warp): returns the number of rows for 16 and for 8 warps (the rest take 4).
extern "C" int cymf_als_row_classes(const int64_t *sorted_lengths_desc, int64_t n, int dtype, int32_t ld,
                                    int64_t *n_wide16, int64_t *n_wide8) {
    CYMF_REQUIRE(sorted_lengths_desc && n_wide16 && n_wide8 && n >= 0 && ld > 0, "bad argument");
    // Measured on B200 (tools/als_tune.py, ml-20m shape, K=128, transformed solver): rows averaging 130 entries
    // take 11.4 ms with 4 warps per row, 13.5 ms with 8, 23 ms with 16; rows averaging 673 entries (heavy tail up
    // to 31 k) take 12.4 / 8.9 / 11.4 ms.  Per-iteration barriers against per-warp work set the crossover.
    (void)dtype; (void)ld;
    const int64_t wide8 = 384, wide16 = 4096;
    int64_t a = 0, b = 0;
    while (a < n && sorted_lengths_desc[a] > wide16) ++a;
    b = a;
    while (b < n && sorted_lengths_desc[b] > wide8) ++b;
    *n_wide16 = a;
    *n_wide8 = b - a;
    return 0;
}

extern "C" int cymf_als_half_host(const int32_t *indptr, const int32_t *indices, double *X, const double *Y,
                                  int64_t rows, int64_t n, int32_t K, double weight_decay, double weight,
                                  int dtype, double cg_tol, int32_t cg_max_iter, int64_t *cg_iterations_out) {
    CYMF_REQUIRE(indptr && indices && X && Y, "null pointer");
    CYMF_REQUIRE(rows > 0 && n > 0 && K > 0 && K <= 256, "bad shape (WMF supports num_components <= 256)");
    CYMF_REQUIRE(dtype == CYMF_F32 || dtype == CYMF_F64, "unknown dtype");
    int ndev = 0;
    CYMF_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev < 1) { set_error("no CUDA device: cymf_b200 has no CPU fallback"); return CYMF_EUNSUPPORTED; }
    if (cg_tol <= 0) cg_tol = dtype == CYMF_F32 ? 1e-6 : 1e-10;
    if (cg_max_iter <= 0) cg_max_iter = 2 * K;
    const size_t es = dtype == CYMF_F32 ? 4 : 8;
    const bool wide = K > 128;                                         // plain CG on the untransformed systems
    const bool tc_rows = dtype == CYMF_F32 && tc_enabled() && !wide;   // one-pass tensor-core row solver (als_tc.cu)
    const int32_t ld = tc_rows ? (K + 31) / 32 * 32 : (K + 3) / 4 * 4;
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
    // heaviest rows with 16 warps per row, medium with 8, the rest with 4
    std::vector<int64_t> len((size_t)rows);
    for (int64_t t = 0; t < rows; ++t) len[(size_t)t] = indptr[order[(size_t)t] + 1] - indptr[order[(size_t)t]];
    int64_t n16 = 0, n8 = 0;
    CYMF_TRY(cymf_als_row_classes(len.data(), rows, dtype, ld, &n16, &n8));
    const int64_t start[3] = {0, n16, n16 + n8}, count[3] = {n16, n8, rows - n16 - n8};
    const int32_t width[3] = {16, 8, 4};
    if (wide) {
        // num_components > 128: CG on (G + (w-1) sum y y^T) x = w sum y itself, G p product inside every iteration
        for (int c = 0; c < 3; ++c)
            CYMF_TRY(cymf_als_cg_dev(d_ip, d_ix, d_order + start[c], (int32_t)count[c], dX, dY, dG, nullptr, dtype, K, ld,
                                     weight, cg_tol, cg_max_iter, width[c], 0, d_queue, d_stats, st));
    } else {
        // change of variables y~ = L^-1 y, x~ = L^T x (G = L L^T): the CG iteration then has no dense K x K product
        void *dBy, *dBf, *dBb, *dYt;
        CYMF_TRY(mem.get((char **)&dBy, (size_t)ld * ld * es));
        CYMF_TRY(mem.get((char **)&dBf, (size_t)ld * ld * es));
        CYMF_TRY(mem.get((char **)&dBb, (size_t)ld * ld * es));
        CYMF_TRY(mem.get((char **)&dYt, (size_t)n * ld * es));
        CYMF_TRY(cymf_chol_transforms_dev(g64, K, ld, 0.0, dtype, dBy, dBf, dBb, nullptr, st));
        CYMF_TRY(cymf_rows_times_matrix_dev(dY, dYt, dBy, dtype, n, ld, st));
        CYMF_TRY(cymf_rows_times_matrix_dev(dX, dX, dBf, dtype, rows, ld, st));
        if (tc_rows) {
            CYMF_TRY(cymf_als_rows_tc_dev(d_ip, d_ix, d_order, (int32_t)rows, dX, dYt, dtype, K, ld, weight, cg_tol,
                                          cg_max_iter, d_queue, d_stats, st));
        } else {
            for (int c = 0; c < 3; ++c)
                CYMF_TRY(cymf_als_cg_dev(d_ip, d_ix, d_order + start[c], (int32_t)count[c], dX, dYt, nullptr, nullptr, dtype,
                                         K, ld, weight, cg_tol, cg_max_iter, width[c], 0, d_queue, d_stats, st));
        }
        CYMF_TRY(cymf_rows_times_matrix_dev(dX, dX, dBb, dtype, rows, ld, st));
    }
    CYMF_TRY(cymf_unpack_rows_dev(dX, stage, dtype, rows, K, ld, st));
    CYMF_CUDA(cudaMemcpyAsync(X, stage, (size_t)rows * K * 8, cudaMemcpyDeviceToHost, st));
    unsigned long long stats[2] = {0, 0};
    CYMF_CUDA(cudaMemcpyAsync(stats, d_stats, 16, cudaMemcpyDeviceToHost, st));
    CYMF_CUDA(cudaStreamSynchronize(st));
    if (cg_iterations_out) *cg_iterations_out = (int64_t)stats[0];
    return 0;
}
