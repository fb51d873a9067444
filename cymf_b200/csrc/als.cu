// als.cu -- WMF alternating-least-squares half sweep (replaces WMF._als, cymf/wmf.pyx:136-174, and the dgesv
// call behind cymf/linalg.pyx:144-163).
//
// Reference, per row r with item set S_r:   (G + (w-1) sum_{i in S_r} y_i y_i^T) x_r = w sum_{i in S_r} y_i,
// G = Y^T Y + wd I, solved by dense LU after materialising the K x K matrix (O(|S_r| K^2) scalar work).
// Here the matrix is never formed:
//   gram_partial_kernel / gram_finish_kernel : G (wmf.pyx:142-143).  Row slabs -> per-slab K x K partials (f32 or
//       f64 products, f64 cross-slab sum) -> deterministic reduction; in multi-GPU runs the partial of a rank's own
//       row block is all-reduced before wd*I is added.
//   als_cg_kernel : one 128-thread CTA per row, conjugate gradient on A p = G p + (w-1) sum_i y_i (y_i . p).
//       The row's item vectors are staged once in shared memory (as many as fit the staging budget; the rest is
//       re-read through L2 each iteration), warps split the items, the residual recurrence runs until
//       |r| <= tol |b| (warm start from the current x_r).  Rows come from a heaviest-first work queue.
//   Algorithmic bytes per half sweep: N (K s + 4) + rows (K s + 8) + n K s.
#include <math.h>

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

// sum the slab partials in slab order (deterministic), optionally add wd on the diagonal, write f64 and/or T
template <typename T>
__global__ void gram_finish_kernel(const double *__restrict__ partial, int slabs, int K, double wd,
                                   double *__restrict__ out64, T *__restrict__ outT) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= K * K) return;
    double s = 0.0;
    for (int b = 0; b < slabs; ++b) s += partial[(size_t)b * K * K + t];
    if (t / K == t % K) s += wd;
    if (out64) out64[t] = s;
    if (outT) outT[t] = (T)s;
}

// ---- CG row solver ---------------------------------------------------------------------------------------------
template <typename T> struct AlsArgs {
    const int64_t *indptr;
    const int32_t *indices;
    const int32_t *order;       // rows to solve, heaviest first
    int32_t n_solve;
    T *X;                       // [rows, ldx]  solved in place (warm start = current content)
    const T *Y;                 // [n, ldy]
    const T *G;                 // [K, K]  Y^T Y + wd I
    int32_t K, ldx, ldy, stage_rows, max_iter;
    T weight, tol2;
    int32_t *queue;             // work-queue head (zeroed before launch)
    unsigned long long *stats;  // [0] CG iterations summed over rows, [1] rows that hit max_iter (may be NULL)
};

template <typename T> __device__ __forceinline__ T warp_allsum(T v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

template <typename T> __device__ __forceinline__ T block_sum128(T v, T *red) {   // 4 warps; red[4]
    v = warp_allsum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    const T s = (red[0] + red[1]) + (red[2] + red[3]);
    __syncthreads();
    return s;
}

template <typename T, int M>          // M = ceil(K / 32): elements of a K-vector per lane
__global__ void __launch_bounds__(128) als_cg_kernel(const AlsArgs<T> a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int KP = 32 * M;
    T *p_s = reinterpret_cast<T *>(smem_raw);     // [KP]      search direction, readable by all warps
    T *part = p_s + KP;                           // [4][KP]   per-warp partial sums
    T *red = part + 4 * KP;                       // [4]
    T *Ys = red + 4;                              // [stage_rows][K]
    __shared__ int row_slot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = a.K;
    const bool own = tid < K;                     // thread k owns element k of x, r, p (K <= 128)

    // A v for the vector currently in p_s; returns element `tid`.  Ends with every thread past the last barrier.
    auto apply = [&](int64_t lo, int nnz, int ns) -> T {
        T ps[M], acc[M];
#pragma unroll
        for (int m = 0; m < M; ++m) { ps[m] = p_s[lane + 32 * m]; acc[m] = T(0); }
        for (int i = warp; i < nnz; i += 4) {
            T yv[M];
            if (i < ns) {
#pragma unroll
                for (int m = 0; m < M; ++m) { const int k = lane + 32 * m; yv[m] = k < K ? Ys[i * K + k] : T(0); }
            } else {
                const T *y = a.Y + (size_t)__ldg(a.indices + lo + i) * a.ldy;
#pragma unroll
                for (int m = 0; m < M; ++m) { const int k = lane + 32 * m; yv[m] = k < K ? __ldg(y + k) : T(0); }
            }
            T t = T(0);
#pragma unroll
            for (int m = 0; m < M; ++m) t += yv[m] * ps[m];
            t = warp_allsum(t);
#pragma unroll
            for (int m = 0; m < M; ++m) acc[m] += t * yv[m];
        }
#pragma unroll
        for (int m = 0; m < M; ++m) part[warp * KP + lane + 32 * m] = acc[m];
        T gp = T(0);
        if (own) {
            T g0 = T(0), g1 = T(0);
            int j = 0;
            for (; j + 1 < K; j += 2) {
                g0 += __ldg(a.G + (size_t)j * K + tid) * p_s[j];
                g1 += __ldg(a.G + (size_t)(j + 1) * K + tid) * p_s[j + 1];
            }
            if (j < K) g0 += __ldg(a.G + (size_t)j * K + tid) * p_s[j];
            gp = g0 + g1;
        }
        __syncthreads();
        T out = T(0);
        if (own) out = gp + (a.weight - T(1)) * ((part[tid] + part[KP + tid]) + (part[2 * KP + tid] + part[3 * KP + tid]));
        __syncthreads();
        return out;
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
        T *xr = a.X + (size_t)r * a.ldx;
        if (nnz == 0) {                                                        // wmf.pyx:154-156
            if (tid < a.ldx) xr[tid] = T(0);
            continue;
        }
        const int ns = nnz < a.stage_rows ? nnz : a.stage_rows;

        // stage the row's item vectors and accumulate b = w * sum y_i (wmf.pyx:163)
        T bacc[M];
#pragma unroll
        for (int m = 0; m < M; ++m) bacc[m] = T(0);
        for (int i = warp; i < nnz; i += 4) {
            const T *y = a.Y + (size_t)__ldg(a.indices + lo + i) * a.ldy;
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const int k = lane + 32 * m;
                const T v = k < K ? __ldg(y + k) : T(0);
                bacc[m] += v;
                if (i < ns && k < K) Ys[i * K + k] = v;
            }
        }
#pragma unroll
        for (int m = 0; m < M; ++m) part[warp * KP + lane + 32 * m] = bacc[m];
        __syncthreads();
        T b = T(0), x = T(0);
        if (own) {
            b = a.weight * ((part[tid] + part[KP + tid]) + (part[2 * KP + tid] + part[3 * KP + tid]));
            x = xr[tid];                                                        // warm start
        }
        if (tid < KP) p_s[tid] = own ? x : T(0);
        __syncthreads();
        const T bb = block_sum128(b * b, red);
        unsigned iters = 0;
        bool stalled = false;
        if (bb > T(0)) {
            T res = b - apply(lo, nnz, ns);                                     // r0 = b - A x0
            T p = res;
            T rs = block_sum128(res * res, red);
            while (rs > a.tol2 * bb) {
                if ((int)iters >= a.max_iter) { stalled = true; break; }
                if (tid < KP) p_s[tid] = own ? p : T(0);
                __syncthreads();
                const T Ap = apply(lo, nnz, ns);
                const T pAp = block_sum128(p * Ap, red);
                if (!(pAp > T(0))) { stalled = true; break; }
                const T alpha = rs / pAp;
                x += alpha * p;
                res -= alpha * Ap;
                const T rs_new = block_sum128(res * res, red);
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
    gram_finish_kernel<T><<<(K * K + 255) / 256, 256, 0, st>>>(partial, slabs, K, add_wd ? wd : 0.0, out64, outT);
    CYMF_LAUNCHED();
    return 0;
}

template <typename T, int M> static int launch_cg(const AlsArgs<T> &a, cudaStream_t st) {
    constexpr int KP = 32 * M;
    const size_t smem = sizeof(T) * ((size_t)5 * KP + 4 + (size_t)a.stage_rows * a.K);
    auto kern = als_cg_kernel<T, M>;
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
    if (a.K <= 32) return launch_cg<T, 1>(a, st);
    if (a.K <= 64) return launch_cg<T, 2>(a, st);
    if (a.K <= 96) return launch_cg<T, 3>(a, st);
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
    CYMF_REQUIRE(n >= 0 && K > 0 && K <= 128 && ld >= K, "bad shape (WMF supports num_components <= 128)");
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

// out_native[t] = (T)(in_f64[t] + wd on the diagonal): finishes a Gram matrix that was all-reduced in f64
extern "C" int cymf_gram_finalize_dev(const double *in_f64, int dtype, int32_t K, double weight_decay, void *out_native,
                                      void *stream) {
    CYMF_REQUIRE(in_f64 && out_native && K > 0 && K <= 128, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CYMF_F32)
        gram_finish_kernel<float><<<(K * K + 255) / 256, 256, 0, st>>>(in_f64, 1, K, weight_decay, nullptr, (float *)out_native);
    else
        gram_finish_kernel<double><<<(K * K + 255) / 256, 256, 0, st>>>(in_f64, 1, K, weight_decay, nullptr, (double *)out_native);
    CYMF_LAUNCHED();
    return 0;
}

extern "C" int cymf_als_cg_dev(const int64_t *indptr, const int32_t *indices, const int32_t *order, int32_t n_solve,
                               void *X, const void *Y, const void *G, int dtype, int32_t K, int32_t ldx, int32_t ldy,
                               double weight, double cg_tol, int32_t cg_max_iter, int32_t stage_rows,
                               int32_t *queue, unsigned long long *stats, void *stream) {
    CYMF_REQUIRE(indptr && indices && order && X && Y && G && queue, "null pointer");
    CYMF_REQUIRE(K > 0 && K <= 128 && ldx >= K && ldy >= K && ldx <= 128, "bad shape (WMF supports num_components <= 128)");
    CYMF_REQUIRE(cg_tol > 0 && cg_max_iter > 0, "bad CG parameters");
    if (n_solve <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t es = dtype == CYMF_F32 ? 4 : 8;
    if (stage_rows <= 0) {                       // auto: ~32 KB of staged item vectors per CTA
        stage_rows = (int32_t)(32 * 1024 / (es * K));
        if (stage_rows < 8) stage_rows = 8;
    }
    if (dtype == CYMF_F32) {
        AlsArgs<float> a{indptr, indices, order, n_solve, (float *)X, (const float *)Y, (const float *)G, K, ldx, ldy,
                         stage_rows, cg_max_iter, (float)weight, (float)(cg_tol * cg_tol), queue, stats};
        return cg_impl<float>(a, st);
    }
    if (dtype == CYMF_F64) {
        AlsArgs<double> a{indptr, indices, order, n_solve, (double *)X, (const double *)Y, (const double *)G, K, ldx, ldy,
                          stage_rows, cg_max_iter, weight, cg_tol * cg_tol, queue, stats};
        return cg_impl<double>(a, st);
    }
    set_error("als: unknown dtype %d", dtype);
    return CYMF_EINVAL;
}

#include <algorithm>
#include <vector>

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
    CYMF_TRY(mem.get((char **)&dG, (size_t)K * K * es));
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
    CYMF_TRY(cymf_als_cg_dev(d_ip, d_ix, d_order, (int32_t)rows, dX, dY, dG, dtype, K, ld, ld, weight, cg_tol,
                             cg_max_iter, 0, d_queue, d_stats, st));
    CYMF_TRY(cymf_unpack_rows_dev(dX, stage, dtype, rows, K, ld, st));
    CYMF_CUDA(cudaMemcpyAsync(X, stage, (size_t)rows * K * 8, cudaMemcpyDeviceToHost, st));
    unsigned long long stats[2] = {0, 0};
    CYMF_CUDA(cudaMemcpyAsync(stats, d_stats, 16, cudaMemcpyDeviceToHost, st));
    CYMF_CUDA(cudaStreamSynchronize(st));
    if (cg_iterations_out) *cg_iterations_out = (int64_t)stats[0];
    return 0;
}
