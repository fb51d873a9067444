// common.cuh -- shared device helpers for the cymf_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/cymf_b200.h"

namespace cymf {

// ---- status plumbing (capi.cu) ---------------------------------------------------------------------------
void set_error(const char *fmt, ...);
int cuda_status(cudaError_t e, const char *what, const char *file, int line);
extern std::atomic<int64_t> g_launches;
int sm_count();

#define CYMF_CUDA(expr)                                                           \
    do {                                                                          \
        cudaError_t e_ = (expr);                                                  \
        if (e_ != cudaSuccess) return ::cymf::cuda_status(e_, #expr, __FILE__, __LINE__); \
    } while (0)
#define CYMF_REQUIRE(cond, msg)                                  \
    do {                                                         \
        if (!(cond)) {                                           \
            ::cymf::set_error("%s: %s", __func__, msg);          \
            return CYMF_EINVAL;                                  \
        }                                                        \
    } while (0)
#define CYMF_LAUNCHED()                                            \
    do {                                                           \
        ::cymf::g_launches.fetch_add(1, std::memory_order_relaxed); \
        CYMF_CUDA(cudaGetLastError());                             \
    } while (0)

// tcgen05 (tensor core) implementations of the GEMM-shaped pieces, tc_gemm.cu; f32, ld in {32, 64, 96, 128}
bool tc_shape_ok(int dtype, int ld);
int tc_rows_times_matrix(const float *in, float *const *outs, int n_outs, const float *B, int64_t rows, int ld,
                         cudaStream_t st);
int tc_rows_times_matrix_tma(const float *in, float *const *outs, int n_outs, const float *B, int64_t rows, int ld,
                             cudaStream_t st);      // tc_gemm_tma.cu; CYMF_EUNSUPPORTED if the tensor map cannot be encoded
int64_t tc_gram_slabs(int64_t n);
int tc_gram_partial(const float *Y, int64_t n, int K, int ld, double *partial, cudaStream_t st);
int tc_gram_gather(const float *Y, const int64_t *indptr, const int32_t *indices, const int32_t *order,
                   const int32_t *first_slab, int n_heavy, int n_slabs, int K, int ld, double *partial, double *bsum,
                   cudaStream_t st);
int tc_als_rows(const int64_t *indptr, const int32_t *indices, const int32_t *order, int32_t n_solve, float *X,
                const float *Y, int ld, float weight, float tol2, int32_t max_iter, int32_t *queue,
                unsigned long long *stats, cudaStream_t st);      // als_tc.cu
int tc_als_rows6(const int64_t *indptr, const int32_t *indices, const int32_t *order, int32_t n_solve, float *X,
                 const float *Y, int ld, float weight, float tol2, int32_t max_iter, int32_t *queue,
                 unsigned long long *stats, cudaStream_t st);     // als_tc6.cu
bool tc_enabled();      // false when the environment sets CYMF_NO_TCGEN05=1 (A/B comparisons in tests and tools)

#define CYMF_TRY(expr)             \
    do {                           \
        int rc_ = (expr);          \
        if (rc_) return rc_;       \
    } while (0)

// Device allocations of one `*_host` call: everything handed out is freed when the call returns, on every path.
struct DeviceArena {
    void *blocks[64];
    int count = 0;
    ~DeviceArena() {
        for (int t = 0; t < count; ++t) cudaFree(blocks[t]);
    }
    template <typename P> int get(P **out, size_t bytes) {
        void *p = nullptr;
        *out = nullptr;
        if (count >= 64) { set_error("DeviceArena: too many blocks"); return CYMF_ENOMEM; }
        cudaError_t e = cudaMalloc(&p, bytes ? bytes : 1);
        if (e != cudaSuccess) return cuda_status(e, "cudaMalloc", __FILE__, __LINE__);
        blocks[count++] = p;
        *out = (P *)p;
        return 0;
    }
};

// grid for a flat grid-stride kernel of n items (256 threads per block, at most 8 blocks per SM)
static inline unsigned flat_grid(int64_t n) {
    int64_t b = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

// ---- Philox4x32-10 (counter-based; one call per triplet, no state to carry) ---------------------------------
__host__ __device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = mulhi32(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = mulhi32(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}

// Unbiased map of up to four 32-bit words onto [0, n): multiply-shift with rejection of the biased low
// words (same rule libstdc++ applies to the reference's mt19937 stream); the 4th word is taken as is.
__host__ __device__ __forceinline__ uint32_t bounded_from_words(uint4 r, uint32_t n) {
    uint64_t prod = (uint64_t)r.x * n;
    uint32_t low = (uint32_t)prod;
    if (low < n) {
        const uint32_t thr = (0u - n) % n;
        if (low < thr) {
            prod = (uint64_t)r.y * n; low = (uint32_t)prod;
            if (low < thr) {
                prod = (uint64_t)r.z * n; low = (uint32_t)prod;
                if (low < thr) prod = (uint64_t)r.w * n;
            }
        }
    }
    return (uint32_t)(prod >> 32);
}

__host__ __device__ __forceinline__ uint32_t philox_negative(uint64_t seed, uint32_t epoch, uint64_t l, uint32_t n) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)l, (uint32_t)(l >> 32), epoch, 0x42505231u /* "BPR1" */),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    return bounded_from_words(r, n);
}

#ifdef __CUDACC__
// ---- 4-element row slots: 16 B (f32) or 32 B (f64), L2-coherent loads, write-through stores -------------------
template <typename T> struct Slot { T v[4]; };

__device__ __forceinline__ Slot<float> load_slot(const float *p) {
    const float4 t = __ldcg(reinterpret_cast<const float4 *>(p));
    Slot<float> s; s.v[0] = t.x; s.v[1] = t.y; s.v[2] = t.z; s.v[3] = t.w; return s;
}
__device__ __forceinline__ Slot<double> load_slot(const double *p) {
    const double2 a = __ldcg(reinterpret_cast<const double2 *>(p));
    const double2 b = __ldcg(reinterpret_cast<const double2 *>(p) + 1);
    Slot<double> s; s.v[0] = a.x; s.v[1] = a.y; s.v[2] = b.x; s.v[3] = b.y; return s;
}
__device__ __forceinline__ void store_slot(float *p, const Slot<float> &s) {
    __stcg(reinterpret_cast<float4 *>(p), make_float4(s.v[0], s.v[1], s.v[2], s.v[3]));
}
__device__ __forceinline__ void store_slot(double *p, const Slot<double> &s) {
    __stcg(reinterpret_cast<double2 *>(p), make_double2(s.v[0], s.v[1]));
    __stcg(reinterpret_cast<double2 *>(p) + 1, make_double2(s.v[2], s.v[3]));
}
// additive scatter: one 128-bit reduction per lane (f32) / four 64-bit reductions (f64), no return value
__device__ __forceinline__ void red_add_slot(float *p, const Slot<float> &s) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(s.v[0]), "f"(s.v[1]), "f"(s.v[2]),
                 "f"(s.v[3])
                 : "memory");
}
__device__ __forceinline__ void red_add_slot(double *p, const Slot<double> &s) {
#pragma unroll
    for (int e = 0; e < 4; ++e)
        asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p + e), "d"(s.v[e]) : "memory");
}
template <typename T> __device__ __forceinline__ Slot<T> zero_slot() {
    Slot<T> s; s.v[0] = s.v[1] = s.v[2] = s.v[3] = T(0); return s;
}

// ---- lane groups: LPT consecutive lanes of a warp cooperate on one sample ------------------------------------
template <int LPT> __device__ __forceinline__ unsigned group_mask(int lane) {
    if constexpr (LPT == 32) return 0xffffffffu;
    else return ((1u << LPT) - 1u) << ((lane / LPT) * LPT);
}

template <int LPT, typename T> __device__ __forceinline__ T group_sum(T x, unsigned gmask) {
#pragma unroll
    for (int off = LPT / 2; off > 0; off >>= 1) x += __shfl_xor_sync(gmask, x, off);
    return x;
}

// Is `key` present in the sorted run indices[lo, hi)?  (LPT+1)-ary search: every step each lane probes one
// splitter, one ballot narrows the run by a factor LPT+1, so a 1,000-entry row costs two dependent loads
// instead of ten.  Must be called by all LPT lanes of the group with identical arguments.
template <int LPT>
__device__ __forceinline__ bool group_contains(const int32_t *__restrict__ indices, int64_t lo, int64_t hi,
                                               int32_t key, int sub, unsigned gmask, int gshift) {
    while (hi - lo > LPT) {
        const int64_t span = hi - lo;
        const int64_t p = lo + (span * (sub + 1)) / (LPT + 1);
        const int32_t v = __ldg(indices + p);
        const unsigned ge = (__ballot_sync(gmask, v >= key) & gmask) >> gshift;
        const unsigned eq = (__ballot_sync(gmask, v == key) & gmask);
        if (eq) return true;
        if (ge == 0u) {
            lo = lo + (span * LPT) / (LPT + 1) + 1;
        } else {
            const int t = __ffs(ge) - 1;                  // first splitter above the key
            hi = lo + (span * (t + 1)) / (LPT + 1);
            if (t > 0) lo = lo + (span * t) / (LPT + 1) + 1;
        }
    }
    const bool mine = (lo + sub < hi) && (__ldg(indices + lo + sub) == key);
    return (__ballot_sync(gmask, mine) & gmask) != 0u;
}
// Position of `key` in the sorted run indices[lo, hi), or -1.  Same (LPT+1)-ary search as group_contains.
template <int LPT>
__device__ __forceinline__ int64_t group_find(const int32_t *__restrict__ indices, int64_t lo, int64_t hi, int32_t key,
                                              int sub, unsigned gmask, int gshift) {
    while (hi - lo > LPT) {
        const int64_t span = hi - lo;
        const int64_t p = lo + (span * (sub + 1)) / (LPT + 1);
        const int32_t v = __ldg(indices + p);
        const unsigned ge = (__ballot_sync(gmask, v >= key) & gmask) >> gshift;
        const unsigned eq = (__ballot_sync(gmask, v == key) & gmask) >> gshift;
        if (eq) { const int t = __ffs(eq) - 1; return lo + (span * (t + 1)) / (LPT + 1); }
        if (ge == 0u) {
            lo = lo + (span * LPT) / (LPT + 1) + 1;
        } else {
            const int t = __ffs(ge) - 1;
            hi = lo + (span * (t + 1)) / (LPT + 1);
            if (t > 0) lo = lo + (span * t) / (LPT + 1) + 1;
        }
    }
    const bool mine = (lo + sub < hi) && (__ldg(indices + lo + sub) == key);
    const unsigned hit = (__ballot_sync(gmask, mine) & gmask) >> gshift;
    return hit ? lo + (__ffs(hit) - 1) : (int64_t)-1;
}
#endif  // __CUDACC__

}  // namespace cymf
