// capi.cu -- status plumbing, the host-side mt19937 stream, layout conversion kernels and the
// host-buffer (`*_host`) entry points of the C ABI declared in include/cymf_b200.h.
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <charconv>
#include <cmath>
#include <string>
#include <vector>

#include "common.cuh"

namespace cymf {

static thread_local char tls_error[512] = "";
std::atomic<int64_t> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tls_error, sizeof(tls_error), fmt, ap);
    va_end(ap);
}

int cuda_status(cudaError_t e, const char *what, const char *file, int line) {
    set_error("CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
    return (int)e;
}

bool tc_enabled() {
    const char *v = getenv("CYMF_NO_TCGEN05");
    return !(v && v[0] == '1');
}

int sm_count() {
    static int cached = 0;
    if (!cached) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            cached = 148;
    }
    return cached;
}

// ---- layout conversion: dense f64 [rows, K] (host view of model.W) <-> device [rows, ld] of T, zero padded ----
template <typename T>
__global__ void pack_rows_kernel(const double *__restrict__ src, T *__restrict__ dst, int64_t rows, int K, int ld) {
    const int64_t total = rows * ld;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / ld;
        const int c = (int)(t - r * ld);
        dst[t] = c < K ? (T)src[r * K + c] : T(0);
    }
}
template <typename T>
__global__ void unpack_rows_kernel(const T *__restrict__ src, double *__restrict__ dst, int64_t rows, int K, int ld) {
    const int64_t total = rows * K;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = t / K;
        const int c = (int)(t - r * K);
        dst[t] = (double)src[r * ld + c];
    }
}
template <typename T> __global__ void fill_kernel(T *dst, int64_t n, T value) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        dst[t] = value;
}
__global__ void widen_indptr_kernel(const int32_t *__restrict__ src, int64_t *__restrict__ dst, int64_t n) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        dst[t] = src[t];
}

static inline unsigned grid_for(int64_t n) {
    int64_t b = (n + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (unsigned)b;
}

}  // namespace cymf

using namespace cymf;

extern "C" int cymf_abi_version(void) { return CYMF_ABI_VERSION; }
extern "C" const char *cymf_last_error(void) { return tls_error; }
extern "C" int64_t cymf_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

// ---- mt19937 + bounded draw: the reference's UniformGenerator (cymf/math.pyx:12-18) ---------------------------
struct cymf_rng {
    uint32_t state[624];
    int next;
    explicit cymf_rng(uint32_t seed) : next(624) {
        uint32_t prev = state[0] = seed;
        for (uint32_t t = 1; t < 624; ++t) prev = state[t] = 1812433253u * (prev ^ (prev >> 30)) + t;
    }
    void twist() {
        auto mix = [](uint32_t upper, uint32_t lower, uint32_t far) {
            const uint32_t y = (upper & 0x80000000u) | (lower & 0x7fffffffu);
            return far ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        };
        for (int t = 0; t < 227; ++t) state[t] = mix(state[t], state[t + 1], state[t + 397]);
        for (int t = 227; t < 623; ++t) state[t] = mix(state[t], state[t + 1], state[t - 227]);
        state[623] = mix(state[623], state[0], state[396]);
        next = 0;
    }
    uint32_t word() {
        if (next == 624) twist();
        uint32_t y = state[next++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        return y ^ (y >> 18);
    }
    // std::uniform_int_distribution<long>(0, n-1) on a 32-bit engine, libstdc++ >= 11: multiply-shift,
    // redraw while the low word falls in the biased zone
    uint32_t below(uint32_t n) {
        uint64_t m = (uint64_t)word() * n;
        if ((uint32_t)m < n) {
            const uint32_t zone = (0u - n) % n;
            while ((uint32_t)m < zone) m = (uint64_t)word() * n;
        }
        return (uint32_t)(m >> 32);
    }
    // the same distribution over any n >= 1: a range of exactly 2^32 takes the raw word, wider ranges draw the
    // high part recursively, add a raw low word and redraw while the sum leaves the range (libstdc++'s upscaling)
    uint64_t below64(uint64_t n) {
        const uint64_t top = n - 1;
        if (top < 0xffffffffull) return below((uint32_t)n);
        if (top == 0xffffffffull) return word();
        for (;;) {
            const uint64_t high = below64((top >> 32) + 1) << 32;
            const uint64_t v = high + word();
            if (v <= top && v >= high) return v;
        }
    }
};

extern "C" cymf_rng *cymf_rng_create(uint32_t seed) { return new (std::nothrow) cymf_rng(seed); }
extern "C" void cymf_rng_destroy(cymf_rng *g) { delete g; }
extern "C" int cymf_rng_fill_below(cymf_rng *g, uint32_t n, int32_t *out, int64_t count) {
    CYMF_REQUIRE(g && out && n > 0 && count >= 0, "bad argument");
    for (int64_t t = 0; t < count; ++t) out[t] = (int32_t)g->below(n);
    return 0;
}

extern "C" int cymf_rng_fill_below64(cymf_rng *g, uint64_t n, int64_t *out, int64_t count) {
    CYMF_REQUIRE(g && out && n > 0 && count >= 0, "bad argument");
    for (int64_t t = 0; t < count; ++t) out[t] = (int64_t)g->below64(n);
    return 0;
}

// The Hogwild kernels' negative for triplet l of `epoch` (same __host__ __device__ code the kernel runs), so a
// run can be audited or replayed on the host.
extern "C" int cymf_bpr_negatives_host(uint64_t seed, uint32_t epoch, int64_t first, int64_t count, uint32_t n,
                                       int32_t *out) {
    CYMF_REQUIRE(out && n > 0 && first >= 0 && count >= 0, "bad argument");
    for (int64_t t = 0; t < count; ++t) out[t] = (int32_t)philox_negative(seed, epoch, (uint64_t)(first + t), n);
    return 0;
}

// ---- helpers exported for the Python host (device pointers) ---------------------------------------------------
extern "C" int cymf_pack_rows_dev(const double *src, void *dst, int dtype, int64_t rows, int32_t K, int32_t ld,
                                  void *stream) {
    CYMF_REQUIRE(src && dst && rows >= 0 && K > 0 && ld >= K, "bad argument");
    if (rows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CYMF_F32) pack_rows_kernel<float><<<grid_for(rows * ld), 256, 0, st>>>(src, (float *)dst, rows, K, ld);
    else pack_rows_kernel<double><<<grid_for(rows * ld), 256, 0, st>>>(src, (double *)dst, rows, K, ld);
    CYMF_LAUNCHED();
    return 0;
}
extern "C" int cymf_unpack_rows_dev(const void *src, double *dst, int dtype, int64_t rows, int32_t K, int32_t ld,
                                    void *stream) {
    CYMF_REQUIRE(src && dst && rows >= 0 && K > 0 && ld >= K, "bad argument");
    if (rows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CYMF_F32)
        unpack_rows_kernel<float><<<grid_for(rows * K), 256, 0, st>>>((const float *)src, dst, rows, K, ld);
    else
        unpack_rows_kernel<double><<<grid_for(rows * K), 256, 0, st>>>((const double *)src, dst, rows, K, ld);
    CYMF_LAUNCHED();
    return 0;
}
extern "C" int cymf_fill_dev(void *dst, int dtype, int64_t n, double value, void *stream) {
    CYMF_REQUIRE(dst && n >= 0, "bad argument");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CYMF_F32) fill_kernel<float><<<grid_for(n), 256, 0, st>>>((float *)dst, n, (float)value);
    else fill_kernel<double><<<grid_for(n), 256, 0, st>>>((double *)dst, n, value);
    CYMF_LAUNCHED();
    return 0;
}

// dense f64 vector -> vector of `dtype` (propensities, X.data)
extern "C" int cymf_convert_dev(const double *src, void *dst, int dtype, int64_t n, void *stream) {
    CYMF_REQUIRE(src && dst && n >= 0, "bad argument");
    if (n == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CYMF_F32) pack_rows_kernel<float><<<grid_for(n), 256, 0, st>>>(src, (float *)dst, n, 1, 1);
    else pack_rows_kernel<double><<<grid_for(n), 256, 0, st>>>(src, (double *)dst, n, 1, 1);
    CYMF_LAUNCHED();
    return 0;
}

// ---- host-buffer BPR fit ---------------------------------------------------------------------------------------
extern "C" int cymf_bpr_fit_host(double *W, double *H, int32_t U, int32_t I, int32_t K,
                                 const int32_t *users, const int32_t *positives, int64_t N,
                                 const int32_t *indptr, const int32_t *indices,
                                 int32_t num_epochs, double learning_rate, double weight_decay,
                                 int optimizer, int mode, uint64_t seed, int64_t *applied_out) {
    CYMF_REQUIRE(W && H && users && positives && indptr && indices, "null pointer");
    CYMF_REQUIRE(U > 0 && I > 0 && K > 0 && N >= 0 && num_epochs >= 0, "bad shape");
    CYMF_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0 (hogwild f32), 1 (hogwild f64) or 2 (replay f64)");
    CYMF_REQUIRE(optimizer >= CYMF_SGD && optimizer <= CYMF_ADAM, "unknown optimizer");
    int ndev = 0;
    CYMF_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev < 1) { set_error("no CUDA device: cymf_b200 has no CPU fallback"); return CYMF_EUNSUPPORTED; }

    const int dtype = mode == 0 ? CYMF_F32 : CYMF_F64;
    const size_t es = dtype == CYMF_F32 ? 4 : 8;
    const int32_t ld = (K + 3) / 4 * 4;
    const int64_t nnz = indptr[U];
    DeviceArena mem;
    cudaStream_t st = nullptr;

    double *stage = nullptr;                                   // f64 staging for pack / unpack
    const int64_t big = (int64_t)(U > I ? U : I) * K;
    CYMF_TRY(mem.get(&stage, (size_t)big * 8));
    cymf_factors f{};
    CYMF_TRY(mem.get((char **)&f.W, (size_t)U * ld * es));
    CYMF_TRY(mem.get((char **)&f.H, (size_t)I * ld * es));
    CYMF_CUDA(cudaMemcpyAsync(stage, W, (size_t)U * K * 8, cudaMemcpyHostToDevice, st));
    CYMF_TRY(cymf_pack_rows_dev(stage, f.W, dtype, U, K, ld, st));
    CYMF_CUDA(cudaMemcpyAsync(stage, H, (size_t)I * K * 8, cudaMemcpyHostToDevice, st));
    CYMF_TRY(cymf_pack_rows_dev(stage, f.H, dtype, I, K, ld, st));
    if (optimizer != CYMF_SGD) {                               // state rebuilt on every fit (bpr.pyx:149-156)
        const double init = optimizer == CYMF_ADAGRAD ? 1.0 : 0.0;
        CYMF_TRY(mem.get((char **)&f.s1W, (size_t)U * ld * es));
        CYMF_TRY(mem.get((char **)&f.s1H, (size_t)I * ld * es));
        CYMF_TRY(cymf_fill_dev(f.s1W, dtype, (int64_t)U * ld, init, st));
        CYMF_TRY(cymf_fill_dev(f.s1H, dtype, (int64_t)I * ld, init, st));
        if (optimizer == CYMF_ADAM) {
            CYMF_TRY(mem.get((char **)&f.s2W, (size_t)U * ld * es));
            CYMF_TRY(mem.get((char **)&f.s2H, (size_t)I * ld * es));
            CYMF_TRY(cymf_fill_dev(f.s2W, dtype, (int64_t)U * ld, 0.0, st));
            CYMF_TRY(cymf_fill_dev(f.s2H, dtype, (int64_t)I * ld, 0.0, st));
        }
    }
    int32_t *d_users, *d_pos, *d_idx, *d_ip32, *d_neg = nullptr;
    int64_t *d_ip;
    unsigned long long *d_applied;
    CYMF_TRY(mem.get(&d_users, (size_t)N * 4));
    CYMF_TRY(mem.get(&d_pos, (size_t)N * 4));
    CYMF_TRY(mem.get(&d_idx, (size_t)nnz * 4));
    CYMF_TRY(mem.get(&d_ip32, (size_t)(U + 1) * 4));
    CYMF_TRY(mem.get(&d_ip, (size_t)(U + 1) * 8));
    CYMF_TRY(mem.get(&d_applied, 8));
    CYMF_CUDA(cudaMemcpyAsync(d_users, users, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    CYMF_CUDA(cudaMemcpyAsync(d_pos, positives, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    CYMF_CUDA(cudaMemcpyAsync(d_idx, indices, (size_t)nnz * 4, cudaMemcpyHostToDevice, st));
    CYMF_CUDA(cudaMemcpyAsync(d_ip32, indptr, (size_t)(U + 1) * 4, cudaMemcpyHostToDevice, st));
    widen_indptr_kernel<<<grid_for(U + 1), 256, 0, st>>>(d_ip32, d_ip, U + 1);
    CYMF_LAUNCHED();
    CYMF_CUDA(cudaMemsetAsync(d_applied, 0, 8, st));

    if (mode == 2) {
        CYMF_TRY(mem.get(&d_neg, (size_t)N * 4));
        cymf_rng gen((uint32_t)seed);                          // bpr.pyx:141: one generator for the whole fit
        std::vector<int32_t> neg((size_t)N);
        for (int32_t epoch = 0; epoch < num_epochs; ++epoch) {
            for (int64_t t = 0; t < N; ++t) neg[(size_t)t] = (int32_t)gen.below((uint32_t)I);
            CYMF_CUDA(cudaMemcpyAsync(d_neg, neg.data(), (size_t)N * 4, cudaMemcpyHostToDevice, st));
            CYMF_TRY(cymf_bpr_replay_epoch_dev(&f, optimizer, d_users, d_pos, d_neg, N, d_ip, d_idx, U, I, K, ld,
                                               learning_rate, weight_decay, d_applied, st));
            CYMF_CUDA(cudaStreamSynchronize(st));              // neg[] is reused by the next epoch
        }
    } else {
        const int scatter = dtype == CYMF_F32 ? 1 : 0;      // 128-bit reductions: parameters and optimizer state
        const int64_t inflight = N / 256 > 1024 ? N / 256 : 1024;
        for (int32_t epoch = 0; epoch < num_epochs; ++epoch)
            CYMF_TRY(cymf_bpr_hogwild_epoch_dev(&f, dtype, optimizer, scatter, d_users, d_pos, N, d_ip, d_idx, U, I, K,
                                                ld, learning_rate, weight_decay, seed, (uint32_t)epoch, inflight,
                                                d_applied, st));
    }
    CYMF_TRY(cymf_unpack_rows_dev(f.W, stage, dtype, U, K, ld, st));
    CYMF_CUDA(cudaMemcpyAsync(W, stage, (size_t)U * K * 8, cudaMemcpyDeviceToHost, st));
    CYMF_TRY(cymf_unpack_rows_dev(f.H, stage, dtype, I, K, ld, st));
    CYMF_CUDA(cudaMemcpyAsync(H, stage, (size_t)I * K * 8, cudaMemcpyDeviceToHost, st));
    unsigned long long applied = 0;
    CYMF_CUDA(cudaMemcpyAsync(&applied, d_applied, 8, cudaMemcpyDeviceToHost, st));
    CYMF_CUDA(cudaStreamSynchronize(st));
    if (applied_out) *applied_out = (int64_t)applied;
    return 0;
}

// ---- host-buffer RelMF fit -------------------------------------------------------------------------------------
extern "C" int cymf_relmf_fit_host(double *W, double *H, int32_t U, int32_t I, int32_t K,
                                   const int32_t *indptr, const int32_t *indices, const double *values,
                                   const double *propensities, int32_t num_epochs, double learning_rate,
                                   double weight_decay, double clip_value, int optimizer, int mode, uint64_t seed) {
    CYMF_REQUIRE(W && H && indptr && indices && propensities, "null pointer");
    CYMF_REQUIRE(U > 0 && I > 0 && K > 0 && num_epochs >= 0, "bad shape");
    CYMF_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0 (hogwild f32), 1 (hogwild f64) or 2 (replay f64)");
    CYMF_REQUIRE(optimizer >= CYMF_SGD && optimizer <= CYMF_ADAM, "unknown optimizer");
    int ndev = 0;
    CYMF_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev < 1) { set_error("no CUDA device: cymf_b200 has no CPU fallback"); return CYMF_EUNSUPPORTED; }

    const int dtype = mode == 0 ? CYMF_F32 : CYMF_F64;
    const size_t es = dtype == CYMF_F32 ? 4 : 8;
    const int32_t ld = (K + 3) / 4 * 4;
    const int64_t nnz = indptr[U];
    const int64_t n_samples = (int64_t)U * I;                   // relmf.pyx:121: N = U * I cells per epoch
    DeviceArena mem;
    cudaStream_t st = nullptr;

    double *stage = nullptr;
    int64_t big = (int64_t)(U > I ? U : I) * K;
    if (big < nnz) big = nnz;
    CYMF_TRY(mem.get(&stage, (size_t)big * 8));
    cymf_factors f{};
    CYMF_TRY(mem.get((char **)&f.W, (size_t)U * ld * es));
    CYMF_TRY(mem.get((char **)&f.H, (size_t)I * ld * es));
    CYMF_CUDA(cudaMemcpyAsync(stage, W, (size_t)U * K * 8, cudaMemcpyHostToDevice, st));
    CYMF_TRY(cymf_pack_rows_dev(stage, f.W, dtype, U, K, ld, st));
    CYMF_CUDA(cudaMemcpyAsync(stage, H, (size_t)I * K * 8, cudaMemcpyHostToDevice, st));
    CYMF_TRY(cymf_pack_rows_dev(stage, f.H, dtype, I, K, ld, st));
    if (optimizer != CYMF_SGD) {                               // state rebuilt on every fit (relmf.pyx:129-137)
        const double init = optimizer == CYMF_ADAGRAD ? 1.0 : 0.0;
        CYMF_TRY(mem.get((char **)&f.s1W, (size_t)U * ld * es));
        CYMF_TRY(mem.get((char **)&f.s1H, (size_t)I * ld * es));
        CYMF_TRY(cymf_fill_dev(f.s1W, dtype, (int64_t)U * ld, init, st));
        CYMF_TRY(cymf_fill_dev(f.s1H, dtype, (int64_t)I * ld, init, st));
        if (optimizer == CYMF_ADAM) {
            CYMF_TRY(mem.get((char **)&f.s2W, (size_t)U * ld * es));
            CYMF_TRY(mem.get((char **)&f.s2H, (size_t)I * ld * es));
            CYMF_TRY(cymf_fill_dev(f.s2W, dtype, (int64_t)U * ld, 0.0, st));
            CYMF_TRY(cymf_fill_dev(f.s2H, dtype, (int64_t)I * ld, 0.0, st));
        }
    }
    int32_t *d_idx, *d_ip32;
    int64_t *d_ip, *d_cells = nullptr;
    char *d_val = nullptr, *d_prop;
    CYMF_TRY(mem.get(&d_idx, (size_t)nnz * 4));
    CYMF_TRY(mem.get(&d_ip32, (size_t)(U + 1) * 4));
    CYMF_TRY(mem.get(&d_ip, (size_t)(U + 1) * 8));
    CYMF_TRY(mem.get(&d_prop, (size_t)I * es));
    CYMF_CUDA(cudaMemcpyAsync(d_idx, indices, (size_t)nnz * 4, cudaMemcpyHostToDevice, st));
    CYMF_CUDA(cudaMemcpyAsync(d_ip32, indptr, (size_t)(U + 1) * 4, cudaMemcpyHostToDevice, st));
    widen_indptr_kernel<<<grid_for(U + 1), 256, 0, st>>>(d_ip32, d_ip, U + 1);
    CYMF_LAUNCHED();
    // 1-column "matrices": pack converts f64 -> dtype
    CYMF_CUDA(cudaMemcpyAsync(stage, propensities, (size_t)I * 8, cudaMemcpyHostToDevice, st));
    CYMF_TRY(cymf_convert_dev(stage, d_prop, dtype, I, st));
    if (values && nnz > 0) {
        CYMF_TRY(mem.get(&d_val, (size_t)nnz * es));
        CYMF_CUDA(cudaMemcpyAsync(stage, values, (size_t)nnz * 8, cudaMemcpyHostToDevice, st));
        CYMF_TRY(cymf_convert_dev(stage, d_val, dtype, nnz, st));
    }

    if (mode == 2) {
        CYMF_TRY(mem.get(&d_cells, (size_t)n_samples * 8));
        cymf_rng gen((uint32_t)seed);                          // relmf.pyx:127: one generator for the whole fit
        std::vector<int64_t> cells((size_t)n_samples);
        for (int32_t epoch = 0; epoch < num_epochs; ++epoch) {
            for (int64_t t = 0; t < n_samples; ++t) cells[(size_t)t] = (int64_t)gen.below64((uint64_t)n_samples);
            CYMF_CUDA(cudaMemcpyAsync(d_cells, cells.data(), (size_t)n_samples * 8, cudaMemcpyHostToDevice, st));
            CYMF_TRY(cymf_relmf_replay_epoch_dev(&f, optimizer, d_cells, n_samples, d_ip, d_idx, (const double *)d_val,
                                                 (const double *)d_prop, U, I, K, ld, learning_rate, weight_decay,
                                                 clip_value, st));
            CYMF_CUDA(cudaStreamSynchronize(st));
        }
    } else {
        const int scatter = dtype == CYMF_F32 ? 1 : 0;      // 128-bit reductions: parameters and optimizer state
        const int64_t inflight = n_samples / 256 > 1024 ? n_samples / 256 : 1024;
        for (int32_t epoch = 0; epoch < num_epochs; ++epoch)
            CYMF_TRY(cymf_relmf_hogwild_epoch_dev(&f, dtype, optimizer, scatter, d_ip, d_idx, d_val, d_prop, U, I, K, ld,
                                                  n_samples, learning_rate, weight_decay, clip_value, seed,
                                                  (uint32_t)epoch, inflight, st));
    }
    CYMF_TRY(cymf_unpack_rows_dev(f.W, stage, dtype, U, K, ld, st));
    CYMF_CUDA(cudaMemcpyAsync(W, stage, (size_t)U * K * 8, cudaMemcpyDeviceToHost, st));
    CYMF_TRY(cymf_unpack_rows_dev(f.H, stage, dtype, I, K, ld, st));
    CYMF_CUDA(cudaMemcpyAsync(H, stage, (size_t)I * K * 8, cudaMemcpyDeviceToHost, st));
    CYMF_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// ---- GloVe.save_word2vec_format (cymf/glove.pyx:164-177) ---------------------------------------------------------
// The reference writes  f"{word} " + " ".join(map(str, W[i])) + "\n"  per row: str(np.float64) is the shortest
// round-trip decimal laid out by CPython's repr rules.  std::to_chars (scientific) yields the same shortest digits;
// the layout below restates format_float_short('r'): exponent form when decpt <= -4 or decpt > 16 (mantissa without
// ".0", exponent of at least two digits), otherwise positional with ".0" appended to integral values.
static char *py_float_repr(double x, char *out) {
    if (x != x) { memcpy(out, "nan", 3); return out + 3; }
    if (x < 0 || (x == 0 && std::signbit(x))) { *out++ = '-'; x = -x; }
    if (x > 1.7976931348623157e308) { memcpy(out, "inf", 3); return out + 3; }
    char buf[48];
    const auto r = std::to_chars(buf, buf + sizeof(buf), x, std::chars_format::scientific);
    char digits[24];
    int nd = 0, exp10 = 0;
    const char *p = buf;
    for (; p < r.ptr && *p != 'e'; ++p)
        if (*p != '.') digits[nd++] = *p;
    if (p < r.ptr) {                                              // "e[+-]dd"
        ++p;
        const bool neg = *p == '-';
        if (*p == '+' || *p == '-') ++p;
        for (; p < r.ptr; ++p) exp10 = exp10 * 10 + (*p - '0');
        if (neg) exp10 = -exp10;
    }
    while (nd > 1 && digits[nd - 1] == '0') --nd;                 // to_chars never pads, kept for safety
    const int decpt = exp10 + 1;
    if (decpt <= -4 || decpt > 16) {
        *out++ = digits[0];
        if (nd > 1) { *out++ = '.'; memcpy(out, digits + 1, nd - 1); out += nd - 1; }
        *out++ = 'e';
        int e = decpt - 1;
        *out++ = e < 0 ? '-' : '+';
        if (e < 0) e = -e;
        char eb[8];
        int ne = 0;
        do { eb[ne++] = (char)('0' + e % 10); e /= 10; } while (e);
        if (ne < 2) eb[ne++] = '0';
        while (ne) *out++ = eb[--ne];
    } else if (decpt <= 0) {
        *out++ = '0'; *out++ = '.';
        for (int z = 0; z < -decpt; ++z) *out++ = '0';
        memcpy(out, digits, nd); out += nd;
    } else if (decpt >= nd) {
        memcpy(out, digits, nd); out += nd;
        for (int z = 0; z < decpt - nd; ++z) *out++ = '0';
        *out++ = '.'; *out++ = '0';
    } else {
        memcpy(out, digits, decpt); out += decpt;
        *out++ = '.';
        memcpy(out, digits + decpt, nd - decpt); out += nd - decpt;
    }
    return out;
}

// W: dense row-major f64 [V, K] HOST array; words: V NUL-terminated byte strings (already encoded).  Writes the
// gensim word2vec text format byte for byte as the reference does.
extern "C" int cymf_word2vec_write_host(const char *path, const double *W, int64_t V, int32_t K,
                                        const char *const *words) {
    CYMF_REQUIRE(path && (W || V == 0 || K == 0) && (words || V == 0) && V >= 0 && K >= 0, "bad argument");
    FILE *f = fopen(path, "wb");
    if (!f) { set_error("cannot open %s for writing", path); return CYMF_EINVAL; }
    std::string line;
    line = std::to_string(V) + " " + std::to_string(K) + "\n";
    bool ok = fwrite(line.data(), 1, line.size(), f) == line.size();
    std::vector<char> row((size_t)K * 32 + 16);
    for (int64_t i = 0; i < V && ok; ++i) {
        const size_t wl = strlen(words[i]);
        ok = fwrite(words[i], 1, wl, f) == wl && fputc(' ', f) != EOF;
        char *o = row.data();
        for (int32_t k = 0; k < K; ++k) {
            if (k) *o++ = ' ';
            o = py_float_repr(W[(size_t)i * K + k], o);
        }
        *o++ = '\n';
        ok = ok && fwrite(row.data(), 1, (size_t)(o - row.data()), f) == (size_t)(o - row.data());
    }
    if (fclose(f) != 0) ok = false;
    if (!ok) { set_error("write to %s failed", path); return CYMF_EINVAL; }
    return 0;
}
