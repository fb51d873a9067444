// glove.cu -- GloVe AdaGrad sample update (replaces the prange loop of cymf/glove.pyx:151-153 with
// GloVeModel.forward/backward, cymf/model.pyx:166-204, and GloVeAdaGrad, cymf/optimizer.pyx:85-123).
//
// Same gather/scatter shape as the BPR kernel: one lane group per co-occurrence sample (c, x, count); four
// rows (W[c], H[x] and their AdaGrad accumulators) and four scalars (two biases + accumulators) are read,
// updated and written back.  Algorithmic bytes per sample = 8*K*s + 8*s + 8 + s  (s = sizeof(T)).
//
// Quirk kept on purpose (model.pyx:195-204): the two bias updates sit inside the k loop, so each sample
// applies K AdaGrad steps with the same gradient g to each bias:
//     acc_t = acc_0 + t g^2,   b -= lr g sum_{t=1..K} 1/sqrt(acc_0 + t g^2).
// The Hogwild kernel evaluates that sum in parallel over the lanes of the group; the replay kernel runs the K
// steps one after the other, exactly like the reference.
#include <math.h>

#include "common.cuh"

namespace cymf {

template <typename T> struct GloveArgs {
    T *W, *H, *bW, *bH, *aW, *aH, *abW, *abH;
    const int32_t *central, *context;
    const T *counts;
    int64_t N, groups;
    int32_t K, ld;
    T lr, x_max, alpha;
    double *loss_sum;
};

__device__ __forceinline__ float glove_weight(float n, float x_max, float alpha) {
    return fminf(__powf(n / x_max, alpha), 1.0f);                 // model.pyx:34-35
}
__device__ __forceinline__ double glove_weight(double n, double x_max, double alpha) {
    return fmin(pow(n / x_max, alpha), 1.0);
}
__device__ __forceinline__ float log_t(float x) { return __logf(x); }
__device__ __forceinline__ double log_t(double x) { return log(x); }
__device__ __forceinline__ float rsq(float x) { return rsqrtf(x); }
__device__ __forceinline__ double rsq(double x) { return 1.0 / sqrt(x); }
__device__ __forceinline__ void red_add_scalar(float *p, float v) { asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory"); }
__device__ __forceinline__ void red_add_scalar(double *p, double v) { asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory"); }

template <typename T, int LPT, int NV, bool RED>
__global__ void __launch_bounds__(256) glove_hogwild_kernel(const GloveArgs<T> a) {
    constexpr int GPW = 32 / LPT;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPT;
    const unsigned gmask = group_mask<LPT>(lane);
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t stride = a.groups;
    const int64_t gid = warp * GPW + lane / LPT;
    const bool active = gid < stride;
    double loss_local = 0.0;

    int64_t l = gid;
    int32_t c_next = 0, x_next = 0;
    T n_next = T(1);
    if (active && l < a.N) { c_next = __ldcs(a.central + l); x_next = __ldcs(a.context + l); n_next = __ldcs(a.counts + l); }

    for (int64_t base = warp * GPW; base < a.N && warp * GPW < stride; base += stride, l += stride) {
        const bool valid = active && l < a.N;
        const int32_t c = c_next, x = x_next;
        const T n = n_next;
        const int64_t ln = l + stride;
        if (active && ln < a.N) { c_next = __ldcs(a.central + ln); x_next = __ldcs(a.context + ln); n_next = __ldcs(a.counts + ln); }

        T *pw = a.W + (size_t)c * a.ld, *ph = a.H + (size_t)x * a.ld;
        T *paw = a.aW + (size_t)c * a.ld, *pah = a.aH + (size_t)x * a.ld;
        Slot<T> w[NV], h[NV], aw[NV], ah[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int e = (sub + v * LPT) * 4;
            if (e < a.ld) { w[v] = load_slot(pw + e); h[v] = load_slot(ph + e); aw[v] = load_slot(paw + e); ah[v] = load_slot(pah + e); }
            else { w[v] = zero_slot<T>(); h[v] = zero_slot<T>(); aw[v] = zero_slot<T>(); ah[v] = zero_slot<T>(); }
        }
        const T bw = __ldcg(a.bW + c), bh = __ldcg(a.bH + x), abw = __ldcg(a.abW + c), abh = __ldcg(a.abH + x);

        T dot = T(0);                                                          // model.pyx:174-175
#pragma unroll
        for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int e = 0; e < 4; ++e) dot += w[v].v[e] * h[v].v[e];
        dot = group_sum<LPT>(dot, gmask);
        const T raw = dot + (bw + bh) - log_t(n);                              // model.pyx:176-177
        const T g = raw * glove_weight(n, a.x_max, a.alpha);                   // model.pyx:179
        if (a.loss_sum && valid && sub == 0) loss_local += 0.5 * (double)g * (double)raw;

        // K AdaGrad steps on each bias with the same gradient (model.pyx:199-204): partial sums over the lanes
        const T g2 = g * g;
        T sw = T(0), sh = T(0);
        for (int t = sub + 1; t <= a.K; t += LPT) {
            sw += rsq(abw + (T)t * g2);
            sh += rsq(abh + (T)t * g2);
        }
        sw = group_sum<LPT>(sw, gmask);
        sh = group_sum<LPT>(sh, gmask);

        if (valid) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int e0 = (sub + v * LPT) * 4;
                if (e0 >= a.ld) continue;
                Slot<T> dw, dh, daw, dah;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const T gw = g * h[v].v[e], gh = g * w[v].v[e];            // model.pyx:196-197 (pre-update values)
                    daw.v[e] = gw * gw;                                        // optimizer.pyx:103-111
                    dah.v[e] = gh * gh;
                    const T naw = aw[v].v[e] + daw.v[e], nah = ah[v].v[e] + dah.v[e];
                    // pad columns: w = h = 0 -> g* = 0, accumulators (0) stay 0, and 0 * rsqrt(0) must not make NaN
                    dw.v[e] = gw == T(0) ? T(0) : -a.lr * gw * rsq(naw);
                    dh.v[e] = gh == T(0) ? T(0) : -a.lr * gh * rsq(nah);
                    if (!RED) { dw.v[e] += w[v].v[e]; dh.v[e] += h[v].v[e]; daw.v[e] = naw; dah.v[e] = nah; }
                }
                if (RED) { red_add_slot(pw + e0, dw); red_add_slot(ph + e0, dh); red_add_slot(paw + e0, daw); red_add_slot(pah + e0, dah); }
                else     { store_slot(pw + e0, dw);   store_slot(ph + e0, dh);   store_slot(paw + e0, daw);   store_slot(pah + e0, dah); }
            }
            if (sub == 0) {
                const T kg2 = (T)a.K * g2;
                if (RED) {
                    red_add_scalar(a.bW + c, -a.lr * g * sw); red_add_scalar(a.bH + x, -a.lr * g * sh);
                    red_add_scalar(a.abW + c, kg2);           red_add_scalar(a.abH + x, kg2);
                } else {
                    __stcg(a.bW + c, bw - a.lr * g * sw);     __stcg(a.bH + x, bh - a.lr * g * sh);
                    __stcg(a.abW + c, abw + kg2);             __stcg(a.abH + x, abh + kg2);
                }
            }
        }
    }
    if (a.loss_sum) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) loss_local += __shfl_xor_sync(0xffffffffu, loss_local, off);
        if (lane == 0 && loss_local != 0.0) atomicAdd(a.loss_sum, loss_local);
    }
}

// ---- serialized f64 replay (reference operation order, num_threads = 1) ---------------------------------------
constexpr int GLOVE_REPLAY_MAX_K = 1024;

__global__ void __launch_bounds__(32) glove_replay_kernel(const GloveArgs<double> a, double *loss /* [N] or NULL */) {
    __shared__ double prod[GLOVE_REPLAY_MAX_K];
    const int lane = threadIdx.x;
    const int K = a.K;
    for (int64_t l = 0; l < a.N; ++l) {
        const int32_t c = a.central[l], x = a.context[l];
        const double n = a.counts[l];
        double *pw = a.W + (size_t)c * a.ld, *ph = a.H + (size_t)x * a.ld;
        double *paw = a.aW + (size_t)c * a.ld, *pah = a.aH + (size_t)x * a.ld;
        for (int k = lane; k < K; k += 32) prod[k] = __dmul_rn(__ldcg(pw + k), __ldcg(ph + k));
        __syncwarp();
        double d = 0.0;
        for (int k = 0; k < K; ++k) d = __dadd_rn(d, prod[k]);                 // model.pyx:174-175, k ascending
        __syncwarp();
        double bw = __ldcg(a.bW + c), bh = __ldcg(a.bH + x);
        d = __dadd_rn(d, __dadd_rn(bw, bh));                                   // model.pyx:176
        d = __dsub_rn(d, log(n));                                              // model.pyx:177
        const double raw = d;
        const double g = __dmul_rn(d, fmin(pow(__ddiv_rn(n, a.x_max), a.alpha), 1.0));   // model.pyx:179
        if (loss && lane == 0) loss[l] = __dmul_rn(__dmul_rn(0.5, g), raw);    // model.pyx:180
        for (int k = lane; k < K; k += 32) {                                   // model.pyx:195-198
            const double wk = __ldcg(pw + k), hk = __ldcg(ph + k);
            const double gw = __dmul_rn(g, hk), gh = __dmul_rn(g, wk);
            const double naw = __dadd_rn(__ldcg(paw + k), __dmul_rn(gw, gw));
            const double nah = __dadd_rn(__ldcg(pah + k), __dmul_rn(gh, gh));
            __stcg(paw + k, naw);
            __stcg(pah + k, nah);
            __stcg(pw + k, __dsub_rn(wk, __ddiv_rn(__dmul_rn(a.lr, gw), __dsqrt_rn(naw))));
            __stcg(ph + k, __dsub_rn(hk, __ddiv_rn(__dmul_rn(a.lr, gh), __dsqrt_rn(nah))));
        }
        // model.pyx:199-204: K sequential bias steps (all lanes compute the same chain; lane 0 stores)
        double abw = __ldcg(a.abW + c), abh = __ldcg(a.abH + x);
        const double g2 = __dmul_rn(g, g), lg = __dmul_rn(a.lr, g);
        for (int k = 0; k < K; ++k) {
            abw = __dadd_rn(abw, g2);
            bw = __dsub_rn(bw, __ddiv_rn(lg, __dsqrt_rn(abw)));
            abh = __dadd_rn(abh, g2);
            bh = __dsub_rn(bh, __ddiv_rn(lg, __dsqrt_rn(abh)));
        }
        if (lane == 0) {
            __stcg(a.bW + c, bw); __stcg(a.abW + c, abw);
            __stcg(a.bH + x, bh); __stcg(a.abH + x, abh);
        }
        __syncwarp();
        __threadfence_block();
    }
}

template <typename T, int LPT, int NV, bool RED>
static int launch_glove(const GloveArgs<T> &a, int64_t max_groups, cudaStream_t st) {
    auto kern = glove_hogwild_kernel<T, LPT, NV, RED>;
    int per_sm = 0;
    CYMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));
    if (per_sm < 1) per_sm = 1;
    constexpr int GPB = 8 * (32 / LPT);
    int64_t blocks = (int64_t)sm_count() * per_sm;
    const int64_t need = (a.N + GPB - 1) / GPB;
    if (blocks > need) blocks = need;
    if (max_groups > 0 && blocks * GPB > max_groups) blocks = (max_groups + GPB - 1) / GPB;
    if (blocks < 1) blocks = 1;
    GloveArgs<T> b = a;
    b.groups = blocks * GPB;
    if (max_groups > 0 && b.groups > max_groups) b.groups = max_groups;
    kern<<<(unsigned)blocks, 256, 0, st>>>(b);
    CYMF_LAUNCHED();
    return 0;
}

template <typename T, bool RED>
static int dispatch_glove(const GloveArgs<T> &a, int64_t max_groups, cudaStream_t st) {
    const int slots = a.ld / 4;
    if (slots <= 4)   return launch_glove<T, 4, 1, RED>(a, max_groups, st);
    if (slots <= 8)   return launch_glove<T, 8, 1, RED>(a, max_groups, st);
    if (slots <= 16)  return launch_glove<T, 16, 1, RED>(a, max_groups, st);
    if (slots <= 32)  return launch_glove<T, 32, 1, RED>(a, max_groups, st);
    if (slots <= 64)  return launch_glove<T, 32, 2, RED>(a, max_groups, st);
    if (slots <= 96)  return launch_glove<T, 32, 3, RED>(a, max_groups, st);
    if (slots <= 128) return launch_glove<T, 32, 4, RED>(a, max_groups, st);
    set_error("glove: num_components > 512 is not supported (ld=%d)", a.ld);
    return CYMF_EUNSUPPORTED;
}

template <typename T>
static void fill_args(GloveArgs<T> &a, const cymf_glove_params *p, const int32_t *central, const int32_t *context,
                      const void *counts, int64_t N, int32_t K, int32_t ld, double lr, double x_max, double alpha) {
    a.W = (T *)p->W; a.H = (T *)p->H; a.bW = (T *)p->bW; a.bH = (T *)p->bH;
    a.aW = (T *)p->aW; a.aH = (T *)p->aH; a.abW = (T *)p->abW; a.abH = (T *)p->abH;
    a.central = central; a.context = context; a.counts = (const T *)counts;
    a.N = N; a.K = K; a.ld = ld; a.lr = (T)lr; a.x_max = (T)x_max; a.alpha = (T)alpha;
}

}  // namespace cymf

using namespace cymf;

static bool glove_params_ok(const cymf_glove_params *p) {
    return p && p->W && p->H && p->bW && p->bH && p->aW && p->aH && p->abW && p->abH;
}

extern "C" int cymf_glove_hogwild_epoch_dev(const cymf_glove_params *p, int dtype, int scatter,
                                            const int32_t *central, const int32_t *context, const void *counts,
                                            int64_t N, int32_t K, int32_t ld, double learning_rate, double x_max,
                                            double alpha, int64_t max_inflight, double *loss_sum, void *stream) {
    CYMF_REQUIRE(glove_params_ok(p) && central && context && counts, "null pointer");
    CYMF_REQUIRE(K > 0 && ld >= K && ld % 4 == 0, "bad shape (ld must be a multiple of 4, >= K)");
    if (N <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CYMF_F32) {
        GloveArgs<float> a{};
        fill_args(a, p, central, context, counts, N, K, ld, learning_rate, x_max, alpha);
        a.loss_sum = loss_sum;
        return scatter ? dispatch_glove<float, true>(a, max_inflight, st) : dispatch_glove<float, false>(a, max_inflight, st);
    } else if (dtype == CYMF_F64) {
        GloveArgs<double> a{};
        fill_args(a, p, central, context, counts, N, K, ld, learning_rate, x_max, alpha);
        a.loss_sum = loss_sum;
        return scatter ? dispatch_glove<double, true>(a, max_inflight, st) : dispatch_glove<double, false>(a, max_inflight, st);
    }
    set_error("glove: unknown dtype %d", dtype);
    return CYMF_EINVAL;
}

extern "C" int cymf_glove_replay_epoch_dev(const cymf_glove_params *p, const int32_t *central, const int32_t *context,
                                           const double *counts, int64_t N, int32_t K, int32_t ld,
                                           double learning_rate, double x_max, double alpha, double *loss,
                                           void *stream) {
    CYMF_REQUIRE(glove_params_ok(p) && central && context && counts, "null pointer");
    CYMF_REQUIRE(K > 0 && ld >= K && K <= GLOVE_REPLAY_MAX_K, "bad shape (replay supports num_components <= 1024)");
    if (N <= 0) return 0;
    GloveArgs<double> a{};
    fill_args(a, p, central, context, counts, N, K, ld, learning_rate, x_max, alpha);
    glove_replay_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(a, loss);
    CYMF_LAUNCHED();
    return 0;
}

extern "C" int cymf_glove_fit_host(const int32_t *central, const int32_t *context, const double *counts, int64_t N,
                                   double *central_W, double *central_bias, double *context_W, double *context_bias,
                                   int64_t Vw, int64_t Vh, int32_t K, int32_t num_epochs,
                                   double learning_rate, double x_max, double alpha, int mode, double *loss_out) {
    CYMF_REQUIRE(central && context && counts && central_W && central_bias && context_W && context_bias, "null pointer");
    CYMF_REQUIRE(Vw > 0 && Vh > 0 && K > 0 && N >= 0 && num_epochs >= 0, "bad shape");
    CYMF_REQUIRE(mode >= 0 && mode <= 2, "mode must be 0 (hogwild f32), 1 (hogwild f64) or 2 (replay f64)");
    int ndev = 0;
    CYMF_CUDA(cudaGetDeviceCount(&ndev));
    if (ndev < 1) { set_error("no CUDA device: cymf_b200 has no CPU fallback"); return CYMF_EUNSUPPORTED; }
    const int dtype = mode == 0 ? CYMF_F32 : CYMF_F64;
    const size_t es = dtype == CYMF_F32 ? 4 : 8;
    const int32_t ld = (K + 3) / 4 * 4;
    DeviceArena mem;
    cudaStream_t st = nullptr;
    const int64_t vmax = Vw > Vh ? Vw : Vh;
    int64_t big = vmax * K;
    if (N > big) big = N;
    double *stage = nullptr, *d_loss = nullptr;
    CYMF_TRY(mem.get(&stage, (size_t)big * 8));
    cymf_glove_params p{};
    struct Item { void **dev; double *host; int64_t rows; int32_t k, l; };
    Item items[4] = {{&p.W, central_W, Vw, K, ld}, {&p.H, context_W, Vh, K, ld},
                     {&p.bW, central_bias, Vw, 1, 1}, {&p.bH, context_bias, Vh, 1, 1}};
    for (Item &it : items) {
        CYMF_TRY(mem.get((char **)it.dev, (size_t)it.rows * it.l * es));
        CYMF_CUDA(cudaMemcpyAsync(stage, it.host, (size_t)it.rows * it.k * 8, cudaMemcpyHostToDevice, st));
        CYMF_TRY(cymf_pack_rows_dev(stage, *it.dev, dtype, it.rows, it.k, it.l, st));
    }
    struct Acc { void **dev; int64_t n; };
    Acc accs[4] = {{&p.aW, Vw * ld}, {&p.aH, Vh * ld}, {&p.abW, Vw}, {&p.abH, Vh}};
    for (Acc &ac : accs) {                                       // optimizer.pyx:91-99: every accumulator starts at 1
        CYMF_TRY(mem.get((char **)ac.dev, (size_t)ac.n * es));
        CYMF_TRY(cymf_fill_dev(*ac.dev, dtype, ac.n, 1.0, st));
    }
    int32_t *d_c, *d_x;
    void *d_n;
    CYMF_TRY(mem.get(&d_c, (size_t)N * 4));
    CYMF_TRY(mem.get(&d_x, (size_t)N * 4));
    CYMF_TRY(mem.get((char **)&d_n, (size_t)N * es));
    CYMF_CUDA(cudaMemcpyAsync(d_c, central, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    CYMF_CUDA(cudaMemcpyAsync(d_x, context, (size_t)N * 4, cudaMemcpyHostToDevice, st));
    CYMF_CUDA(cudaMemcpyAsync(stage, counts, (size_t)N * 8, cudaMemcpyHostToDevice, st));
    if (N > 0) CYMF_TRY(cymf_pack_rows_dev(stage, d_n, dtype, N, 1, 1, st));
    CYMF_TRY(mem.get(&d_loss, 8 * (size_t)(num_epochs > 0 ? num_epochs : 1)));
    CYMF_CUDA(cudaMemsetAsync(d_loss, 0, 8 * (size_t)(num_epochs > 0 ? num_epochs : 1), st));
    const int64_t inflight = N / 256 > 1024 ? N / 256 : 1024;
    for (int32_t epoch = 0; epoch < num_epochs; ++epoch) {
        if (mode == 2) {
            // per-sample losses are not kept in this entry point; the mean is only a progress readout
            CYMF_TRY(cymf_glove_replay_epoch_dev(&p, d_c, d_x, (const double *)d_n, N, K, ld, learning_rate, x_max,
                                                 alpha, nullptr, st));
        } else {
            CYMF_TRY(cymf_glove_hogwild_epoch_dev(&p, dtype, dtype == CYMF_F32, d_c, d_x, d_n, N, K, ld, learning_rate,
                                                  x_max, alpha, inflight, loss_out ? d_loss + epoch : nullptr, st));
        }
    }
    for (Item &it : items) {
        CYMF_TRY(cymf_unpack_rows_dev(*it.dev, stage, dtype, it.rows, it.k, it.l, st));
        CYMF_CUDA(cudaMemcpyAsync(it.host, stage, (size_t)it.rows * it.k * 8, cudaMemcpyDeviceToHost, st));
    }
    if (loss_out && num_epochs > 0) {
        CYMF_CUDA(cudaMemcpyAsync(loss_out, d_loss, 8 * (size_t)num_epochs, cudaMemcpyDeviceToHost, st));
        CYMF_CUDA(cudaStreamSynchronize(st));
        for (int32_t e = 0; e < num_epochs; ++e) loss_out[e] = N > 0 ? loss_out[e] / (double)N : 0.0;
    }
    CYMF_CUDA(cudaStreamSynchronize(st));
    return 0;
}
