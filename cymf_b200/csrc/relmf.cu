// relmf.cu -- RelMF pointwise update kernels (replaces the prange loop of cymf/relmf.pyx:143-148 together with
// RelMfModel.forward/backward, cymf/model.pyx:99-142, and the Sgd/AdaGrad/Adam element updates of
// cymf/optimizer.pyx).  Same gather/scatter skeleton as bpr.cu with two rows per sample instead of three.
//
//   relmf_hogwild_kernel : one lane group per sampled cell (u, i) ~ U[0,U) x U[0,I) (Philox4x32-10 keyed by
//                          (seed, epoch, l)); the label X[u,i] comes from an (LPT+1)-ary search of the user's
//                          CSR row (the reference densifies X, relmf.pyx:79-80: 29.6 GB at the ml-20m shape);
//                          the two row gathers are issued before the search.  With scatter = 1 the parameter
//                          step AND the optimizer-state increments (AdaGrad g^2; Adam's m, v deltas) are applied
//                          with red.global.add, so concurrent samples of one row all count and (m, v) can never
//                          be torn -- plain stores let Adam diverge once a row has several samples in flight
//                          (measured on the ml-100k shape, tools/relmf_probe.py).  HBM/L2-bound:
//                          algorithmic bytes per SGD sample = 4*K*sizeof(T) + sizeof(T) (two rows RW + p_i).
//   relmf_replay_kernel  : parity path.  One warp applies the reference's own mt19937 cell stream in order,
//                          f64, sequential-k dot, no FMA contraction (num_threads = 1 semantics).
#include "common.cuh"
#include "optim.cuh"

namespace cymf {

template <typename T> struct RelmfArgs {
    T *W, *H, *s1W, *s1H, *s2W, *s2H;
    const int64_t *indptr;
    const int32_t *indices;
    const T *values;              // X.data in CSR order, or NULL when every stored cell is 1
    const T *prop;                // propensities [I]
    const int64_t *cells;         // replay only: r = u * I + i per sample
    int64_t n, groups;
    int32_t U, I, ld;
    T lr, wd, clip;
    uint64_t seed;
    uint32_t epoch;
};

// Two-word bounded draw: multiply-shift; a first word that falls in the biased zone is replaced by the second.
__host__ __device__ __forceinline__ uint32_t bounded_from_pair(uint32_t a, uint32_t b, uint32_t n) {
    uint64_t prod = (uint64_t)a * n;
    const uint32_t low = (uint32_t)prod;
    if (low < n && low < (0u - n) % n) prod = (uint64_t)b * n;
    return (uint32_t)(prod >> 32);
}

// The cell sample l of `epoch` touches.  u and i are drawn independently, which is the same distribution as the
// reference's r ~ U[0, U*I), u = r / I, i = r % I (relmf.pyx:144-146).
__host__ __device__ __forceinline__ void philox_cell(uint64_t seed, uint32_t epoch, uint64_t l, uint32_t U, uint32_t I,
                                                     int32_t *u, int32_t *i) {
    const uint4 r = philox4x32_10(make_uint4((uint32_t)l, (uint32_t)(l >> 32), epoch, 0x524d4631u /* "RMF1" */),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    *u = (int32_t)bounded_from_pair(r.x, r.y, U);
    *i = (int32_t)bounded_from_pair(r.z, r.w, I);
}

template <typename T, int OPT, int LPT, int NV, bool RED>
__global__ void __launch_bounds__(256) relmf_hogwild_kernel(const RelmfArgs<T> a) {
    constexpr int GPW = 32 / LPT;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPT;
    const int gshift = (lane / LPT) * LPT;
    const unsigned gmask = group_mask<LPT>(lane);
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t stride = a.groups;
    const int64_t gid = warp * GPW + lane / LPT;
    const bool active = gid < stride;
    int64_t l = gid;

    for (int64_t base = warp * GPW; base < a.n && warp * GPW < stride; base += stride, l += stride) {
        const bool valid = active && l < a.n;
        int32_t u, i;
        philox_cell(a.seed, a.epoch, (uint64_t)l, (uint32_t)a.U, (uint32_t)a.I, &u, &i);            // relmf.pyx:144-146
        T *pw = a.W + (size_t)u * a.ld, *ph = a.H + (size_t)i * a.ld;
        Slot<T> w[NV], h[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int e = (sub + v * LPT) * 4;
            if (e < a.ld) { w[v] = load_slot(pw + e); h[v] = load_slot(ph + e); }
            else { w[v] = zero_slot<T>(); h[v] = zero_slot<T>(); }
        }
        const T p = __ldg(a.prop + i);
        int64_t lo = 0, hi = 0;
        if (valid) { lo = __ldg(a.indptr + u); hi = __ldg(a.indptr + u + 1); }
        const int64_t pos = group_find<LPT>(a.indices, lo, hi, i, sub, gmask, gshift);               // X[u, i]
        T x = T(0);
        if (pos >= 0) x = a.values ? __ldg(a.values + pos) : T(1);

        T t = T(0);                                                                                   // model.pyx:114-115
#pragma unroll
        for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int e = 0; e < 4; ++e) t += w[v].v[e] * h[v].v[e];
        t = group_sum<LPT>(t, gmask);
        const T c = x / (p >= a.clip ? p : a.clip);                                                   // r / dmax(p, M)
        const T coef = c * (T(1) - t) + (T(1) - c) * (T(0) - t);                                     // model.pyx:130-131
        if (valid) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int e0 = (sub + v * LPT) * 4;
                if (e0 >= a.ld) continue;
                Slot<T> aw = zero_slot<T>(), ah = aw, bw = aw, bh = aw;
                if (OPT != CYMF_SGD) {
                    aw = load_slot(a.s1W + (size_t)u * a.ld + e0);
                    ah = load_slot(a.s1H + (size_t)i * a.ld + e0);
                }
                if (OPT == CYMF_ADAM) {
                    bw = load_slot(a.s2W + (size_t)u * a.ld + e0);
                    bh = load_slot(a.s2H + (size_t)i * a.ld + e0);
                }
                Slot<T> dw, dh;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const T wk = w[v].v[e], hk = h[v].v[e];
                    const T gw = -(coef * hk) + a.wd * wk;                                            // model.pyx:129-132
                    const T gh = -(coef * wk) + a.wd * hk;                                            // model.pyx:134-137
                    dw.v[e] = opt_step<T, OPT, RED>(gw, a.lr, aw.v[e], bw.v[e]);
                    dh.v[e] = opt_step<T, OPT, RED>(gh, a.lr, ah.v[e], bh.v[e]);
                    if (!RED) { dw.v[e] += wk; dh.v[e] += hk; }
                }
                if (RED) {
                    red_add_slot(pw + e0, dw); red_add_slot(ph + e0, dh);
                    if (OPT != CYMF_SGD) {
                        red_add_slot(a.s1W + (size_t)u * a.ld + e0, aw);
                        red_add_slot(a.s1H + (size_t)i * a.ld + e0, ah);
                    }
                    if (OPT == CYMF_ADAM) {
                        red_add_slot(a.s2W + (size_t)u * a.ld + e0, bw);
                        red_add_slot(a.s2H + (size_t)i * a.ld + e0, bh);
                    }
                } else {
                    store_slot(pw + e0, dw); store_slot(ph + e0, dh);
                    if (OPT != CYMF_SGD) {
                        store_slot(a.s1W + (size_t)u * a.ld + e0, aw);
                        store_slot(a.s1H + (size_t)i * a.ld + e0, ah);
                    }
                    if (OPT == CYMF_ADAM) {
                        store_slot(a.s2W + (size_t)u * a.ld + e0, bw);
                        store_slot(a.s2H + (size_t)i * a.ld + e0, bh);
                    }
                }
            }
        }
    }
}

// ---- serialized f64 replay -------------------------------------------------------------------------------------
constexpr int RELMF_REPLAY_MAX_K = 1024;

template <int OPT>
__global__ void __launch_bounds__(32) relmf_replay_kernel(const RelmfArgs<double> a, int32_t K) {
    __shared__ double prod[RELMF_REPLAY_MAX_K];
    const int lane = threadIdx.x;
    int64_t r_next = a.n > 0 ? a.cells[0] : 0;
    for (int64_t l = 0; l < a.n; ++l) {
        const int64_t r = r_next;
        if (l + 1 < a.n) r_next = a.cells[l + 1];
        const int32_t u = (int32_t)(r / a.I), i = (int32_t)(r % a.I);                                // relmf.pyx:145-146
        const int64_t pos = group_find<32>(a.indices, a.indptr[u], a.indptr[u + 1], i, lane, 0xffffffffu, 0);
        const double x = pos < 0 ? 0.0 : (a.values ? a.values[pos] : 1.0);
        const double p = a.prop[i];
        const double pm = p >= a.clip ? p : a.clip;                                                  // math.pxd:47-51
        double *pw = a.W + (size_t)u * a.ld, *ph = a.H + (size_t)i * a.ld;
        for (int k = lane; k < K; k += 32) prod[k] = __dmul_rn(__ldcg(pw + k), __ldcg(ph + k));
        __syncwarp();
        double t = 0.0;                                                                              // model.pyx:114-115
        for (int k = 0; k < K; ++k) t = __dadd_rn(t, prod[k]);
        __syncwarp();
        const double c = __ddiv_rn(x, pm);
        const double pos_term = __dmul_rn(c, __dsub_rn(1.0, t));                                     // (r/pm) * (1. - t)
        const double neg_term = __dmul_rn(__dsub_rn(1.0, c), __dsub_rn(0.0, t));                     // (1 - r/pm) * (0. - t)
        for (int k = lane; k < K; k += 32) {
            const double wk = __ldcg(pw + k), hk = __ldcg(ph + k);
            const double gw = __dadd_rn(-__dadd_rn(__dmul_rn(pos_term, hk), __dmul_rn(neg_term, hk)), __dmul_rn(a.wd, wk));
            const double gh = __dadd_rn(-__dadd_rn(__dmul_rn(pos_term, wk), __dmul_rn(neg_term, wk)), __dmul_rn(a.wd, hk));
            const size_t ow = (size_t)u * a.ld + k, oh = (size_t)i * a.ld + k;
            __stcg(pw + k, opt_step_exact<OPT>(wk, gw, a.lr, a.s1W ? a.s1W + ow : nullptr, a.s2W ? a.s2W + ow : nullptr));
            __stcg(ph + k, opt_step_exact<OPT>(hk, gh, a.lr, a.s1H ? a.s1H + oh : nullptr, a.s2H ? a.s2H + oh : nullptr));
        }
        __syncwarp();
    }
}

// ---- launch plumbing -----------------------------------------------------------------------------------------
template <typename T, int OPT, int LPT, int NV, bool RED>
static int launch_relmf(const RelmfArgs<T> &a, int64_t max_groups, cudaStream_t st) {
    auto kern = relmf_hogwild_kernel<T, OPT, LPT, NV, RED>;
    int per_sm = 0;
    CYMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));
    if (per_sm < 1) per_sm = 1;
    constexpr int GPB = 8 * (32 / LPT);
    int64_t blocks = (int64_t)sm_count() * per_sm;
    const int64_t need = (a.n + GPB - 1) / GPB;
    if (blocks > need) blocks = need;
    if (max_groups > 0 && blocks * GPB > max_groups) blocks = (max_groups + GPB - 1) / GPB;
    if (blocks < 1) blocks = 1;
    RelmfArgs<T> b = a;
    b.groups = blocks * GPB;
    if (max_groups > 0 && b.groups > max_groups) b.groups = max_groups;
    kern<<<(unsigned)blocks, 256, 0, st>>>(b);
    CYMF_LAUNCHED();
    return 0;
}

template <typename T, int OPT, bool RED>
static int relmf_shape(const RelmfArgs<T> &a, int64_t max_groups, cudaStream_t st) {
    const int slots = a.ld / 4;
    if (slots <= 4)   return launch_relmf<T, OPT, 4, 1, RED>(a, max_groups, st);
    if (slots <= 8)   return launch_relmf<T, OPT, 8, 1, RED>(a, max_groups, st);
    if (slots <= 16)  return launch_relmf<T, OPT, 16, 1, RED>(a, max_groups, st);
    if (slots <= 32)  return launch_relmf<T, OPT, 32, 1, RED>(a, max_groups, st);
    if (slots <= 64)  return launch_relmf<T, OPT, 32, 2, RED>(a, max_groups, st);
    if (slots <= 96)  return launch_relmf<T, OPT, 32, 3, RED>(a, max_groups, st);
    if (slots <= 128) return launch_relmf<T, OPT, 32, 4, RED>(a, max_groups, st);
    set_error("relmf: num_components > 512 is not supported (ld=%d)", a.ld);
    return CYMF_EUNSUPPORTED;
}

template <typename T>
static int relmf_dispatch(const cymf_factors *f, int optimizer, int scatter, RelmfArgs<T> a, int64_t max_groups,
                          cudaStream_t st) {
    a.W = (T *)f->W; a.H = (T *)f->H;
    a.s1W = (T *)f->s1W; a.s1H = (T *)f->s1H; a.s2W = (T *)f->s2W; a.s2H = (T *)f->s2H;
    switch (optimizer) {
        case CYMF_SGD:
            return scatter ? relmf_shape<T, CYMF_SGD, true>(a, max_groups, st)
                           : relmf_shape<T, CYMF_SGD, false>(a, max_groups, st);
        case CYMF_ADAGRAD:
            CYMF_REQUIRE(f->s1W && f->s1H, "adagrad needs s1W/s1H");
            return scatter ? relmf_shape<T, CYMF_ADAGRAD, true>(a, max_groups, st)
                           : relmf_shape<T, CYMF_ADAGRAD, false>(a, max_groups, st);
        case CYMF_ADAM:
            CYMF_REQUIRE(f->s1W && f->s1H && f->s2W && f->s2H, "adam needs s1*/s2*");
            return scatter ? relmf_shape<T, CYMF_ADAM, true>(a, max_groups, st)
                           : relmf_shape<T, CYMF_ADAM, false>(a, max_groups, st);
    }
    set_error("relmf: unknown optimizer %d", optimizer);
    return CYMF_EINVAL;
}
}  // namespace cymf

using namespace cymf;

extern "C" int cymf_relmf_hogwild_epoch_dev(const cymf_factors *f, int dtype, int optimizer, int scatter,
                                            const int64_t *indptr, const int32_t *indices, const void *values,
                                            const void *propensities, int32_t U, int32_t I, int32_t K, int32_t ld,
                                            int64_t n_samples, double learning_rate, double weight_decay,
                                            double clip_value, uint64_t seed, uint32_t epoch, int64_t max_inflight,
                                            void *stream) {
    CYMF_REQUIRE(f && f->W && f->H && indptr && indices && propensities, "null pointer");
    CYMF_REQUIRE(U > 0 && I > 0 && K > 0 && ld >= K && ld % 4 == 0, "bad shape (ld must be a multiple of 4, >= K)");
    if (n_samples <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CYMF_F32) {
        RelmfArgs<float> a{};
        a.indptr = indptr; a.indices = indices; a.values = (const float *)values; a.prop = (const float *)propensities;
        a.n = n_samples; a.U = U; a.I = I; a.ld = ld;
        a.lr = (float)learning_rate; a.wd = (float)weight_decay; a.clip = (float)clip_value;
        a.seed = seed; a.epoch = epoch;
        return relmf_dispatch<float>(f, optimizer, scatter, a, max_inflight, st);
    } else if (dtype == CYMF_F64) {
        RelmfArgs<double> a{};
        a.indptr = indptr; a.indices = indices; a.values = (const double *)values; a.prop = (const double *)propensities;
        a.n = n_samples; a.U = U; a.I = I; a.ld = ld;
        a.lr = learning_rate; a.wd = weight_decay; a.clip = clip_value;
        a.seed = seed; a.epoch = epoch;
        return relmf_dispatch<double>(f, optimizer, scatter, a, max_inflight, st);
    }
    set_error("relmf: unknown dtype %d", dtype);
    return CYMF_EINVAL;
}

extern "C" int cymf_relmf_cells_host(uint64_t seed, uint32_t epoch, int64_t first, int64_t count, int32_t U, int32_t I,
                                     int64_t *out) {
    CYMF_REQUIRE(out && U > 0 && I > 0 && first >= 0 && count >= 0, "bad argument");
    for (int64_t t = 0; t < count; ++t) {
        int32_t u, i;
        philox_cell(seed, epoch, (uint64_t)(first + t), (uint32_t)U, (uint32_t)I, &u, &i);
        out[t] = (int64_t)u * I + i;
    }
    return 0;
}

extern "C" int cymf_relmf_replay_epoch_dev(const cymf_factors *f, int optimizer, const int64_t *cells, int64_t n_samples,
                                           const int64_t *indptr, const int32_t *indices, const double *values,
                                           const double *propensities, int32_t U, int32_t I, int32_t K, int32_t ld,
                                           double learning_rate, double weight_decay, double clip_value, void *stream) {
    CYMF_REQUIRE(f && f->W && f->H && cells && indptr && indices && propensities, "null pointer");
    CYMF_REQUIRE(U > 0 && I > 0 && K > 0 && ld >= K, "bad shape");
    CYMF_REQUIRE(K <= RELMF_REPLAY_MAX_K, "replay supports num_components <= 1024");
    if (n_samples <= 0) return 0;
    RelmfArgs<double> a{};
    a.W = (double *)f->W; a.H = (double *)f->H;
    a.s1W = (double *)f->s1W; a.s1H = (double *)f->s1H; a.s2W = (double *)f->s2W; a.s2H = (double *)f->s2H;
    a.indptr = indptr; a.indices = indices; a.values = values; a.prop = propensities; a.cells = cells;
    a.n = n_samples; a.U = U; a.I = I; a.ld = ld; a.lr = learning_rate; a.wd = weight_decay; a.clip = clip_value;
    cudaStream_t st = (cudaStream_t)stream;
    switch (optimizer) {
        case CYMF_SGD: relmf_replay_kernel<CYMF_SGD><<<1, 32, 0, st>>>(a, K); break;
        case CYMF_ADAGRAD:
            CYMF_REQUIRE(f->s1W && f->s1H, "adagrad needs s1W/s1H");
            relmf_replay_kernel<CYMF_ADAGRAD><<<1, 32, 0, st>>>(a, K); break;
        case CYMF_ADAM:
            CYMF_REQUIRE(f->s1W && f->s1H && f->s2W && f->s2H, "adam needs s1*/s2*");
            relmf_replay_kernel<CYMF_ADAM><<<1, 32, 0, st>>>(a, K); break;
        default: set_error("relmf: unknown optimizer %d", optimizer); return CYMF_EINVAL;
    }
    CYMF_LAUNCHED();
    return 0;
}
