// bpr.cu -- BPR triplet update kernels (replaces the prange loop of cymf/bpr.pyx:160-171 together with
// BprModel.forward/backward, cymf/model.pyx:47-87, and the Sgd/AdaGrad/Adam element updates,
// cymf/optimizer.pyx:52-58,74-82,150-160).
//
//   bpr_hogwild_kernel : throughput path.  One lane group (4..32 lanes, one 16/32-byte slot per lane and
//                        pass) per triplet; Philox negative; (LPT+1)-ary membership probe of the user's CSR
//                        row overlapped with the three speculative row gathers; shuffle-reduced dot; update
//                        scattered with vector stores or vector reductions.  HBM/L2-bound gather-scatter:
//                        algorithmic bytes per applied SGD update = 6*K*sizeof(T) + 8.
//   bpr_replay_kernel  : parity path.  One warp applies the reference's own (u, i, j) stream strictly in
//                        order, f64, sequential-k dot, no FMA contraction (num_threads = 1 semantics).
#include <math.h>

#include "common.cuh"
#include "exp_table.h"
#include "optim.cuh"

namespace cymf {

template <typename T> struct BprArgs {
    T *W, *H, *s1W, *s1H, *s2W, *s2H;
    const int32_t *users, *positives, *negatives;
    const int64_t *indptr;
    const int32_t *indices;
    int64_t N, groups;
    int32_t I, ld;
    T lr, wd;
    uint64_t seed;
    uint32_t epoch;
    unsigned long long *applied;
    int64_t first, end;           // hogwild: the launch covers pairs [first, end) of the epoch; `users` / `positives` are
                                  // pre-shifted by -first so that both they and the Philox counter are indexed by l
};

// RANGE = false: the launch covers the whole epoch (first == 0) -- the code the throughput numbers are measured on.
// RANGE = true : pairs [first, end) of the epoch (an epoch issued in pieces while later pieces cross PCIe).
template <typename T, int OPT, int LPT, int NV, bool RED, bool RANGE>
__global__ void __launch_bounds__(256) bpr_hogwild_kernel(const BprArgs<T> a) {
    constexpr int GPW = 32 / LPT;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPT;
    const int gshift = (lane / LPT) * LPT;
    const unsigned gmask = group_mask<LPT>(lane);
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t stride = a.groups;                    // active lane groups in the grid (<= launched groups)
    unsigned long long n_applied = 0;

    const int64_t gid = warp * GPW + lane / LPT;
    const bool active = gid < stride;
#define CYMF_END (RANGE ? a.end : a.N)
    int64_t l = gid;
    if constexpr (RANGE) l += a.first;
    int32_t u_next = 0, i_next = 0;
    if (active && l < CYMF_END) { u_next = __ldcs(a.users + l); i_next = __ldcs(a.positives + l); }

    // warp-uniform trip count: the first group of the warp runs out last
    for (int64_t base = warp * GPW; base < a.N && warp * GPW < stride; base += stride, l += stride) {
        const bool valid = active && l < CYMF_END;
        const int32_t u = u_next, i = i_next;
        const int64_t ln = l + stride;
        if (active && ln < CYMF_END) { u_next = __ldcs(a.users + ln); i_next = __ldcs(a.positives + ln); }
        const int32_t j = (int32_t)philox_negative(a.seed, a.epoch, (uint64_t)l, (uint32_t)a.I);   // bpr.pyx:165

        // speculative gathers of the three rows (j is rarely a positive of u), issued before the membership probe
        T *pw = a.W + (size_t)u * a.ld, *pi = a.H + (size_t)i * a.ld, *pj = a.H + (size_t)j * a.ld;
        Slot<T> w[NV], hi[NV], hj[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
            const int e = (sub + v * LPT) * 4;
            if (e < a.ld) { w[v] = load_slot(pw + e); hi[v] = load_slot(pi + e); hj[v] = load_slot(pj + e); }
            else { w[v] = zero_slot<T>(); hi[v] = zero_slot<T>(); hj[v] = zero_slot<T>(); }
        }
        int64_t lo = 0, hi_ptr = 0;
        if (valid) { lo = __ldg(a.indptr + u); hi_ptr = __ldg(a.indptr + u + 1); }
        const bool hit = group_contains<LPT>(a.indices, lo, hi_ptr, j, sub, gmask, gshift);          // bpr.pyx:166

        T x = T(0);                                                                                    // model.pyx:55-56
#pragma unroll
        for (int v = 0; v < NV; ++v)
#pragma unroll
            for (int e = 0; e < 4; ++e) x += w[v].v[e] * (hi[v].v[e] - hj[v].v[e]);
        x = group_sum<LPT>(x, gmask);
        const T s = sigmoid_neg(x);                                                                    // model.pyx:78
        const bool apply = valid && !hit;
        if (apply) {
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int e0 = (sub + v * LPT) * 4;
                if (e0 >= a.ld) continue;
                Slot<T> aw = zero_slot<T>(), ai = aw, aj = aw, bw = aw, bi = aw, bj = aw;     // optimizer state slots
                if (OPT != CYMF_SGD) {
                    aw = load_slot(a.s1W + (size_t)u * a.ld + e0);
                    ai = load_slot(a.s1H + (size_t)i * a.ld + e0);
                    aj = load_slot(a.s1H + (size_t)j * a.ld + e0);
                }
                if (OPT == CYMF_ADAM) {
                    bw = load_slot(a.s2W + (size_t)u * a.ld + e0);
                    bi = load_slot(a.s2H + (size_t)i * a.ld + e0);
                    bj = load_slot(a.s2H + (size_t)j * a.ld + e0);
                }
                Slot<T> dw, di, dj;
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const T wk = w[v].v[e], hik = hi[v].v[e], hjk = hj[v].v[e];
                    const T gw = -(s * (hik - hjk) - a.wd * wk);                                      // model.pyx:81
                    const T gi = -(s * wk - a.wd * hik);                                              // model.pyx:82
                    const T gj = -(s * (-wk) - a.wd * hjk);                                           // model.pyx:83
                    dw.v[e] = opt_step<T, OPT, RED>(gw, a.lr, aw.v[e], bw.v[e]);
                    di.v[e] = opt_step<T, OPT, RED>(gi, a.lr, ai.v[e], bi.v[e]);
                    dj.v[e] = opt_step<T, OPT, RED>(gj, a.lr, aj.v[e], bj.v[e]);
                    if (!RED) { dw.v[e] += wk; di.v[e] += hik; dj.v[e] += hjk; }
                }
                if (RED) {
                    red_add_slot(pw + e0, dw); red_add_slot(pi + e0, di); red_add_slot(pj + e0, dj);
                    if (OPT != CYMF_SGD) {
                        red_add_slot(a.s1W + (size_t)u * a.ld + e0, aw);
                        red_add_slot(a.s1H + (size_t)i * a.ld + e0, ai);
                        red_add_slot(a.s1H + (size_t)j * a.ld + e0, aj);
                    }
                    if (OPT == CYMF_ADAM) {
                        red_add_slot(a.s2W + (size_t)u * a.ld + e0, bw);
                        red_add_slot(a.s2H + (size_t)i * a.ld + e0, bi);
                        red_add_slot(a.s2H + (size_t)j * a.ld + e0, bj);
                    }
                } else {
                    store_slot(pw + e0, dw); store_slot(pi + e0, di); store_slot(pj + e0, dj);
                    if (OPT != CYMF_SGD) {
                        store_slot(a.s1W + (size_t)u * a.ld + e0, aw);
                        store_slot(a.s1H + (size_t)i * a.ld + e0, ai);
                        store_slot(a.s1H + (size_t)j * a.ld + e0, aj);
                    }
                    if (OPT == CYMF_ADAM) {
                        store_slot(a.s2W + (size_t)u * a.ld + e0, bw);
                        store_slot(a.s2H + (size_t)i * a.ld + e0, bi);
                        store_slot(a.s2H + (size_t)j * a.ld + e0, bj);
                    }
                }
            }
            if (sub == 0) ++n_applied;
        }
    }
    if (a.applied) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) n_applied += __shfl_xor_sync(0xffffffffu, n_applied, off);
        if (lane == 0 && n_applied) atomicAdd(a.applied, n_applied);
    }
#undef CYMF_END
}

// ---- serialized f64 replay -------------------------------------------------------------------------------------
constexpr int REPLAY_MAX_K = 1024;

// exp(x) as the reference's libm evaluates it (cymf/model.pyx:78 calls libm exp; glibc >= 2.28 on an FMA host):
// x = k ln2/128 + r, exp(x) = 2^(k/128) (1 + tail + r + r^2 P(r)), every a*b+c fused exactly where the C
// library's FMA build fuses it.  Checked bit-for-bit against glibc 2.39 exp() on 2e7 arguments on the host
// (DESIGN.md "replay exactness"); CUDA's own exp() agrees with libm only to 1 ulp.  Outside the table-driven
// range (|x| >= 512, never reached by a BPR score) it defers to CUDA's exp().
__device__ __forceinline__ double exp_libm(double x) {
    const double ax = fabs(x);
    if (ax < 0x1p-54) return __dadd_rn(1.0, x);
    if (!(ax < 512.0)) return exp(x);
    const double inv_ln2_n = 0x1.71547652b82fep0 * 128, neg_ln2_hi_n = -0x1.62e42fefa0000p-8,
                 neg_ln2_lo_n = -0x1.cf79abc9e3b3ap-47, shift = 0x1.8p52;
    const double c2 = 0x1.ffffffffffdbdp-2, c3 = 0x1.555555555543cp-3, c4 = 0x1.55555cf172b91p-5,
                 c5 = 0x1.1111167a4d017p-7;
    double kd = __dadd_rn(__dmul_rn(inv_ln2_n, x), shift);
    const unsigned long long ki = (unsigned long long)__double_as_longlong(kd);
    kd = __dsub_rn(kd, shift);
    const double r = __fma_rn(kd, neg_ln2_lo_n, __fma_rn(kd, neg_ln2_hi_n, x));
    const unsigned idx = 2u * (unsigned)(ki & 127ull);
    const double tail = __longlong_as_double((long long)cymf_exp_table[idx]);
    const double scale = __longlong_as_double((long long)(cymf_exp_table[idx + 1] + (ki << 45)));
    const double r2 = __dmul_rn(r, r);
    const double p_lo = __fma_rn(r, c3, c2), p_hi = __fma_rn(r, c5, c4);
    const double t = __fma_rn(__dmul_rn(r2, r2), p_hi, __fma_rn(r2, p_lo, __dadd_rn(tail, r)));
    return __fma_rn(scale, t, scale);
}

template <int OPT>
__global__ void __launch_bounds__(32) bpr_replay_kernel(const BprArgs<double> a, int32_t K) {
    __shared__ double prod[REPLAY_MAX_K];
    const int lane = threadIdx.x;
    unsigned long long n_applied = 0;
    int32_t u_next = 0, i_next = 0, j_next = 0;
    if (a.N > 0) { u_next = a.users[0]; i_next = a.positives[0]; j_next = a.negatives[0]; }
    for (int64_t l = 0; l < a.N; ++l) {
        const int32_t u = u_next, i = i_next, j = j_next;
        if (l + 1 < a.N) { u_next = a.users[l + 1]; i_next = a.positives[l + 1]; j_next = a.negatives[l + 1]; }
        const bool hit = group_contains<32>(a.indices, a.indptr[u], a.indptr[u + 1], j, lane, 0xffffffffu, 0);
        if (hit) continue;                                                    // bpr.pyx:166-167
        double *pw = a.W + (size_t)u * a.ld, *pi = a.H + (size_t)i * a.ld, *pj = a.H + (size_t)j * a.ld;
        for (int k = lane; k < K; k += 32)
            prod[k] = __dmul_rn(__ldcg(pw + k), __dsub_rn(__ldcg(pi + k), __ldcg(pj + k)));
        __syncwarp();
        double x = 0.0;                                                       // model.pyx:53-56, k ascending
        for (int k = 0; k < K; ++k) x = __dadd_rn(x, prod[k]);
        __syncwarp();
        const double s = __ddiv_rn(1.0, __dadd_rn(1.0, exp_libm(x)));         // model.pyx:78
        for (int k = lane; k < K; k += 32) {
            const double wk = __ldcg(pw + k), hik = __ldcg(pi + k), hjk = __ldcg(pj + k);
            // model.pyx:81-83, every product and sum rounded separately as on the reference's x86-64 build
            const double gw = -__dsub_rn(__dmul_rn(s, __dsub_rn(hik, hjk)), __dmul_rn(a.wd, wk));
            const double gi = -__dsub_rn(__dmul_rn(s, wk), __dmul_rn(a.wd, hik));
            const double gj = -__dsub_rn(__dmul_rn(s, -wk), __dmul_rn(a.wd, hjk));
            const double g[3] = {gw, gi, gj};
            double *th[3] = {pw + k, pi + k, pj + k};
            const size_t off[3] = {(size_t)u * a.ld + k, (size_t)i * a.ld + k, (size_t)j * a.ld + k};
            const double cur[3] = {wk, hik, hjk};
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                double *S1 = r == 0 ? a.s1W : a.s1H, *S2 = r == 0 ? a.s2W : a.s2H;
                double next;
                if (OPT == CYMF_SGD) {
                    next = __dsub_rn(cur[r], __dmul_rn(a.lr, g[r]));
                } else if (OPT == CYMF_ADAGRAD) {
                    const double acc = __dadd_rn(S1[off[r]], __dmul_rn(g[r], g[r]));
                    S1[off[r]] = acc;
                    next = __dsub_rn(cur[r], __ddiv_rn(__dmul_rn(a.lr, g[r]), __dsqrt_rn(acc)));
                } else {
                    const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
                    const double m = __dadd_rn(__dmul_rn(b1, S1[off[r]]), __dmul_rn(1 - b1, g[r]));
                    const double v = __dadd_rn(__dmul_rn(b2, S2[off[r]]), __dmul_rn(1 - b2, __dmul_rn(g[r], g[r])));
                    S1[off[r]] = m;
                    S2[off[r]] = v;
                    const double num = __dmul_rn(a.lr, __ddiv_rn(m, 1 - b1));
                    const double den = __dadd_rn(__dsqrt_rn(__ddiv_rn(v, 1 - b2)), eps);
                    next = __dsub_rn(cur[r], __ddiv_rn(num, den));
                }
                __stcg(th[r], next);
            }
        }
        ++n_applied;
    }
    if (a.applied && lane == 0) atomicAdd(a.applied, n_applied);
}

// ---- launch plumbing -----------------------------------------------------------------------------------------
template <typename T, int OPT, int LPT, int NV, bool RED>
static int launch_hogwild(const BprArgs<T> &a, int64_t max_groups, cudaStream_t st) {
    auto kern = a.first > 0 ? bpr_hogwild_kernel<T, OPT, LPT, NV, RED, true> : bpr_hogwild_kernel<T, OPT, LPT, NV, RED, false>;
    int per_sm = 0;
    CYMF_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, 0));
    if (per_sm < 1) per_sm = 1;
    constexpr int GPB = 8 * (32 / LPT);                       // groups per 256-thread block
    int64_t blocks = (int64_t)sm_count() * per_sm;
    const int64_t need = (a.N + GPB - 1) / GPB;
    if (blocks > need) blocks = need;
    if (max_groups > 0 && blocks * GPB > max_groups) blocks = (max_groups + GPB - 1) / GPB;
    if (blocks < 1) blocks = 1;
    BprArgs<T> b = a;
    b.groups = blocks * GPB;
    if (max_groups > 0 && b.groups > max_groups) b.groups = max_groups;
    b.end = a.first + a.N;
    b.users = a.users - a.first;              // never dereferenced below index `first`
    b.positives = a.positives - a.first;
    kern<<<(unsigned)blocks, 256, 0, st>>>(b);
    CYMF_LAUNCHED();
    return 0;
}

template <typename T, int OPT, bool RED>
static int dispatch_shape(const BprArgs<T> &a, int64_t max_groups, cudaStream_t st) {
    const int slots = a.ld / 4;
    if (slots <= 4)   return launch_hogwild<T, OPT, 4, 1, RED>(a, max_groups, st);
    if (slots <= 8)   return launch_hogwild<T, OPT, 8, 1, RED>(a, max_groups, st);
    if (slots <= 16)  return launch_hogwild<T, OPT, 16, 1, RED>(a, max_groups, st);
    if (slots <= 32)  return launch_hogwild<T, OPT, 32, 1, RED>(a, max_groups, st);
    if (slots <= 64)  return launch_hogwild<T, OPT, 32, 2, RED>(a, max_groups, st);
    if (slots <= 96)  return launch_hogwild<T, OPT, 32, 3, RED>(a, max_groups, st);
    if (slots <= 128) return launch_hogwild<T, OPT, 32, 4, RED>(a, max_groups, st);
    set_error("bpr: num_components > 512 is not supported (ld=%d)", a.ld);
    return CYMF_EUNSUPPORTED;
}

template <typename T>
static int dispatch_hogwild(const cymf_factors *f, int optimizer, int scatter, BprArgs<T> a, int64_t max_groups,
                            cudaStream_t st) {
    a.W = (T *)f->W; a.H = (T *)f->H;
    a.s1W = (T *)f->s1W; a.s1H = (T *)f->s1H; a.s2W = (T *)f->s2W; a.s2H = (T *)f->s2H;
    switch (optimizer) {
        case CYMF_SGD:
            return scatter ? dispatch_shape<T, CYMF_SGD, true>(a, max_groups, st)
                           : dispatch_shape<T, CYMF_SGD, false>(a, max_groups, st);
        case CYMF_ADAGRAD:
            CYMF_REQUIRE(f->s1W && f->s1H, "adagrad needs s1W/s1H");
            return scatter ? dispatch_shape<T, CYMF_ADAGRAD, true>(a, max_groups, st)
                           : dispatch_shape<T, CYMF_ADAGRAD, false>(a, max_groups, st);
        case CYMF_ADAM:
            CYMF_REQUIRE(f->s1W && f->s1H && f->s2W && f->s2H, "adam needs s1*/s2*");
            return scatter ? dispatch_shape<T, CYMF_ADAM, true>(a, max_groups, st)
                           : dispatch_shape<T, CYMF_ADAM, false>(a, max_groups, st);
    }
    set_error("bpr: unknown optimizer %d", optimizer);
    return CYMF_EINVAL;
}

}  // namespace cymf

using namespace cymf;

extern "C" int cymf_bpr_hogwild_range_dev(const cymf_factors *f, int dtype, int optimizer, int scatter,
                                          const int32_t *users, const int32_t *positives, int64_t N,
                                          const int64_t *indptr, const int32_t *indices,
                                          int32_t U, int32_t I, int32_t K, int32_t ld,
                                          double learning_rate, double weight_decay,
                                          uint64_t seed, uint32_t epoch, int64_t max_inflight,
                                          unsigned long long *applied, int64_t first, void *stream) {
    CYMF_REQUIRE(f && f->W && f->H && users && positives && indptr && indices, "null pointer");
    CYMF_REQUIRE(U > 0 && I > 0 && K > 0 && ld >= K && ld % 4 == 0 && first >= 0, "bad shape (ld must be a multiple of 4, >= K)");
    if (N <= 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == CYMF_F32) {
        BprArgs<float> a{};
        a.users = users; a.positives = positives; a.indptr = indptr; a.indices = indices;
        a.N = N; a.I = I; a.ld = ld; a.lr = (float)learning_rate; a.wd = (float)weight_decay;
        a.seed = seed; a.epoch = epoch; a.applied = applied; a.first = first;
        return dispatch_hogwild<float>(f, optimizer, scatter, a, max_inflight, st);
    } else if (dtype == CYMF_F64) {
        BprArgs<double> a{};
        a.users = users; a.positives = positives; a.indptr = indptr; a.indices = indices;
        a.N = N; a.I = I; a.ld = ld; a.lr = learning_rate; a.wd = weight_decay;
        a.seed = seed; a.epoch = epoch; a.applied = applied; a.first = first;
        return dispatch_hogwild<double>(f, optimizer, scatter, a, max_inflight, st);
    }
    set_error("bpr: unknown dtype %d", dtype);
    return CYMF_EINVAL;
}

extern "C" int cymf_bpr_hogwild_epoch_dev(const cymf_factors *f, int dtype, int optimizer, int scatter,
                                          const int32_t *users, const int32_t *positives, int64_t N,
                                          const int64_t *indptr, const int32_t *indices,
                                          int32_t U, int32_t I, int32_t K, int32_t ld,
                                          double learning_rate, double weight_decay,
                                          uint64_t seed, uint32_t epoch, int64_t max_inflight,
                                          unsigned long long *applied, void *stream) {
    return cymf_bpr_hogwild_range_dev(f, dtype, optimizer, scatter, users, positives, N, indptr, indices, U, I, K, ld,
                                      learning_rate, weight_decay, seed, epoch, max_inflight, applied, 0, stream);
}

extern "C" int cymf_bpr_replay_epoch_dev(const cymf_factors *f, int optimizer,
                                         const int32_t *users, const int32_t *positives, const int32_t *negatives,
                                         int64_t N, const int64_t *indptr, const int32_t *indices,
                                         int32_t U, int32_t I, int32_t K, int32_t ld,
                                         double learning_rate, double weight_decay,
                                         unsigned long long *applied, void *stream) {
    CYMF_REQUIRE(f && f->W && f->H && users && positives && negatives && indptr && indices, "null pointer");
    CYMF_REQUIRE(U > 0 && I > 0 && K > 0 && ld >= K, "bad shape");
    CYMF_REQUIRE(K <= REPLAY_MAX_K, "replay supports num_components <= 1024");
    if (N <= 0) return 0;
    BprArgs<double> a{};
    a.W = (double *)f->W; a.H = (double *)f->H;
    a.s1W = (double *)f->s1W; a.s1H = (double *)f->s1H; a.s2W = (double *)f->s2W; a.s2H = (double *)f->s2H;
    a.users = users; a.positives = positives; a.negatives = negatives; a.indptr = indptr; a.indices = indices;
    a.N = N; a.I = I; a.ld = ld; a.lr = learning_rate; a.wd = weight_decay; a.applied = applied;
    cudaStream_t st = (cudaStream_t)stream;
    switch (optimizer) {
        case CYMF_SGD: bpr_replay_kernel<CYMF_SGD><<<1, 32, 0, st>>>(a, K); break;
        case CYMF_ADAGRAD:
            CYMF_REQUIRE(f->s1W && f->s1H, "adagrad needs s1W/s1H");
            bpr_replay_kernel<CYMF_ADAGRAD><<<1, 32, 0, st>>>(a, K); break;
        case CYMF_ADAM:
            CYMF_REQUIRE(f->s1W && f->s1H && f->s2W && f->s2H, "adam needs s1*/s2*");
            bpr_replay_kernel<CYMF_ADAM><<<1, 32, 0, st>>>(a, K); break;
        default: set_error("bpr: unknown optimizer %d", optimizer); return CYMF_EINVAL;
    }
    CYMF_LAUNCHED();
    return 0;
}
