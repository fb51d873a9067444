// optim.cuh -- element update rules of the reference's optimizers (cymf/optimizer.pyx), shared by the BPR and RelMF
// kernels.  All three keep the reference's quirks: AdaGrad accumulators start at 1 with no epsilon, Adam has no
// timestep and a constant 1/(1-beta) bias correction.
#pragma once
#include "common.cuh"

namespace cymf {

__device__ __forceinline__ float sigmoid_neg(float x) { return __frcp_rn(1.0f + __expf(x)); }
__device__ __forceinline__ double sigmoid_neg(double x) { return 1.0 / (1.0 + exp(x)); }
__device__ __forceinline__ float rsqrt_t(float x) { return rsqrtf(x); }
__device__ __forceinline__ double rsqrt_t(double x) { return 1.0 / sqrt(x); }
__device__ __forceinline__ float sqrt_t(float x) { return sqrtf(x); }
__device__ __forceinline__ double sqrt_t(double x) { return sqrt(x); }

// One element of the update rule; g is the reference's gradient, the return value the additive step d such that
// theta_new = theta + d.  Advances the optimizer state (s1: AdaGrad accumulator / Adam M, s2: Adam V) in place.
template <typename T, int OPT>
__device__ __forceinline__ T opt_step(T g, T lr, T &s1, T &s2) {
    if (OPT == CYMF_SGD) {
        return -lr * g;                                           // optimizer.pyx:52-58
    } else if (OPT == CYMF_ADAGRAD) {
        s1 += g * g;                                              // optimizer.pyx:74-82 (state starts at 1)
        return -lr * g * rsqrt_t(s1);
    } else {
        const T b1 = T(0.9), b2 = T(0.999), eps = T(1e-8);        // optimizer.pyx:127-160, no timestep
        s1 = b1 * s1 + (T(1) - b1) * g;
        s2 = b2 * s2 + (T(1) - b2) * g * g;
        return -lr * (s1 / (T(1) - b1)) / (sqrt_t(s2 / (T(1) - b2)) + eps);
    }
}

// The same rules with every operation individually IEEE-rounded in the reference's order (serialized f64 replay).
// Returns the new parameter value.
template <int OPT>
__device__ __forceinline__ double opt_step_exact(double theta, double g, double lr, double *s1, double *s2) {
    if (OPT == CYMF_SGD) {
        return __dsub_rn(theta, __dmul_rn(lr, g));
    } else if (OPT == CYMF_ADAGRAD) {
        const double acc = __dadd_rn(*s1, __dmul_rn(g, g));
        *s1 = acc;
        return __dsub_rn(theta, __ddiv_rn(__dmul_rn(lr, g), __dsqrt_rn(acc)));
    } else {
        const double b1 = 0.9, b2 = 0.999, eps = 1e-8;
        const double m = __dadd_rn(__dmul_rn(b1, *s1), __dmul_rn(1 - b1, g));
        const double v = __dadd_rn(__dmul_rn(b2, *s2), __dmul_rn(1 - b2, __dmul_rn(g, g)));
        *s1 = m;
        *s2 = v;
        const double num = __dmul_rn(lr, __ddiv_rn(m, 1 - b1));
        const double den = __dadd_rn(__dsqrt_rn(__ddiv_rn(v, 1 - b2)), eps);
        return __dsub_rn(theta, __ddiv_rn(num, den));
    }
}

}  // namespace cymf
