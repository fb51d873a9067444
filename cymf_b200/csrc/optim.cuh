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
__device__ __forceinline__ float fmin_t(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double fmin_t(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ float fmax_t(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double fmax_t(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float sqrt_t(float x) { return sqrtf(x); }
__device__ __forceinline__ double sqrt_t(double x) { return sqrt(x); }

// One element of the update rule; g is the reference's gradient, the return value the additive step d such that
// theta_new = theta + d.  DELTA = false: advances the optimizer state (s1: AdaGrad accumulator / Adam M, s2: Adam V)
// in place.  DELTA = true: s1 / s2 are replaced by their INCREMENTS (to be applied with red.add), so that
// concurrent samples of one row all count.
template <typename T, int OPT, bool DELTA = false>
__device__ __forceinline__ T opt_step(T g, T lr, T &s1, T &s2) {
    if (OPT == CYMF_SGD) {
        return -lr * g;                                           // optimizer.pyx:52-58
    } else if (OPT == CYMF_ADAGRAD) {
        const T acc = s1 + g * g;                                 // optimizer.pyx:74-82 (state starts at 1)
        s1 = DELTA ? g * g : acc;
        return -lr * g * rsqrt_t(acc);
    } else {
        const T b1 = T(0.9), b2 = T(0.999), eps = T(1e-8);        // optimizer.pyx:127-160, no timestep
        const T m = b1 * s1 + (T(1) - b1) * g;
        const T v = b2 * s2 + (T(1) - b2) * g * g;
        s1 = DELTA ? (T(1) - b1) * (g - s1) : m;
        s2 = DELTA ? (T(1) - b2) * (g * g - s2) : v;
        // In serial execution |m^ / sqrt(v^)| <= (1 - b1^2 / b2)^(-1/2) = 2.2991 (Cauchy-Schwarz on the two
        // exponential sums), so the clamp never binds there; under Hogwild it bounds the step a lane takes from a
        // torn (m, v) pair -- m already holding a concurrent sample's gradient that v does not hold yet -- which
        // otherwise reaches lr |g| / eps and sends rows to infinity.
        const T ratio = (m / (T(1) - b1)) / (sqrt_t(v / (T(1) - b2)) + eps);
        return -lr * fmin_t(fmax_t(ratio, T(-2.3)), T(2.3));
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
