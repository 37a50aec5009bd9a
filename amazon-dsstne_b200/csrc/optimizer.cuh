// optimizer.cuh -- per-element optimizer rules (hot-path row a12), shared by the streaming
// update kernels (optimizer.cu) and the fused sparse-gradient epilogue (sparse_wgrad.cu).
// Formulas exactly as coded in k{SGD,Momentum,AdaGrad,Nesterov,RMSProp,AdaDelta,Adam}Update*
// (E/kernels.cu:2746-3199); accurate sqrtf/powf/division (the reference builds with
// -use_fast_math; parity is judged against the accurate CPU oracle).
#pragma once
#include "common.cuh"

namespace dsb {

struct OptArgs {
    int   mode;
    float alpha, lambda, lambda1, mu, mu1, t;
    // Adam bias corrections, hoisted out of the per-element path: 1/(1-mu^(t+1)), 1/(1-mu1^(t+1))
    float adamC1, adamC2;
};

__host__ inline OptArgs make_opt(int mode, float alpha, float lambda, float lambda1, float mu, float mu1, float t)
{
    OptArgs o; o.mode = mode; o.alpha = alpha; o.lambda = lambda; o.lambda1 = lambda1; o.mu = mu; o.mu1 = mu1; o.t = t;
    o.adamC1 = 0.0f; o.adamC2 = 0.0f;
    return o;
}

// weight rule: g is the (already negative-scaled) gradient, w the weight; v / gv the state
// (ignored by modes that do not use them).  Returns the new weight.
template <int MODE>
__device__ __forceinline__ float opt_weight(const OptArgs& o, float g, float w, float& v, float& gv)
{
    if (MODE == DSB200_SGD) {                        // E/kernels.cu:2746-2757
        return w + o.alpha * (g - o.lambda * w - o.lambda1 * sgnf(w));
    } else if (MODE == DSB200_MOMENTUM) {            // :2798-2812
        v = o.mu * v + o.alpha * (g - o.lambda * w - o.lambda1 * sgnf(w));
        return w + v;
    } else if (MODE == DSB200_ADAGRAD) {             // :2854-2869
        g -= o.lambda * w + o.lambda1 * sgnf(w);
        v += g * g;
        return w + o.alpha * g * (1.0f / sqrtf(fmaxf(0.000000001f, v)));
    } else if (MODE == DSB200_NESTEROV) {            // :3047-3062
        const float vOld = v;
        const float vNew = o.mu * vOld + o.alpha * (g - o.lambda * w - o.lambda1 * sgnf(w));
        v = vNew;
        return w + vNew + o.mu * (vNew - vOld);
    } else if (MODE == DSB200_RMSPROP) {             // :3144-3159
        g -= o.lambda * w + o.lambda1 * sgnf(w);
        v = o.mu * v + (1.0f - o.mu) * g * g;
        return w + o.alpha * g * (1.0f / sqrtf(fmaxf(0.000000001f, v)));
    } else if (MODE == DSB200_ADADELTA) {            // :2911-2930
        g -= o.lambda * w + o.lambda1 * sgnf(w);
        gv = o.mu * gv + (1.0f - o.mu) * g * g;
        const float dw = sqrtf(fmaxf(0.000000001f, v) / fmaxf(0.000000001f, gv)) * g;
        v = o.mu * v + (1.0f - o.mu) * dw * dw;
        return w + dw;
    } else {                                         // Adam, :2976-2998 (t+1 inside the kernel)
        float dw = g - (o.lambda * w + o.lambda1 * sgnf(w));
        v  = o.mu * v + (1.0f - o.mu) * dw;
        gv = o.mu1 * gv + (1.0f - o.mu1) * dw * dw;
        const float tt = o.t + 1.0f;
        const float vdw = v / (1.0f - powf(o.mu, tt));
        const float sdw = gv / (1.0f - powf(o.mu1, tt));
        return w + o.alpha * vdw / (sqrtf(sdw) + 1.0e-8f);
    }
}

// bias rule: gbar = mean over the batch of delta[:,c]; note the sign (biases SUBTRACT)
template <int MODE>
__device__ __forceinline__ float opt_bias(const OptArgs& o, float gbar, float b, float& v, float& gv)
{
    if (MODE == DSB200_SGD) {                        // E/kernels.cu:2766-2788
        return b - o.alpha * gbar;
    } else if (MODE == DSB200_MOMENTUM) {            // :2821-2845
        v = o.mu * v - o.alpha * gbar;
        return b + v;
    } else if (MODE == DSB200_ADAGRAD) {             // :2878-2902
        v += gbar * gbar;
        return b - o.alpha * gbar * (1.0f / sqrtf(fmaxf(0.000000001f, v)));
    } else if (MODE == DSB200_NESTEROV) {            // :3071-3095
        const float vOld = v;
        const float vNew = o.mu * vOld - o.alpha * gbar;
        v = vNew;
        return b + vNew + o.mu * (vNew - vOld);
    } else if (MODE == DSB200_RMSPROP) {             // :3168-3192
        v = o.mu * v + (1.0f - o.mu) * gbar * gbar;
        return b - o.alpha * gbar * (1.0f / sqrtf(fmaxf(0.000000001f, v)));
    } else if (MODE == DSB200_ADADELTA) {            // :2939-2967
        gv = o.mu * gv + (1.0f - o.mu) * gbar * gbar;
        const float dw = sqrtf(fmaxf(0.000000001f, v) / fmaxf(0.000000001f, gv)) * gbar;
        v = o.mu * v + (1.0f - o.mu) * dw * dw;
        return b - dw;
    } else {                                         // Adam, :3007-3038
        v  = o.mu * v + (1.0f - o.mu) * gbar;
        gv = o.mu1 * gv + (1.0f - o.mu1) * gbar * gbar;
        const float tt = o.t + 1.0f;
        const float vdw = v / (1.0f - powf(o.mu, tt));
        const float sdw = gv / (1.0f - powf(o.mu1, tt));
        return b - o.alpha * vdw / (sqrtf(sdw) + 1.0e-8f);
    }
}

constexpr bool opt_uses_v(int mode)  { return mode != DSB200_SGD; }
constexpr bool opt_uses_gv(int mode) { return mode == DSB200_ADADELTA || mode == DSB200_ADAM; }

}  // namespace dsb
