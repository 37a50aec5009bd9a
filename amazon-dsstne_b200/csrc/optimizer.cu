// optimizer.cu -- fused optimizer step (hot-path row a12).
//
// Replaces k{SGD,Momentum,AdaGrad,Nesterov,RMSProp,AdaDelta,Adam}Update{Weights,Biases} and
// kCalculateRegularizationError (E/kernels.cu:2719-3199; dispatch E/NNWeight.cpp:718-851).
//  * weights: one streaming pass, 128-bit loads/stores, grid = multiple of the SM count;
//  * biases: the reference sums delta columns with ONE thread per column looping over the batch
//    (27k threads for a 27,278-wide layer, i.e. a fraction of one wave).  Here a CTA owns 32
//    columns x a slice of the batch rows, lanes read 128-byte row segments, the 8 warps and --
//    for narrow layers -- several row-slice CTAs are combined in a fixed order (last-arriving
//    CTA adds the slice partials), so the column mean is deterministic.
#include "optimizer.cuh"
#include "launch.h"

namespace dsb {

template <int MODE>
__global__ void __launch_bounds__(256)
update_weights_kernel(const OptArgs o, uint64_t size, float* __restrict__ v, const float* __restrict__ g,
                      float* __restrict__ gv, float* __restrict__ w, int vec)
{
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t nth = (uint64_t)gridDim.x * blockDim.x;
    pdl_launch_dependents();
    pdl_wait();
    if (vec) {
        const uint64_t n4 = size >> 2;
        for (uint64_t i = tid; i < n4; i += nth) {
            float4 g4 = ldg_cs_f4(reinterpret_cast<const float4*>(g) + i);
            float4 w4 = reinterpret_cast<float4*>(w)[i];
            float4 v4 = make_float4(0, 0, 0, 0), s4 = make_float4(0, 0, 0, 0);
            if (opt_uses_v(MODE))  v4 = reinterpret_cast<float4*>(v)[i];
            if (opt_uses_gv(MODE)) s4 = reinterpret_cast<float4*>(gv)[i];
            w4.x = opt_weight<MODE>(o, g4.x, w4.x, v4.x, s4.x);
            w4.y = opt_weight<MODE>(o, g4.y, w4.y, v4.y, s4.y);
            w4.z = opt_weight<MODE>(o, g4.z, w4.z, v4.z, s4.z);
            w4.w = opt_weight<MODE>(o, g4.w, w4.w, v4.w, s4.w);
            reinterpret_cast<float4*>(w)[i] = w4;
            if (opt_uses_v(MODE))  reinterpret_cast<float4*>(v)[i] = v4;
            if (opt_uses_gv(MODE)) reinterpret_cast<float4*>(gv)[i] = s4;
        }
        for (uint64_t i = (n4 << 2) + tid; i < size; i += nth) {
            float vv = opt_uses_v(MODE) ? v[i] : 0.0f, ss = opt_uses_gv(MODE) ? gv[i] : 0.0f;
            w[i] = opt_weight<MODE>(o, g[i], w[i], vv, ss);
            if (opt_uses_v(MODE)) v[i] = vv;
            if (opt_uses_gv(MODE)) gv[i] = ss;
        }
    } else {
        for (uint64_t i = tid; i < size; i += nth) {
            float vv = opt_uses_v(MODE) ? v[i] : 0.0f, ss = opt_uses_gv(MODE) ? gv[i] : 0.0f;
            w[i] = opt_weight<MODE>(o, g[i], w[i], vv, ss);
            if (opt_uses_v(MODE)) v[i] = vv;
            if (opt_uses_gv(MODE)) gv[i] = ss;
        }
    }
}

// CTA = 32 columns x 8 warps; gridDim.y row slices.
template <int MODE>
__global__ void __launch_bounds__(256)
update_biases_kernel(const OptArgs o, uint32_t batch, uint32_t width, const float* __restrict__ delta,
                     float* __restrict__ v, float* __restrict__ gv, float* __restrict__ bias,
                     float* __restrict__ partials, uint32_t* __restrict__ counters)
{
    __shared__ float sPart[8][33];
    __shared__ uint32_t sLast;
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t c = blockIdx.x * 32 + lane;
    const uint32_t R = gridDim.y, ry = blockIdx.y;
    const uint32_t r0 = (uint32_t)(((uint64_t)batch * ry) / R), r1 = (uint32_t)(((uint64_t)batch * (ry + 1)) / R);
    float sum = 0.0f;
    if (c < width) {
        const float* p = delta + c;
        uint32_t b = r0 + warp;
        for (; b + 7 * 8 < r1; b += 64) {
            float x[8];
#pragma unroll
            for (int u = 0; u < 8; u++) x[u] = __ldg(p + (size_t)(b + u * 8) * width);
#pragma unroll
            for (int u = 0; u < 8; u++) sum += x[u];
        }
        for (; b < r1; b += 8) sum += __ldg(p + (size_t)b * width);
    }
    sPart[warp][lane] = sum;
    __syncthreads();
    if (warp == 0) {
        float tot = 0.0f;
#pragma unroll
        for (int w8 = 0; w8 < 8; w8++) tot += sPart[w8][lane];
        if (R > 1) {
            if (c < width) partials[(size_t)ry * width + c] = tot;
            __threadfence();
            __syncwarp();
            if (lane == 0) {
                const uint32_t old = atomicAdd(counters + blockIdx.x, 1u);
                const uint32_t last = (old == R - 1);
                if (last) counters[blockIdx.x] = 0;
                sLast = last;
            }
            __syncwarp();
            if (!sLast) return;
            __threadfence();
            tot = 0.0f;
            if (c < width)
                for (uint32_t y = 0; y < R; y++) tot += ldg_cg_f(partials + (size_t)y * width + c);
        }
        if (c < width) {
            const float gbar = tot / (float)batch;
            float vv = opt_uses_v(MODE) ? v[c] : 0.0f, ss = opt_uses_gv(MODE) ? gv[c] : 0.0f;
            bias[c] = opt_bias<MODE>(o, gbar, bias[c], vv, ss);
            if (opt_uses_v(MODE)) v[c] = vv;
            if (opt_uses_gv(MODE)) gv[c] = ss;
        }
    }
}

__global__ void __launch_bounds__(256)
regularization_kernel(float halfLambda, float lambda1, const float* __restrict__ w, uint64_t size,
                      unsigned long long* __restrict__ acc)
{
    __shared__ double sW[8];
    double e = 0.0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < size; i += (uint64_t)gridDim.x * blockDim.x) {
        const float x = __ldg(w + i);
        e += (double)(halfLambda * x * x + lambda1 * fabsf(x));      // E/kernels.cu:2729-2731
    }
    e = warp_sum(e);
    if ((threadIdx.x & 31) == 0) sW[threadIdx.x >> 5] = e;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int i = 0; i < 8; i++) tot += sW[i];
        atomicAdd(acc, (unsigned long long)llrint(tot * (double)kErrorScaleF));
    }
}

template <int MODE>
static int launch_weights(dsb200_ctx* ctx, const OptArgs& o, uint64_t size, float* v, const float* g, float* gv, float* w)
{
    const int vec = ((((uintptr_t)g | (uintptr_t)w | (uintptr_t)v | (uintptr_t)gv) % 16) == 0) ? 1 : 0;
    uint64_t work = vec ? (size + 3) / 4 : size;
    uint64_t blocks = (work + 255) / 256;
    const uint64_t cap = (uint64_t)ctx->numSMs * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    DSB_CUDA_OK(launch_pdl(update_weights_kernel<MODE>, dim3((unsigned)blocks), dim3(256), 0, ctx->stream, o, size, v, g, gv, w, vec));
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int MODE>
static int launch_biases(dsb200_ctx* ctx, const OptArgs& o, uint32_t batch, uint32_t width, const float* delta,
                         float* v, float* gv, float* bias)
{
    const uint32_t tiles = (width + 31) / 32;
    uint32_t R = 1;
    if (tiles < (uint32_t)ctx->numSMs * 2) {
        R = ((uint32_t)ctx->numSMs * 2 + tiles - 1) / tiles;
        const uint32_t maxR = (batch + 63) / 64;                  // >= 64 rows per slice
        if (R > maxR) R = maxR;
        if (R < 1) R = 1;
    }
    if (R > 1) {
        int rc = dsb200_ctx_reserve(ctx, tiles, (size_t)R * width);
        if (rc) return rc;
    }
    dim3 grid(tiles, R);
    DSB_CUDA_OK(launch_pdl(update_biases_kernel<MODE>, grid, dim3(256), 0, ctx->stream, o, batch, width, delta, v, gv, bias, ctx->dPartials, ctx->dRowCounters));
    count_launch();
    return 0;
}

// bias update from column-sum partials [nPartials][width] (written by the fused output-layer forward pass, which sees every delta
// element anyway): fixed summation order -> deterministic
template <int MODE>
__global__ void __launch_bounds__(256)
update_biases_partials_kernel(const OptArgs o, uint32_t batch, uint32_t width, const float* __restrict__ partials, uint32_t nPartials,
                              float* __restrict__ v, float* __restrict__ gv, float* __restrict__ bias)
{
    const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    if (c >= width) return;
    float tot = 0.0f;
    for (uint32_t y = 0; y < nPartials; y++) tot += __ldg(partials + (size_t)y * width + c);
    const float gbar = tot / (float)batch;
    float vv = opt_uses_v(MODE) ? v[c] : 0.0f, ss = opt_uses_gv(MODE) ? gv[c] : 0.0f;
    bias[c] = opt_bias<MODE>(o, gbar, bias[c], vv, ss);
    if (opt_uses_v(MODE)) v[c] = vv;
    if (opt_uses_gv(MODE)) gv[c] = ss;
}

template <int MODE>
static int launch_biases_partials(dsb200_ctx* ctx, const OptArgs& o, uint32_t batch, uint32_t width, const float* partials, uint32_t nPartials,
                                  float* v, float* gv, float* bias)
{
    DSB_CUDA_OK(launch_pdl(update_biases_partials_kernel<MODE>, dim3((width + 255) / 256), dim3(256), 0, ctx->stream, o, batch, width, partials, nPartials, v, gv, bias));
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace dsb

extern "C" {

int dsb200_update_weights(dsb200_ctx* ctx, int mode, float alpha, float lambda, float lambda1, float mu, float mu1, float t,
                          uint64_t size, float* v, const float* g, float* gv, float* w)
{
    DSB_PROFILE_T(ctx, "update_weights", size);
    using namespace dsb;
    if (!ctx || !g || !w) return fail(ctx, DSB200_EINVAL, "update_weights: null argument");
    if (mode < 0 || mode > DSB200_ADAM) return fail(ctx, DSB200_EINVAL, "update_weights: bad mode");
    if (opt_uses_v(mode) && !v) return fail(ctx, DSB200_EINVAL, "update_weights: velocity buffer missing");
    if (opt_uses_gv(mode) && !gv) return fail(ctx, DSB200_EINVAL, "update_weights: gradient-velocity buffer missing");
    if (!size) return 0;
    const OptArgs o = make_opt(mode, alpha, lambda, lambda1, mu, mu1, t);
    switch (mode) {
    case DSB200_SGD:      return launch_weights<DSB200_SGD>(ctx, o, size, nullptr, g, nullptr, w);
    case DSB200_MOMENTUM: return launch_weights<DSB200_MOMENTUM>(ctx, o, size, v, g, nullptr, w);
    case DSB200_ADAGRAD:  return launch_weights<DSB200_ADAGRAD>(ctx, o, size, v, g, nullptr, w);
    case DSB200_NESTEROV: return launch_weights<DSB200_NESTEROV>(ctx, o, size, v, g, nullptr, w);
    case DSB200_RMSPROP:  return launch_weights<DSB200_RMSPROP>(ctx, o, size, v, g, nullptr, w);
    case DSB200_ADADELTA: return launch_weights<DSB200_ADADELTA>(ctx, o, size, v, g, gv, w);
    default:              return launch_weights<DSB200_ADAM>(ctx, o, size, v, g, gv, w);
    }
}

int dsb200_update_biases(dsb200_ctx* ctx, int mode, float alpha, float mu, float mu1, float t, uint32_t batch, uint32_t width,
                         const float* delta, float* v, float* gv, float* bias)
{
    DSB_PROFILE_T(ctx, "update_biases", width);
    using namespace dsb;
    if (!ctx || !delta || !bias) return fail(ctx, DSB200_EINVAL, "update_biases: null argument");
    if (mode < 0 || mode > DSB200_ADAM) return fail(ctx, DSB200_EINVAL, "update_biases: bad mode");
    if (opt_uses_v(mode) && !v) return fail(ctx, DSB200_EINVAL, "update_biases: velocity buffer missing");
    if (opt_uses_gv(mode) && !gv) return fail(ctx, DSB200_EINVAL, "update_biases: gradient-velocity buffer missing");
    if (!width || !batch) return 0;
    const OptArgs o = make_opt(mode, alpha, 0.0f, 0.0f, mu, mu1, t);
    switch (mode) {
    case DSB200_SGD:      return launch_biases<DSB200_SGD>(ctx, o, batch, width, delta, nullptr, nullptr, bias);
    case DSB200_MOMENTUM: return launch_biases<DSB200_MOMENTUM>(ctx, o, batch, width, delta, v, nullptr, bias);
    case DSB200_ADAGRAD:  return launch_biases<DSB200_ADAGRAD>(ctx, o, batch, width, delta, v, nullptr, bias);
    case DSB200_NESTEROV: return launch_biases<DSB200_NESTEROV>(ctx, o, batch, width, delta, v, nullptr, bias);
    case DSB200_RMSPROP:  return launch_biases<DSB200_RMSPROP>(ctx, o, batch, width, delta, v, nullptr, bias);
    case DSB200_ADADELTA: return launch_biases<DSB200_ADADELTA>(ctx, o, batch, width, delta, v, gv, bias);
    default:              return launch_biases<DSB200_ADAM>(ctx, o, batch, width, delta, v, gv, bias);
    }
}

int dsb200_update_biases_partials(dsb200_ctx* ctx, int mode, float alpha, float mu, float mu1, float t, uint32_t batch, uint32_t width,
                                  const float* partials, uint32_t nPartials, float* v, float* gv, float* bias)
{
    DSB_PROFILE_T(ctx, "update_biases_partials", width);
    using namespace dsb;
    if (!ctx || !partials || !bias) return fail(ctx, DSB200_EINVAL, "update_biases_partials: null argument");
    if (mode < 0 || mode > DSB200_ADAM) return fail(ctx, DSB200_EINVAL, "update_biases_partials: bad mode");
    if (opt_uses_v(mode) && !v) return fail(ctx, DSB200_EINVAL, "update_biases_partials: velocity buffer missing");
    if (opt_uses_gv(mode) && !gv) return fail(ctx, DSB200_EINVAL, "update_biases_partials: gradient-velocity buffer missing");
    if (!width || !batch) return 0;
    const OptArgs o = make_opt(mode, alpha, 0.0f, 0.0f, mu, mu1, t);
    switch (mode) {
    case DSB200_SGD:      return launch_biases_partials<DSB200_SGD>(ctx, o, batch, width, partials, nPartials, nullptr, nullptr, bias);
    case DSB200_MOMENTUM: return launch_biases_partials<DSB200_MOMENTUM>(ctx, o, batch, width, partials, nPartials, v, nullptr, bias);
    case DSB200_ADAGRAD:  return launch_biases_partials<DSB200_ADAGRAD>(ctx, o, batch, width, partials, nPartials, v, nullptr, bias);
    case DSB200_NESTEROV: return launch_biases_partials<DSB200_NESTEROV>(ctx, o, batch, width, partials, nPartials, v, nullptr, bias);
    case DSB200_RMSPROP:  return launch_biases_partials<DSB200_RMSPROP>(ctx, o, batch, width, partials, nPartials, v, nullptr, bias);
    case DSB200_ADADELTA: return launch_biases_partials<DSB200_ADADELTA>(ctx, o, batch, width, partials, nPartials, v, gv, bias);
    default:              return launch_biases_partials<DSB200_ADAM>(ctx, o, batch, width, partials, nPartials, v, gv, bias);
    }
}

int dsb200_regularization_error_async(dsb200_ctx* ctx, float lambda, float lambda1, const float* w, uint64_t size, unsigned long long* pDevAcc)
{
    DSB_PROFILE_T(ctx, "regularization_error", size);
    using namespace dsb;
    if (!ctx || !w || !pDevAcc) return fail(ctx, DSB200_EINVAL, "regularization_error_async: null argument");
    if (!size) return 0;
    uint64_t blocks = (size + 255) / 256;
    const uint64_t cap = (uint64_t)ctx->numSMs * 8;
    if (blocks > cap) blocks = cap;
    regularization_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(0.5f * lambda, lambda1, w, size, pDevAcc);
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    return 0;
}

int dsb200_regularization_error(dsb200_ctx* ctx, float lambda, float lambda1, const float* w, uint64_t size, float* out)
{
    DSB_PROFILE(ctx, "regularization_error");
    using namespace dsb;
    if (!ctx || !w || !out) return fail(ctx, DSB200_EINVAL, "regularization_error: null argument");
    DSB_CUDA_OK(cudaMemsetAsync(ctx->dAccumulator, 0, sizeof(unsigned long long), ctx->stream));
    uint64_t blocks = (size + 255) / 256;
    const uint64_t cap = (uint64_t)ctx->numSMs * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    regularization_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(0.5f * lambda, lambda1, w, size, ctx->dAccumulator);
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    DSB_CUDA_OK(cudaMemcpyAsync(ctx->hAccumulator, ctx->dAccumulator, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    DSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    *out = (float)((double)(long long)ctx->hAccumulator[0] * kOneOverErrorScale);
    return 0;
}

}  // extern "C"
