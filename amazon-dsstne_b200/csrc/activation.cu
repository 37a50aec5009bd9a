// activation.cu -- stand-alone activations (row a9) and hidden-layer backward elementwise
// kernels (row a10).
//
// Replaces kCalculate{Sigmoid,Tanh,RELU,LRELU,ELU,SELU,SoftMax}Activation (E/kActivation.cu:46-236),
// kCalculateHadamardProduct (E/kDelta.cu:9021-9151) and kCalculateSparsenessPenalty
// (E/kDelta.cu:8979-9018).  In the training path these are normally fused into their producers
// (dsb200_sparse_z_bias_act, dsb200_output_pass, the GEMM epilogues); the stand-alone entry points
// exist because the drop-in boundary exposes them.
//  * elementwise kernels: 128-bit accesses, grid = multiple of the SM count;
//  * softmax: one CTA per row, max and sum reduced through shuffles and shared memory (the reference
//    round-trips a fixed-point shared atomic);
//  * sparseness penalty: the reference uses one thread per hidden unit looping over the batch
//    twice with stride `stride`; here a CTA owns 32 units x 8 batch slices, coalesced 128-byte row
//    segments, fixed-order combine => deterministic column means.
#include "common.cuh"
#include "launch.h"

namespace dsb {

__device__ __forceinline__ float act_fwd(int act, float z, float slope, float alpha, float lambda)
{
    switch (act) {
    case DSB200_ACT_SIGMOID: return 1.0f / (1.0f + expf(-z));
    case DSB200_ACT_TANH:    return tanhf(z);
    case DSB200_ACT_RELU:    return fmaxf(0.0f, z);
    case DSB200_ACT_LRELU:   return fmaxf(z, z * slope);
    case DSB200_ACT_ELU:     return (z > 0.0f) ? z : alpha * (expf(z) - 1.0f);
    case DSB200_ACT_SELU:    return (z > 0.0f) ? lambda * z : lambda * alpha * (expf(z) - 1.0f);
    default:                 return z;
    }
}

__global__ void __launch_bounds__(256)
activation_kernel(int act, float* __restrict__ data, uint64_t size, float slope, float alpha, float lambda, int vec)
{
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (uint64_t)gridDim.x * blockDim.x;
    if (vec) {
        const uint64_t n4 = size >> 2;
        for (uint64_t i = tid; i < n4; i += nth) {
            float4 x = reinterpret_cast<float4*>(data)[i];
            x.x = act_fwd(act, x.x, slope, alpha, lambda); x.y = act_fwd(act, x.y, slope, alpha, lambda);
            x.z = act_fwd(act, x.z, slope, alpha, lambda); x.w = act_fwd(act, x.w, slope, alpha, lambda);
            reinterpret_cast<float4*>(data)[i] = x;
        }
        for (uint64_t i = (n4 << 2) + tid; i < size; i += nth) data[i] = act_fwd(act, data[i], slope, alpha, lambda);
    } else {
        for (uint64_t i = tid; i < size; i += nth) data[i] = act_fwd(act, data[i], slope, alpha, lambda);
    }
}

// E/kActivation.cu:155-229: a = min(1, exp(z - max) / sum exp(z - max))
__global__ void __launch_bounds__(256)
softmax_kernel(float* __restrict__ data, uint32_t stride)
{
    __shared__ float sRed[8];
    __shared__ float sBcast;
    float* row = data + (size_t)blockIdx.x * stride;
    const uint32_t tid = threadIdx.x;
    float mx = -9999999999.0f;
    for (uint32_t i = tid; i < stride; i += 256) mx = fmaxf(mx, row[i]);
    mx = warp_max(mx);
    if ((tid & 31) == 0) sRed[tid >> 5] = mx;
    __syncthreads();
    if (tid == 0) { float m = sRed[0]; for (int i = 1; i < 8; i++) m = fmaxf(m, sRed[i]); sBcast = m; }
    __syncthreads();
    mx = sBcast;
    float sum = 0.0f;
    for (uint32_t i = tid; i < stride; i += 256) sum += expf(row[i] - mx);
    sum = warp_sum(sum);
    __syncthreads();
    if ((tid & 31) == 0) sRed[tid >> 5] = sum;
    __syncthreads();
    if (tid == 0) { float s = 0.0f; for (int i = 0; i < 8; i++) s += sRed[i]; sBcast = 1.0f / s; }
    __syncthreads();
    const float norm = sBcast;
    for (uint32_t i = tid; i < stride; i += 256) row[i] = fminf(1.0f, expf(row[i] - mx) * norm);
}

__device__ __forceinline__ float hadamard_elem(int act, float x, float d, float scale, float oneOverScale,
                                               float slope, float alpha, float lambda)
{
    switch (act) {
    case DSB200_ACT_SIGMOID: return x * (1.0f - x) * d;                                   // E/kDelta.cu:9023-9031 (scale unused)
    case DSB200_ACT_TANH:    { x *= oneOverScale; return scale * (1.0f - x * x) * d; }    // :9036-9046
    case DSB200_ACT_RELU:    return (x <= 0.0f) ? 0.0f : d;                                // :9048-9058
    case DSB200_ACT_LRELU:   return (x <= 0.0f) ? d * slope : d;                           // :9060-9072
    case DSB200_ACT_ELU:     return (x <= 0.0f) ? d * (x + alpha) : d;                     // :9074-9084
    case DSB200_ACT_SELU:    return (x > 0.0f) ? d * lambda : d * (x + lambda * alpha);    // :9086-9104
    default:                 return d;                                                     // Linear: no kernel
    }
}

__global__ void __launch_bounds__(256)
hadamard_kernel(int act, uint64_t size, float scale, const float* __restrict__ unit, float* __restrict__ delta,
                float slope, float alpha, float lambda, int vec)
{
    const float oos = 1.0f / scale;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (uint64_t)gridDim.x * blockDim.x;
    if (vec) {
        const uint64_t n4 = size >> 2;
        for (uint64_t i = tid; i < n4; i += nth) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(unit) + i);
            float4 d = reinterpret_cast<float4*>(delta)[i];
            d.x = hadamard_elem(act, x.x, d.x, scale, oos, slope, alpha, lambda);
            d.y = hadamard_elem(act, x.y, d.y, scale, oos, slope, alpha, lambda);
            d.z = hadamard_elem(act, x.z, d.z, scale, oos, slope, alpha, lambda);
            d.w = hadamard_elem(act, x.w, d.w, scale, oos, slope, alpha, lambda);
            reinterpret_cast<float4*>(delta)[i] = d;
        }
        for (uint64_t i = (n4 << 2) + tid; i < size; i += nth)
            delta[i] = hadamard_elem(act, unit[i], delta[i], scale, oos, slope, alpha, lambda);
    } else {
        for (uint64_t i = tid; i < size; i += nth)
            delta[i] = hadamard_elem(act, unit[i], delta[i], scale, oos, slope, alpha, lambda);
    }
}

// E/kDelta.cu:8979-9007.  CTA = 32 units x 8 warps over the batch rows; one CTA per 32 units
// (hidden layers are narrow, the batch loop is short), fixed-order combine.
__global__ void __launch_bounds__(256)
sparseness_penalty_kernel(uint32_t batch, uint32_t stride, const float* __restrict__ unit, float* __restrict__ delta,
                          float p, float beta)
{
    __shared__ float sPart[8][33];
    __shared__ float sPenalty[32];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t c = blockIdx.x * 32 + lane;
    float sum = 0.0f;
    if (c < stride)
        for (uint32_t b = warp; b < batch; b += 8) sum += __ldg(unit + (size_t)b * stride + c);
    sPart[warp][lane] = sum;
    __syncthreads();
    if (warp == 0) {
        float pi = 0.0f;
#pragma unroll
        for (int w8 = 0; w8 < 8; w8++) pi += sPart[w8][lane];
        pi /= (float)batch;
        pi = fmaxf(kMinActivation, fminf(kMaxActivation, pi));
        sPenalty[lane] = beta * (-p / pi + (1.0f - p) / (1.0f - pi));
    }
    __syncthreads();
    if (c < stride) {
        const float pen = sPenalty[lane];
        for (uint32_t b = warp; b < batch; b += 8) delta[(size_t)b * stride + c] += pen;
    }
}

static unsigned grid_for(dsb200_ctx* ctx, uint64_t work)
{
    uint64_t blocks = (work + 255) / 256;
    const uint64_t cap = (uint64_t)ctx->numSMs * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (unsigned)blocks;
}

}  // namespace dsb

extern "C" {

int dsb200_activation(dsb200_ctx* ctx, int act, float* pData, uint32_t batch, uint32_t stride, float slope, float alpha, float lambda)
{
    DSB_PROFILE(ctx, "activation");
    using namespace dsb;
    if (!ctx || !pData) return fail(ctx, DSB200_EINVAL, "activation: null argument");
    const uint64_t size = (uint64_t)batch * stride;
    if (!size || act == DSB200_ACT_LINEAR) return 0;
    if (act == DSB200_ACT_SOFTMAX) {
        softmax_kernel<<<batch, 256, 0, ctx->stream>>>(pData, stride);
    } else if (act == DSB200_ACT_SIGMOID || act == DSB200_ACT_TANH || act == DSB200_ACT_RELU || act == DSB200_ACT_LRELU ||
               act == DSB200_ACT_ELU || act == DSB200_ACT_SELU) {
        const int vec = ((uintptr_t)pData % 16) == 0;
        activation_kernel<<<grid_for(ctx, vec ? size / 4 + 1 : size), 256, 0, ctx->stream>>>(act, pData, size, slope, alpha, lambda, vec);
    } else {
        return fail(ctx, DSB200_EUNSUPPORTED, "activation: not on the hot path");
    }
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    return 0;
}

int dsb200_hadamard(dsb200_ctx* ctx, int act, uint64_t size, float scale, const float* pUnit, float* pDelta,
                    float slope, float alpha, float lambda)
{
    DSB_PROFILE(ctx, "hadamard");
    using namespace dsb;
    if (!ctx || !pUnit || !pDelta) return fail(ctx, DSB200_EINVAL, "hadamard: null argument");
    if (!size || act == DSB200_ACT_LINEAR) return 0;
    const int vec = ((((uintptr_t)pUnit | (uintptr_t)pDelta) % 16) == 0);
    hadamard_kernel<<<grid_for(ctx, vec ? size / 4 + 1 : size), 256, 0, ctx->stream>>>(act, size, scale, pUnit, pDelta, slope, alpha, lambda, vec);
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    return 0;
}

int dsb200_sparseness_penalty(dsb200_ctx* ctx, uint32_t batch, uint32_t stride, const float* pUnit, float* pDelta, float p, float beta)
{
    DSB_PROFILE(ctx, "sparseness_penalty");
    using namespace dsb;
    if (!ctx || !pUnit || !pDelta) return fail(ctx, DSB200_EINVAL, "sparseness_penalty: null argument");
    if (!batch || !stride) return 0;
    sparseness_penalty_kernel<<<(stride + 31) / 32, 256, 0, ctx->stream>>>(batch, stride, pUnit, pDelta, p, beta);
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"

// kAddBuffers (E/kernels.cu:39-57): pDst += pSrc -- used after a reduce-scatter that lands in scratch
namespace dsb {
__global__ void __launch_bounds__(256) add_buffers_kernel(float* __restrict__ dst, const float* __restrict__ src, uint64_t size)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < size; i += (uint64_t)gridDim.x * blockDim.x) dst[i] += src[i];
}
}  // namespace dsb

extern "C" int dsb200_add_buffers(dsb200_ctx* ctx, float* pDst, const float* pSrc, uint64_t size)
{
    DSB_PROFILE(ctx, "add_buffers");
    using namespace dsb;
    if (!ctx || !pDst || !pSrc) return fail(ctx, DSB200_EINVAL, "add_buffers: null argument");
    if (!size) return 0;
    add_buffers_kernel<<<grid_for(ctx, size), 256, 0, ctx->stream>>>(pDst, pSrc, size);
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    return 0;
}
