// random.cu -- counter-based uniform randoms for input denoising ("next" row 4 of SURVEY 8f).
//
// The reference refills a whole-dataset buffer with cuRAND (XORWOW) once per epoch
// (NNDataSet::GenerateDenoisingData, E/NNTypes.cpp:1617-1629; E/NNNetwork.cpp:1586-1593) and the
// kernels compare pRandom[j] < p.  cuRAND's stream cannot be reproduced bit for bit (SURVEY 8c:
// "parity unpinned"), so this library keeps the SAME interface -- a plain device buffer of
// U(0,1] floats, which tests can also fill themselves -- and generates it with a stateless
// integer hash (two rounds of the SplitMix64 finaliser over (seed, stream, index)).
#include "common.cuh"
#include "launch.h"

namespace dsb {

__device__ __forceinline__ uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256)
fill_uniform_kernel(float* __restrict__ out, uint64_t n, uint64_t seed, uint64_t stream)
{
    const uint64_t key = mix64(seed ^ mix64(stream + 0x9e3779b97f4a7c15ull));
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = mix64(key + i * 0x9e3779b97f4a7c15ull);
        // 24 random bits -> (0, 1]  (curandGenerateUniform excludes 0, includes 1)
        out[i] = ((float)(uint32_t)(r >> 40) + 1.0f) * (1.0f / 16777216.0f);
    }
}

// NNLayer::CalculateDropout (E/NNLayer.cpp:1685-1708) = curandGenerateUniform + kCalculateDropout /
// kCalculateScaledBiasedDropout (E/kernels.cu:4497-4537) in ONE pass: the uniform of element (row, col) is the same
// counter-based function of (seed, stream, row * fullStride + colOffset + col) as fill_uniform_kernel, so no random
// buffer is written or read, and the mask does not depend on how the layer is sharded over ranks.
__global__ void __launch_bounds__(256)
dropout_kernel(float* __restrict__ unit, uint32_t batch, uint32_t stride, uint32_t fullStride, uint32_t colOffset, float p, float target,
               float a, float b, uint64_t seed, uint64_t stream)
{
    const uint64_t key = mix64(seed ^ mix64(stream + 0x9e3779b97f4a7c15ull));
    const uint64_t n = (uint64_t)batch * stride;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t row = i / stride, col = i - row * stride;
        const uint64_t g = row * fullStride + colOffset + col;
        const float r = ((float)(uint32_t)(mix64(key + g * 0x9e3779b97f4a7c15ull) >> 40) + 1.0f) * (1.0f / 16777216.0f);
        unit[i] = (r < p) ? target : a * unit[i] + b;
    }
}

}  // namespace dsb

extern "C" int dsb200_dropout(dsb200_ctx* ctx, int activation, float* pUnit, uint32_t batch, uint32_t stride, uint32_t fullStride,
                              uint32_t colOffset, float p, float eluAlpha, float seluLambda, uint64_t seed, uint64_t stream)
{
    DSB_PROFILE(ctx, "dropout");
    using namespace dsb;
    if (!ctx || !pUnit) return fail(ctx, DSB200_EINVAL, "dropout: null argument");
    if (!(p > 0.0f) || !batch || !stride) return 0;
    if (p >= 1.0f) return fail(ctx, DSB200_EINVAL, "dropout: p must be < 1");
    float target, a, b;
    if (activation == DSB200_ACT_ELU || activation == DSB200_ACT_SELU) {           // E/NNLayer.cpp:1688-1692
        const float lambda = (activation == DSB200_ACT_SELU) ? seluLambda : 1.0f;
        const float alpha = -lambda * eluAlpha, q = 1.0f - p;
        a = 1.0f / sqrtf(q + alpha * alpha * p * q);
        b = -a * p * alpha;
        target = a * alpha + b;
    } else {
        target = (activation == DSB200_ACT_SIGMOID) ? 0.5f : 0.0f;
        a = (target == 0.0f) ? 1.0f / (1.0f - p) : 1.0f;
        b = 0.0f;
    }
    const uint64_t n = (uint64_t)batch * stride;
    uint64_t blocks = (n + 255) / 256;
    const uint64_t cap = (uint64_t)ctx->numSMs * 16;
    if (blocks > cap) blocks = cap;
    dropout_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(pUnit, batch, stride, fullStride, colOffset, p, target, a, b, seed, stream);
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int dsb200_fill_uniform(dsb200_ctx* ctx, float* pOut, uint64_t n, uint64_t seed, uint64_t stream)
{
    DSB_PROFILE(ctx, "fill_uniform");
    using namespace dsb;
    if (!ctx || (!pOut && n)) return fail(ctx, DSB200_EINVAL, "fill_uniform: null argument");
    if (!n) return 0;
    uint64_t blocks = (n + 255) / 256;
    const uint64_t cap = (uint64_t)ctx->numSMs * 16;
    if (blocks > cap) blocks = cap;
    fill_uniform_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(pOut, n, seed, stream);
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    return 0;
}
