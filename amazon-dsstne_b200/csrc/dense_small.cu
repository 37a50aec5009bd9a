// dense_small.cu -- SIMT fp32 kernels for the SMALL dense layers of the path (hot-path row a11): the 128 x 128 hidden
// weights of BASELINE config 2 cover eight 128 x 128 tiles, too few for the persistent tcgen05 kernel (gemm_tc.cu) and
// latency bound in the library SGEMM, where each layer costs three launches (kClearUnit + cublasSgemm + activation,
// E/NNLayer.cpp:1009, 1073, 1157) or two (cublasSgemm + kCalculateHadamardProduct, E/NNLayer.cpp:2274, 2137).
//   dsb200_gemm_fwd_bias_act (small shapes)  C[B][n]  = act(A[B][k] * W[k][n] + bias[n])                one launch
//   dsb200_gemm_dx_hadamard                  Dp[B][k] = (D[B][n] * W[k][n]^T) (.) f'(unit[B][k]) * s    one launch
// Exact fp32 FMA arithmetic (the 1e-5 parity mode).  Tile: 32 rows x 64 columns per 512-thread CTA, 2 x 2 outputs per
// thread, K in chunks of 64 through shared memory; 64 CTAs for a 1,024 x 128 layer.
#include "common.cuh"
#include "launch.h"

namespace dsb {

constexpr int kDsRows = 32, kDsCols = 64, kDsK = 64, kDsThreads = 512;   // 2 x 2 outputs per thread: short dependent chains, 16 warps per CTA

__device__ __forceinline__ float ds_act(int act, float z, float slope, float alpha, float lambda)
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

// f'(x) through the activation value, times the incoming delta (kCalculateHadamardProduct, E/kDelta.cu:9021-9151)
__device__ __forceinline__ float ds_hadamard(int act, float x, float d, float scale, float slope, float alpha, float lambda)
{
    switch (act) {
    case DSB200_ACT_SIGMOID: return x * (1.0f - x) * d;
    case DSB200_ACT_TANH:    { const float xs = x * (1.0f / scale); return scale * (1.0f - xs * xs) * d; }
    case DSB200_ACT_RELU:    return (x <= 0.0f) ? 0.0f : d;
    case DSB200_ACT_LRELU:   return (x <= 0.0f) ? d * slope : d;
    case DSB200_ACT_ELU:     return (x <= 0.0f) ? d * (x + alpha) : d;
    case DSB200_ACT_SELU:    return (x > 0.0f) ? d * lambda : d * (x + lambda * alpha);
    default:                 return d;
    }
}

// C[M][N] = epilogue(A[M][K] * op(B)),  op(B)(kk, n) = TRANSB ? B[n * ldb + kk] : B[kk * ldb + n]
// EPI 0: + bias[n], activation;  EPI 1: Hadamard with unit[M][N] (delta of the layer below)
// 512 threads; thread (ty, tx) owns rows 2 ty, 2 ty + 1 and columns 2 tx, 2 tx + 1 (the layers are small: short dependent
// chains and 16 warps per CTA matter more than FMA density).
template <bool TRANSB, int EPI>
__global__ void __launch_bounds__(kDsThreads)
dense_small_kernel(const float* __restrict__ A, uint32_t lda, const float* __restrict__ Bm, uint32_t ldb, float* __restrict__ C, uint32_t ldc,
                   uint32_t M, uint32_t N, uint32_t K, const float* __restrict__ bias, const float* __restrict__ unit, int act, float scale,
                   float slope, float alpha, float lambda)
{
    __shared__ __align__(16) float sA[kDsK][kDsRows + 4];
    __shared__ __align__(16) float sB[kDsK][kDsCols + 4];
    const uint32_t tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;           // 32 column pairs x 16 row pairs
    const uint32_t m0 = blockIdx.y * kDsRows, n0 = blockIdx.x * kDsCols;
    float acc[2][2] = {{0.0f, 0.0f}, {0.0f, 0.0f}};

    for (uint32_t k0 = 0; k0 < K; k0 += kDsK) {
        // A chunk: 32 rows x 64 k, read coalesced along k, stored k-major
#pragma unroll
        for (int i = 0; i < (kDsRows * kDsK) / kDsThreads; i++) {
            const uint32_t e = tid + i * kDsThreads, r = e >> 6, c = e & 63;
            sA[c][r] = (m0 + r < M && k0 + c < K) ? __ldg(A + (size_t)(m0 + r) * lda + k0 + c) : 0.0f;
        }
        // B chunk: 64 k x 64 n
#pragma unroll
        for (int i = 0; i < (kDsK * kDsCols) / kDsThreads; i++) {
            const uint32_t e = tid + i * kDsThreads;
            if (TRANSB) {                                                       // B[n][kk]: coalesced along kk
                const uint32_t nn = e >> 6, kk = e & 63;
                sB[kk][nn] = (n0 + nn < N && k0 + kk < K) ? __ldg(Bm + (size_t)(n0 + nn) * ldb + k0 + kk) : 0.0f;
            } else {                                                            // B[kk][n]: coalesced along n
                const uint32_t kk = e >> 6, nn = e & 63;
                sB[kk][nn] = (n0 + nn < N && k0 + kk < K) ? __ldg(Bm + (size_t)(k0 + kk) * ldb + n0 + nn) : 0.0f;
            }
        }
        __syncthreads();
#pragma unroll 16
        for (int kk = 0; kk < kDsK; kk++) {
            const float2 a = *reinterpret_cast<const float2*>(&sA[kk][ty * 2]);   // warp-uniform: broadcast
            const float2 b = *reinterpret_cast<const float2*>(&sB[kk][tx * 2]);
            acc[0][0] = fmaf(a.x, b.x, acc[0][0]); acc[0][1] = fmaf(a.x, b.y, acc[0][1]);
            acc[1][0] = fmaf(a.y, b.x, acc[1][0]); acc[1][1] = fmaf(a.y, b.y, acc[1][1]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; i++) {
        const uint32_t m = m0 + ty * 2 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const uint32_t n = n0 + tx * 2 + j;
            if (n >= N) continue;
            float v = acc[i][j];
            if (EPI == 0) v = ds_act(act, v + __ldg(bias + n), slope, alpha, lambda);
            else          v = ds_hadamard(act, __ldg(unit + (size_t)m * ldc + n), v, scale, slope, alpha, lambda);
            C[(size_t)m * ldc + n] = v;
        }
    }
}

int dense_small_fwd(dsb200_ctx* ctx, uint32_t B, uint32_t k, uint32_t n, const float* A, const float* W, const float* bias, int act, float* C,
                    float slope, float alpha, float lambda)
{
    dim3 grid((n + kDsCols - 1) / kDsCols, (B + kDsRows - 1) / kDsRows);
    dense_small_kernel<false, 0><<<grid, kDsThreads, 0, ctx->stream>>>(A, k, W, n, C, n, B, n, k, bias, nullptr, act, 1.0f, slope, alpha, lambda);
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace dsb

extern "C" int dsb200_gemm_dx_hadamard(dsb200_ctx* ctx, uint32_t B, uint32_t k, uint32_t n, const float* D, const float* W, int activation,
                                       float scale, const float* pUnit, float* Dp, float slope, float alpha, float lambda)
{
    DSB_PROFILE(ctx, "gemm_dx_hadamard");
    using namespace dsb;
    if (!ctx || !D || !W || !pUnit || !Dp) return fail(ctx, DSB200_EINVAL, "gemm_dx_hadamard: null argument");
    if (!B || !k || !n) return 0;
    dim3 grid((k + kDsCols - 1) / kDsCols, (B + kDsRows - 1) / kDsRows);
    dense_small_kernel<true, 1><<<grid, kDsThreads, 0, ctx->stream>>>(D, n, W, n, Dp, k, B, k, n, nullptr, pUnit, activation, scale, slope, alpha, lambda);
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    return 0;
}
