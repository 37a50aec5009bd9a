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

namespace dsb {
int dense_small_dx_hadamard(dsb200_ctx* ctx, uint32_t B, uint32_t k, uint32_t n, const float* D, const float* W, int activation, float scale,
                            const float* pUnit, float* Dp, float slope, float alpha, float lambda)
{
    dim3 grid((k + kDsCols - 1) / kDsCols, (B + kDsRows - 1) / kDsRows);
    dense_small_kernel<true, 1><<<grid, kDsThreads, 0, ctx->stream>>>(D, n, W, n, Dp, k, B, k, n, nullptr, pUnit, activation, scale, slope, alpha, lambda);
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    return 0;
}
}  // namespace dsb

// =====================================================================================================================
// dsb200_dense_update: weight gradient + optimizer step + bias update of a SMALL dense layer in ONE launch.
// Replaces, for the 128 x 128 hidden weights of BASELINE config 2, cublasSgemm (E/NNLayer.cpp:2223; 2 launches with its split-K
// reduce) + k*UpdateWeights + k*UpdateBiases (E/NNWeight.cpp:729-794): four launches of ~6-14 us each, all latency bound.
//   g[i][j]  = galpha * sum_b X[b][i] * D[b][j]          (never written)
//   W[i][j]  = opt_weight(g, W, V, GV)                    (optimizer.cuh, the rules of E/kernels.cu:2746-3199)
//   bias[j]  = opt_bias(sum_b D[b][j] / B, bias, ...)
// Grid: x = row i of W (0..k-1) plus one extra index for the bias row (X == 1), y = tiles of 128 columns.  512 threads = 128
// columns x 4 batch groups; the X column of the row is staged in shared memory once and read as a broadcast; the four group sums
// are added in a fixed order (deterministic, exact fp32 FMA chains).
// =====================================================================================================================
#include "optimizer.cuh"

namespace dsb {

constexpr int kDuCols = 128, kDuGroups = 4, kDuThreads = kDuCols * kDuGroups, kDuMaxB = 4096;

struct DuArgs {
    const float* X; const float* D; uint32_t B, k, n;
    float galpha;
    float* W; float* V; float* GV;
    float* bias; float* bV; float* bGV;
};

template <int MODE>
__global__ void __launch_bounds__(kDuThreads)
dense_update_kernel(const DuArgs a, const OptArgs ow, const OptArgs ob)
{
    __shared__ float sX[kDuMaxB];
    __shared__ float sAcc[kDuGroups][kDuCols];
    const uint32_t i = blockIdx.x, j = blockIdx.y * kDuCols + (threadIdx.x & (kDuCols - 1)), g = threadIdx.x / kDuCols;
    const bool biasRow = i == a.k;
    for (uint32_t b = threadIdx.x; b < a.B; b += kDuThreads) sX[b] = biasRow ? 1.0f : __ldg(a.X + (size_t)b * a.k + i);
    __syncthreads();
    const uint32_t b0 = (uint32_t)(((uint64_t)a.B * g) / kDuGroups), b1 = (uint32_t)(((uint64_t)a.B * (g + 1)) / kDuGroups);
    float acc = 0.0f;
    if (j < a.n) {
        const float* p = a.D + j;
        uint32_t b = b0;
        for (; b + 8 <= b1; b += 8) {
            float d[8];
#pragma unroll
            for (int u = 0; u < 8; u++) d[u] = __ldg(p + (size_t)(b + u) * a.n);
#pragma unroll
            for (int u = 0; u < 8; u++) acc = fmaf(sX[b + u], d[u], acc);
        }
        for (; b < b1; b++) acc = fmaf(sX[b], __ldg(p + (size_t)b * a.n), acc);
    }
    sAcc[g][threadIdx.x & (kDuCols - 1)] = acc;
    __syncthreads();
    if (g != 0 || j >= a.n) return;
    float sum = 0.0f;
#pragma unroll
    for (int q = 0; q < kDuGroups; q++) sum += sAcc[q][threadIdx.x];
    if (biasRow) {
        float vv = opt_uses_v(MODE) ? a.bV[j] : 0.0f, ss = opt_uses_gv(MODE) ? a.bGV[j] : 0.0f;
        a.bias[j] = opt_bias<MODE>(ob, sum / (float)a.B, a.bias[j], vv, ss);
        if (opt_uses_v(MODE)) a.bV[j] = vv;
        if (opt_uses_gv(MODE)) a.bGV[j] = ss;
    } else {
        const size_t e = (size_t)i * a.n + j;
        float vv = opt_uses_v(MODE) ? a.V[e] : 0.0f, ss = opt_uses_gv(MODE) ? a.GV[e] : 0.0f;
        a.W[e] = opt_weight<MODE>(ow, a.galpha * sum, a.W[e], vv, ss);
        if (opt_uses_v(MODE)) a.V[e] = vv;
        if (opt_uses_gv(MODE)) a.GV[e] = ss;
    }
}

template <int MODE>
static int launch_dense_update(dsb200_ctx* ctx, const DuArgs& a, const OptArgs& ow, const OptArgs& ob)
{
    dim3 grid(a.k + (a.bias ? 1u : 0u), (a.n + kDuCols - 1) / kDuCols);
    dense_update_kernel<MODE><<<grid, kDuThreads, 0, ctx->stream>>>(a, ow, ob);
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace dsb

extern "C" int dsb200_dense_update(dsb200_ctx* ctx, int mode, uint32_t B, uint32_t k, uint32_t n, float galpha, const float* X, const float* D,
                                   float alpha, float lambda, float lambda1, float mu, float mu1, float t, float* pWeightVelocity,
                                   float* pWeightGradientVelocity, float* pWeight, float* pBiasVelocity, float* pBiasGradientVelocity, float* pBias)
{
    DSB_PROFILE_T(ctx, "dense_update", (unsigned long long)k * n);
    using namespace dsb;
    if (!ctx || !X || !D || !pWeight) return fail(ctx, DSB200_EINVAL, "dense_update: null argument");
    if (mode < 0 || mode > DSB200_ADAM) return fail(ctx, DSB200_EINVAL, "dense_update: bad mode");
    if (opt_uses_v(mode) && (!pWeightVelocity || (pBias && !pBiasVelocity))) return fail(ctx, DSB200_EINVAL, "dense_update: velocity buffer missing");
    if (opt_uses_gv(mode) && (!pWeightGradientVelocity || (pBias && !pBiasGradientVelocity))) return fail(ctx, DSB200_EINVAL, "dense_update: gradient-velocity buffer missing");
    if (B > (uint32_t)kDuMaxB) return fail(ctx, DSB200_EUNSUPPORTED, "dense_update: batch above 4,096 (use dsb200_gemm_dw + dsb200_update_weights + dsb200_update_biases)");
    if (!B || !k || !n) return 0;
    DuArgs a{X, D, B, k, n, galpha, pWeight, pWeightVelocity, pWeightGradientVelocity, pBias, pBiasVelocity, pBiasGradientVelocity};
    const OptArgs ow = make_opt(mode, alpha, lambda, lambda1, mu, mu1, t), ob = make_opt(mode, alpha, 0.0f, 0.0f, mu, mu1, t);
    switch (mode) {
    case DSB200_SGD:      return launch_dense_update<DSB200_SGD>(ctx, a, ow, ob);
    case DSB200_MOMENTUM: return launch_dense_update<DSB200_MOMENTUM>(ctx, a, ow, ob);
    case DSB200_ADAGRAD:  return launch_dense_update<DSB200_ADAGRAD>(ctx, a, ow, ob);
    case DSB200_NESTEROV: return launch_dense_update<DSB200_NESTEROV>(ctx, a, ow, ob);
    case DSB200_RMSPROP:  return launch_dense_update<DSB200_RMSPROP>(ctx, a, ow, ob);
    case DSB200_ADADELTA: return launch_dense_update<DSB200_ADADELTA>(ctx, a, ow, ob);
    default:              return launch_dense_update<DSB200_ADAM>(ctx, a, ow, ob);
    }
}
