// dense_small.cu -- exact-fp32 SIMT kernels for the dense layers of the path (hot-path row a11).
//
// Two jobs:
//   1. the SMALL dense layers in every arithmetic mode: the 128 x 128 hidden weights of BASELINE config 2 cover eight 128 x 128
//      tiles -- too few for the persistent tcgen05 kernels (gemm_tc.cu, gemm_stream.cu), and pure latency in a library SGEMM, where
//      each layer costs three launches going forward (kClearUnit + cublasSgemm + activation, E/NNLayer.cpp:1009, 1073, 1157), two
//      going backward (cublasSgemm + kCalculateHadamardProduct, E/NNLayer.cpp:2274, 2137) and four for the update (cublasSgemm with
//      its split-K reduce, k*UpdateWeights, k*UpdateBiases, E/NNLayer.cpp:2223, E/NNWeight.cpp:729-794);
//   2. every dense product of DSB200_GEMM_FP32 (the 1e-5 parity mode), in place of cublasSgemm: no library call is left on the path.
//
//   dsb200_gemm_fwd_bias_act   C[B][n]  = act(A[B][k] * W[k][n] + bias[n])                              FORM_NN, one launch
//   dsb200_gemm_dx_hadamard    Dp[B][k] = (D[B][n] * W[k][n]^T) (.) f'(unit[B][k]) * s                  FORM_NT, one launch
//   dsb200_dense_update        W, bias <- optimizer(X[B][k]^T * D[B][n], column sums of D)              FORM_TN, one launch
//   dsb200_gemm_fwd / _dw / _dx (FP32 mode, small shapes)                                               plain alpha / beta epilogue
//
// Round-2 launch list (profiles/r2_launches.md): the first version of these kernels (32 x 64 tiles, 2 x 2 outputs per thread, 64
// CTAs, every CTA re-reading its operands through one __syncthreads-separated chunk loop) ran 10.5 / 12.2 / 19.7 us per launch, six
// launches = 85 us of a 366 us step for 0.2 GFLOP.  They are latency, not throughput: the design below spends its parallelism on
// the critical path of ONE tile.
//   * CTA = 4 warps, tile = 32 rows x 32 columns (a 2-D tile: an 8 x 128 strip made all 128 CTAs pull the same 64 KB of W through
//     the same L2 slices at the same time -- ncu: 43 % of the samples in the wait for the staged operand); the contraction runs in
//     chunks of 128, and inside a chunk every warp takes its own QUARTER (32 k) for the whole tile -- so a warp stages only what it
//     alone reads (cp.async, 16 bytes per request, the whole quarter in flight at once) and the chunk loop has no block barrier;
//   * thread tile 8 rows x 4 columns: per k two LDS.128 for the row values (amortised), one LDS.128 for the columns, 32 FFMA -- FMA
//     bound, not shared-memory bound; layouts chosen so that every shared-memory access is conflict free;
//   * the four partial tiles meet once, through shared memory, and are added in a fixed order (deterministic); what the epilogue
//     reads from global memory (bias, activations for the Hadamard product, weights and optimizer state) is requested before that;
//   * FORM_TN contracts over the batch: the batch is additionally split over gridDim.z so that a 128 x 128 weight still fills the GPU
//     (20 tiles x 8 segments = 160 CTAs); segment partials go through a workspace, the LAST CTA to arrive (self-resetting
//     counter) adds them in segment order and applies the optimizer rule -- still one launch, still deterministic.  The bias is one
//     more row of the same product (X column of ones).
#include <algorithm>

#include "common.cuh"
#include "launch.h"
#include "optimizer.cuh"

namespace dsb {
namespace dsm {

constexpr int TM = 32, TN = 32, KC = 128, KW = 32, WARPS = KC / KW, THREADS = WARPS * 32;
constexpr int PA = 36;                                                        // pitch of the [row][k] A quarter (floats): rows rg + 4 i conflict free
constexpr int SB_FLOATS = KW * TN, SA_FLOATS = TM * PA;
constexpr int SMEM_FLOATS = WARPS * (SB_FLOATS + SA_FLOATS);                  // 34,816 bytes; the reduction tiles (16 KB) alias the front
static_assert(WARPS * TM * TN <= SMEM_FLOATS, "reduction tiles must fit");

enum { FORM_NN = 0, FORM_NT = 1, FORM_TN = 2 };
enum { EPI_PLAIN = 0, EPI_BIAS_ACT = 1, EPI_HADAMARD = 2, EPI_UPDATE = 3 };

struct Args {
    const float* A; uint32_t lda;       // NN / NT: A[M][K];  TN: A[K][aCols] (row i of the product = column i of A; row aCols = ones when `ones`)
    const float* B; uint32_t ldb;       // NN / TN: B[K][N];  NT: B[N][K]
    float* C; uint32_t ldc;
    uint32_t M, N, K;
    uint32_t aCols; int ones;
    uint32_t segs, chunksPerSeg;        // the contraction split over gridDim.z (TN only)
    float* partial; uint32_t* counters; // segs > 1: [segs][M][N] partial tiles, one self-resetting arrival counter per tile
    float alpha, beta;
    const float* bias; const float* unit; int act; float scale, slope, ealpha, lambda;
    // EPI_UPDATE: C = weights [aCols][N]
    float galpha, invBatch;
    float* V; float* GV; float* bvec; float* bV; float* bGV;
};

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ float act_of(int act, float z, float slope, float alpha, float lambda)
{
    switch (act) {
    case DSB200_ACT_SIGMOID: return __fdividef(1.0f, 1.0f + __expf(-z));   // MUFU ex2 / rcp (2^-22 relative; what the reference's -use_fast_math build runs)
    case DSB200_ACT_TANH:    return tanhf(z);
    case DSB200_ACT_RELU:    return fmaxf(0.0f, z);
    case DSB200_ACT_LRELU:   return fmaxf(z, z * slope);
    case DSB200_ACT_ELU:     return (z > 0.0f) ? z : alpha * (expf(z) - 1.0f);
    case DSB200_ACT_SELU:    return (z > 0.0f) ? lambda * z : lambda * alpha * (expf(z) - 1.0f);
    default:                 return z;
    }
}
// f'(x) through the activation value, times the incoming delta (kCalculateHadamardProduct, E/kDelta.cu:9021-9151)
__device__ __forceinline__ float hadamard_of(int act, float x, float d, float scale, float slope, float alpha, float lambda)
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

template <int FORM, int EPI, int MODE>
__global__ void __launch_bounds__(THREADS)
dense_kernel(const Args a, const OptArgs ow, const OptArgs ob)
{
    __shared__ __align__(16) float smem[SMEM_FLOATS];
    __shared__ uint32_t sLast;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31, rg = lane >> 3, cg = lane & 7;
    float* sB = smem + warp * (SB_FLOATS + SA_FLOATS);                         // [32 k][32 n]
    float* sA = sB + SB_FLOATS;                                                // NN / NT: [32 r][PA] (k contiguous);  TN: [32 k][32 r]
    const uint32_t m0 = blockIdx.x * TM, n0 = blockIdx.y * TN, seg = blockIdx.z;
    const uint32_t sBaddr = smem_u32(sB), sAaddr = smem_u32(sA);
    pdl_launch_dependents();
    pdl_wait();

    // whole-tile facts that decide between 16-byte asynchronous copies and guarded scalar loads
    const bool colsIn = n0 + TN <= a.N;
    const bool bVec = (a.ldb & 3u) == 0 && (((uintptr_t)a.B) & 15u) == 0;
    const bool aVec = (a.lda & 3u) == 0 && (((uintptr_t)a.A) & 15u) == 0;
    const bool rowsIn = FORM == FORM_TN ? m0 + TM <= a.aCols : m0 + TM <= a.M;

    // what the epilogue will read, requested now: thread t finishes row t / 4, columns 8 (t % 4) .. + 7
    const uint32_t em = m0 + (threadIdx.x >> 2), en = n0 + 8 * (threadIdx.x & 3);
    const bool eVec = em < a.M && en + 8 <= a.N && (a.ldc & 3u) == 0 && (((uintptr_t)a.C) & 15u) == 0;
    float pre0[8], pre1[8], pre2[8];                                           // bias | unit | (C, V, GV)
#pragma unroll
    for (int j = 0; j < 8; j++) { pre0[j] = 0.0f; pre1[j] = 0.0f; pre2[j] = 0.0f; }
    if (em < a.M) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (en + j >= a.N) continue;
            if (EPI == EPI_BIAS_ACT) pre0[j] = __ldg(a.bias + en + j);
            if (EPI == EPI_HADAMARD) pre0[j] = __ldg(a.unit + (size_t)em * a.ldc + en + j);
            if (EPI == EPI_PLAIN && a.beta != 0.0f) pre0[j] = a.C[(size_t)em * a.ldc + en + j];
            if (EPI == EPI_UPDATE) {
                if (em < a.aCols) {
                    const size_t e = (size_t)em * a.ldc + en + j;
                    pre0[j] = a.C[e];
                    if (opt_uses_v(MODE)) pre1[j] = a.V[e];
                    if (opt_uses_gv(MODE)) pre2[j] = a.GV[e];
                } else {
                    pre0[j] = a.bvec[en + j];
                    if (opt_uses_v(MODE)) pre1[j] = a.bV[en + j];
                    if (opt_uses_gv(MODE)) pre2[j] = a.bGV[en + j];
                }
            }
        }
    }

    float acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) acc[r][c] = 0.0f;

    const uint32_t chunk0 = seg * a.chunksPerSeg, chunks = (a.K + KC - 1) / KC;
    const uint32_t chunk1 = min(chunks, chunk0 + a.chunksPerSeg);
    for (uint32_t ch = chunk0; ch < chunk1; ch++) {
        const uint32_t k0 = ch * KC + warp * KW;                               // this warp's quarter: [k0, k0 + 32)
        if (k0 >= a.K) break;                                                  // (warp-uniform)
        const bool kIn = k0 + KW <= a.K;
        // ---- B quarter -> sB[kk][nn]
        if (FORM != FORM_NT) {
            if (colsIn && kIn && bVec) {
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    const uint32_t id = lane + 32 * t, kk = id >> 3, j = id & 7;
                    cp_async16(sBaddr + (kk * TN + 4 * j) * 4, a.B + (size_t)(k0 + kk) * a.ldb + n0 + 4 * j);
                }
            } else {
                for (int kk = 0; kk < KW; kk++)
                    sB[kk * TN + lane] = (k0 + kk < a.K && n0 + lane < a.N) ? __ldg(a.B + (size_t)(k0 + kk) * a.ldb + n0 + lane) : 0.0f;
            }
        } else {
            if (colsIn && kIn && bVec) {                                       // lane = column: 32 consecutive k of its row of B
                const float4* src = reinterpret_cast<const float4*>(a.B + (size_t)(n0 + lane) * a.ldb + k0);
                float4 v[8];
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] = __ldg(src + j);
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    float* d = sB + (4 * j) * TN + lane;
                    d[0] = v[j].x; d[TN] = v[j].y; d[2 * TN] = v[j].z; d[3 * TN] = v[j].w;
                }
            } else {
                for (int kk = 0; kk < KW; kk++)
                    sB[kk * TN + lane] = (k0 + kk < a.K && n0 + lane < a.N) ? __ldg(a.B + (size_t)(n0 + lane) * a.ldb + k0 + kk) : 0.0f;
            }
        }
        // ---- A quarter
        if (FORM != FORM_TN) {                                                 // sA[r][kk] = A[m0 + r][k0 + kk]
            if (rowsIn && kIn && aVec) {
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    const uint32_t id = lane + 32 * t, r = id >> 3, j = id & 7;
                    cp_async16(sAaddr + (r * PA + 4 * j) * 4, a.A + (size_t)(m0 + r) * a.lda + k0 + 4 * j);
                }
            } else {
                for (int r = 0; r < TM; r++)
                    sA[r * PA + lane] = (m0 + r < a.M && k0 + lane < a.K) ? __ldg(a.A + (size_t)(m0 + r) * a.lda + k0 + lane) : 0.0f;
            }
        } else {                                                               // sA[kk][r] = A[k0 + kk][m0 + r], the ones row at r = aCols
            if (rowsIn && kIn && aVec) {
#pragma unroll
                for (int t = 0; t < 8; t++) {
                    const uint32_t id = lane + 32 * t, kk = id >> 3, j = id & 7;
                    cp_async16(sAaddr + (kk * TM + 4 * j) * 4, a.A + (size_t)(k0 + kk) * a.lda + m0 + 4 * j);
                }
            } else {
                const uint32_t i = m0 + lane;
                for (int kk = 0; kk < KW; kk++) {
                    float v = 0.0f;
                    if (k0 + kk < a.K) v = i < a.aCols ? __ldg(a.A + (size_t)(k0 + kk) * a.lda + i) : ((i == a.aCols && a.ones) ? 1.0f : 0.0f);
                    sA[kk * TM + lane] = v;
                }
            }
        }
        cp_async_wait_all();
        __syncwarp();
        // ---- 32 k of the tile.  Thread tile: columns 4 cg .. + 3; rows rg + 4 i (NN / NT) or 8 rg + i (TN), i < 8
        if (FORM != FORM_TN) {
#pragma unroll 2
            for (int k4 = 0; k4 < KW / 4; k4++) {
                float4 av[8];
#pragma unroll
                for (int i = 0; i < 8; i++) av[i] = *reinterpret_cast<const float4*>(sA + (rg + 4 * i) * PA + 4 * k4);
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float4 b = *reinterpret_cast<const float4*>(sB + (4 * k4 + j) * TN + 4 * cg);
#pragma unroll
                    for (int i = 0; i < 8; i++) {
                        const float x = j == 0 ? av[i].x : j == 1 ? av[i].y : j == 2 ? av[i].z : av[i].w;
                        acc[i][0] = fmaf(x, b.x, acc[i][0]); acc[i][1] = fmaf(x, b.y, acc[i][1]);
                        acc[i][2] = fmaf(x, b.z, acc[i][2]); acc[i][3] = fmaf(x, b.w, acc[i][3]);
                    }
                }
            }
        } else {
#pragma unroll 8
            for (int kk = 0; kk < KW; kk++) {
                const float4 a0 = *reinterpret_cast<const float4*>(sA + kk * TM + 8 * rg), a1 = *reinterpret_cast<const float4*>(sA + kk * TM + 8 * rg + 4);
                const float4 b = *reinterpret_cast<const float4*>(sB + kk * TN + 4 * cg);
                const float x[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    acc[i][0] = fmaf(x[i], b.x, acc[i][0]); acc[i][1] = fmaf(x[i], b.y, acc[i][1]);
                    acc[i][2] = fmaf(x[i], b.z, acc[i][2]); acc[i][3] = fmaf(x[i], b.w, acc[i][3]);
                }
            }
        }
        __syncwarp();                                                          // the quarter may be overwritten by the next chunk
    }

    // ---- the four quarter sums meet: red[w][row][col]
    __syncthreads();
    float* red = smem;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t row = FORM != FORM_TN ? rg + 4 * i : 8 * rg + i;
        *reinterpret_cast<float4*>(red + (warp * TM + row) * TN + 4 * cg) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    }
    __syncthreads();
    float s[8];
    {
        const uint32_t o = (threadIdx.x >> 2) * TN + 8 * (threadIdx.x & 3);
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const float4 p0 = *reinterpret_cast<const float4*>(red + o + 4 * h), p1 = *reinterpret_cast<const float4*>(red + TM * TN + o + 4 * h);
            const float4 p2 = *reinterpret_cast<const float4*>(red + 2 * TM * TN + o + 4 * h), p3 = *reinterpret_cast<const float4*>(red + 3 * TM * TN + o + 4 * h);
            s[4 * h] = ((p0.x + p1.x) + p2.x) + p3.x; s[4 * h + 1] = ((p0.y + p1.y) + p2.y) + p3.y;
            s[4 * h + 2] = ((p0.z + p1.z) + p2.z) + p3.z; s[4 * h + 3] = ((p0.w + p1.w) + p2.w) + p3.w;
        }
    }
    if (a.segs > 1) {
        // ---- segment partials -> workspace; the last CTA of the tile adds them in segment order
        const size_t plane = (size_t)a.M * a.N;
        if (em < a.M) {
#pragma unroll
            for (int j = 0; j < 8; j++)
                if (en + j < a.N) __stcg(a.partial + seg * plane + (size_t)em * a.N + en + j, s[j]);
        }
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t* ctr = a.counters + blockIdx.y * gridDim.x + blockIdx.x;
            const uint32_t old = atomicAdd(ctr, 1u);
            sLast = old == a.segs - 1;
            if (sLast) *ctr = 0u;                                              // ready for the next launch
        }
        __syncthreads();
        if (!sLast) return;
        __threadfence();
#pragma unroll
        for (int j = 0; j < 8; j++) s[j] = 0.0f;
        if (em < a.M)
            for (uint32_t z = 0; z < a.segs; z++) {
#pragma unroll
                for (int j = 0; j < 8; j++)
                    if (en + j < a.N) s[j] += __ldcg(a.partial + z * plane + (size_t)em * a.N + en + j);
            }
    }
    // ---- epilogue: row em, columns en .. en + 7
    if (em >= a.M) return;
    float out[8];
    if (EPI == EPI_UPDATE) {
        float vOut[8], gvOut[8];
        const bool wRow = em < a.aCols;                                        // else the ones row: column sums of D -> bias rules (E/NNWeight.cpp:760-794)
#pragma unroll
        for (int j = 0; j < 8; j++) {
            float vv = pre1[j], ss = pre2[j];
            out[j] = wRow ? opt_weight<MODE>(ow, a.galpha * s[j], pre0[j], vv, ss) : opt_bias<MODE>(ob, s[j] * a.invBatch, pre0[j], vv, ss);
            vOut[j] = vv; gvOut[j] = ss;
        }
        float* w = wRow ? a.C + (size_t)em * a.ldc + en : a.bvec + en;
        float* v = wRow ? a.V + (size_t)em * a.ldc + en : a.bV + en;
        float* gv = wRow ? a.GV + (size_t)em * a.ldc + en : a.bGV + en;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            if (en + j >= a.N) continue;
            w[j] = out[j];
            if (opt_uses_v(MODE)) v[j] = vOut[j];
            if (opt_uses_gv(MODE)) gv[j] = gvOut[j];
        }
        return;
    }
#pragma unroll
    for (int j = 0; j < 8; j++) {
        if (EPI == EPI_PLAIN) out[j] = a.beta != 0.0f ? fmaf(a.beta, pre0[j], a.alpha * s[j]) : a.alpha * s[j];
        else if (EPI == EPI_BIAS_ACT) out[j] = act_of(a.act, s[j] + pre0[j], a.slope, a.ealpha, a.lambda);
        else out[j] = hadamard_of(a.act, pre0[j], s[j], a.scale, a.slope, a.ealpha, a.lambda);
    }
    float* o = a.C + (size_t)em * a.ldc + en;
    if (eVec) {
        *reinterpret_cast<float4*>(o) = make_float4(out[0], out[1], out[2], out[3]);
        *reinterpret_cast<float4*>(o + 4) = make_float4(out[4], out[5], out[6], out[7]);
    } else {
#pragma unroll
        for (int j = 0; j < 8; j++) if (en + j < a.N) o[j] = out[j];
    }
}

// workspace of the segmented TN form: partial tiles (grow only) and zeroed, self-resetting counters
static int reserve_segments(dsb200_ctx* ctx, size_t partialFloats, size_t tiles, float** partial, uint32_t** counters)
{
    const size_t ctrBytes = ((tiles * sizeof(uint32_t)) + 255) & ~(size_t)255;
    const size_t need = ctrBytes + partialFloats * sizeof(float);
    if (need > ctx->denseWsBytes || tiles > ctx->denseWsTiles) {
        if (ctx->dDenseWs) { DSB_CUDA_OK(cudaStreamSynchronize(ctx->stream)); DSB_CUDA_OK(cudaFree(ctx->dDenseWs)); ctx->dDenseWs = nullptr; ctx->denseWsBytes = 0; }
        const size_t tilesCap = std::max<size_t>(tiles, 1024), ctrCap = ((tilesCap * sizeof(uint32_t)) + 255) & ~(size_t)255;
        const size_t bytes = ctrCap + partialFloats * sizeof(float) + partialFloats / 2;
        DSB_CUDA_OK(cudaMalloc(&ctx->dDenseWs, bytes));
        DSB_CUDA_OK(cudaMemsetAsync(ctx->dDenseWs, 0, ctrCap, ctx->stream));
        ctx->denseWsBytes = bytes; ctx->denseWsTiles = tilesCap;
    }
    const size_t ctrCap = ((ctx->denseWsTiles * sizeof(uint32_t)) + 255) & ~(size_t)255;
    *counters = reinterpret_cast<uint32_t*>(ctx->dDenseWs);
    *partial = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(ctx->dDenseWs) + ctrCap);
    return 0;
}

template <int FORM, int EPI, int MODE>
static int launch(dsb200_ctx* ctx, Args& a, const OptArgs& ow, const OptArgs& ob)
{
    const uint32_t rowBlocks = (a.M + TM - 1) / TM, colTiles = (a.N + TN - 1) / TN, chunks = (a.K + KC - 1) / KC;
    if (colTiles > 65535u) return fail(ctx, DSB200_EUNSUPPORTED, "dense kernel: more than 2,097,120 columns");
    a.segs = 1; a.chunksPerSeg = chunks; a.partial = nullptr; a.counters = nullptr;
    if (FORM == FORM_TN && chunks > 1) {
        // fill about two waves of CTAs; never more than 64 segments
        const uint64_t tiles = (uint64_t)rowBlocks * colTiles;
        uint32_t segs = (uint32_t)std::min<uint64_t>(std::min<uint32_t>(chunks, 64u), std::max<uint64_t>(1, (2ull * ctx->numSMs) / tiles));
        a.chunksPerSeg = (chunks + segs - 1) / segs;
        segs = (chunks + a.chunksPerSeg - 1) / a.chunksPerSeg;
        if (segs > 1) {
            a.segs = segs;
            const int rc = reserve_segments(ctx, (size_t)segs * a.M * a.N, (size_t)tiles, &a.partial, &a.counters);
            if (rc) return rc;
        }
    }
    dim3 grid(rowBlocks, colTiles, a.segs);
    DSB_CUDA_OK(launch_pdl(dense_kernel<FORM, EPI, MODE>, grid, dim3(THREADS), 0, ctx->stream, a, ow, ob));
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace dsm

// C[M][N] = alpha * op(A) * op(B) + beta * C in exact fp32: the three products of a dense layer
//   form 0: A[M][K] * B[K][N]      (forward, E/NNLayer.cpp:1073)
//   form 1: A[M][K] * B[N][K]^T    (input delta, E/NNLayer.cpp:2274)
//   form 2: A[K][M]^T * B[K][N]    (weight gradient, E/NNLayer.cpp:2223)
int dense_gemm(dsb200_ctx* ctx, int form, uint32_t M, uint32_t N, uint32_t K, float alpha, const float* A, uint32_t lda, const float* B, uint32_t ldb,
               float beta, float* C, uint32_t ldc)
{
    using namespace dsm;
    Args a{};
    a.A = A; a.lda = lda; a.B = B; a.ldb = ldb; a.C = C; a.ldc = ldc; a.M = M; a.N = N; a.K = K; a.aCols = M; a.ones = 0; a.alpha = alpha; a.beta = beta;
    const OptArgs o{};
    if (form == FORM_NN) return launch<FORM_NN, EPI_PLAIN, 0>(ctx, a, o, o);
    if (form == FORM_NT) return launch<FORM_NT, EPI_PLAIN, 0>(ctx, a, o, o);
    return launch<FORM_TN, EPI_PLAIN, 0>(ctx, a, o, o);
}

int dense_small_fwd(dsb200_ctx* ctx, uint32_t B, uint32_t k, uint32_t n, const float* A, const float* W, const float* bias, int act, float* C,
                    float slope, float alpha, float lambda)
{
    using namespace dsm;
    Args a{};
    a.A = A; a.lda = k; a.B = W; a.ldb = n; a.C = C; a.ldc = n; a.M = B; a.N = n; a.K = k; a.aCols = B;
    a.bias = bias; a.act = act; a.slope = slope; a.ealpha = alpha; a.lambda = lambda;
    const OptArgs o{};
    return launch<FORM_NN, EPI_BIAS_ACT, 0>(ctx, a, o, o);
}

int dense_small_dx_hadamard(dsb200_ctx* ctx, uint32_t B, uint32_t k, uint32_t n, const float* D, const float* W, int activation, float scale,
                            const float* pUnit, float* Dp, float slope, float alpha, float lambda)
{
    using namespace dsm;
    Args a{};
    a.A = D; a.lda = n; a.B = W; a.ldb = n; a.C = Dp; a.ldc = k; a.M = B; a.N = k; a.K = n; a.aCols = B;
    a.unit = pUnit; a.act = activation; a.scale = scale; a.slope = slope; a.ealpha = alpha; a.lambda = lambda;
    const OptArgs o{};
    return launch<FORM_NT, EPI_HADAMARD, 0>(ctx, a, o, o);
}

}  // namespace dsb

// =====================================================================================================================
// dsb200_dense_update: weight gradient + optimizer step + bias update of a dense layer in ONE launch (see the head of the file).
//   g[i][j]  = galpha * sum_b X[b][i] * D[b][j]          (never written)
//   W[i][j]  = opt_weight(g, W, V, GV)                    (optimizer.cuh, the rules of E/kernels.cu:2746-3199)
//   bias[j]  = opt_bias(sum_b D[b][j] / B, bias, ...)
// =====================================================================================================================
extern "C" int dsb200_dense_update(dsb200_ctx* ctx, int mode, uint32_t B, uint32_t k, uint32_t n, float galpha, const float* X, const float* D,
                                   float alpha, float lambda, float lambda1, float mu, float mu1, float t, float* pWeightVelocity,
                                   float* pWeightGradientVelocity, float* pWeight, float* pBiasVelocity, float* pBiasGradientVelocity, float* pBias)
{
    DSB_PROFILE_T(ctx, "dense_update", (unsigned long long)k * n);
    using namespace dsb;
    using namespace dsb::dsm;
    if (!ctx || !X || !D || !pWeight) return fail(ctx, DSB200_EINVAL, "dense_update: null argument");
    if (mode < 0 || mode > DSB200_ADAM) return fail(ctx, DSB200_EINVAL, "dense_update: bad mode");
    if (opt_uses_v(mode) && (!pWeightVelocity || (pBias && !pBiasVelocity))) return fail(ctx, DSB200_EINVAL, "dense_update: velocity buffer missing");
    if (opt_uses_gv(mode) && (!pWeightGradientVelocity || (pBias && !pBiasGradientVelocity))) return fail(ctx, DSB200_EINVAL, "dense_update: gradient-velocity buffer missing");
    if (!B || !k || !n) return 0;
    Args a{};
    a.A = X; a.lda = k; a.B = D; a.ldb = n; a.C = pWeight; a.ldc = n; a.M = k + (pBias ? 1u : 0u); a.N = n; a.K = B; a.aCols = k; a.ones = pBias ? 1 : 0;
    a.galpha = galpha; a.invBatch = 1.0f / (float)B;
    a.V = pWeightVelocity; a.GV = pWeightGradientVelocity; a.bvec = pBias; a.bV = pBiasVelocity; a.bGV = pBiasGradientVelocity;
    const OptArgs ow = make_opt(mode, alpha, lambda, lambda1, mu, mu1, t), ob = make_opt(mode, alpha, 0.0f, 0.0f, mu, mu1, t);
    switch (mode) {
    case DSB200_SGD:      return launch<FORM_TN, EPI_UPDATE, DSB200_SGD>(ctx, a, ow, ob);
    case DSB200_MOMENTUM: return launch<FORM_TN, EPI_UPDATE, DSB200_MOMENTUM>(ctx, a, ow, ob);
    case DSB200_ADAGRAD:  return launch<FORM_TN, EPI_UPDATE, DSB200_ADAGRAD>(ctx, a, ow, ob);
    case DSB200_NESTEROV: return launch<FORM_TN, EPI_UPDATE, DSB200_NESTEROV>(ctx, a, ow, ob);
    case DSB200_RMSPROP:  return launch<FORM_TN, EPI_UPDATE, DSB200_RMSPROP>(ctx, a, ow, ob);
    case DSB200_ADADELTA: return launch<FORM_TN, EPI_UPDATE, DSB200_ADADELTA>(ctx, a, ow, ob);
    default:              return launch<FORM_TN, EPI_UPDATE, DSB200_ADAM>(ctx, a, ow, ob);
    }
}
