// output_pass.cu -- sparse-target output layer: activation + loss + delta (rows a7, a8, fused a9).
//
// Replaces kCalculateSparse{L2,CrossEntropy,ScaledMarginalCrossEntropy,Multinomial*}Error
// (E/kLoss.cu:595-691, 1749-1980, 2213-2352, 2566-2599) and
// kCalculateSparse{,CrossEntropy,ScaledMarginalCrossEntropy}OutputDelta
// (E/kDelta.cu:2193-2618, 6533-6608, 7182-7305), and fuses kCalculateSigmoidActivation
// (E/kActivation.cu:46-64) in front of them.
//
// The reference makes six full passes over the [batch][N] output per training step (activation
// R+W, loss "Raw" R, loss "NonZero", delta "Raw" R+W, delta "NonZero").  Here ONE kernel reads Z
// once and writes delta once: a CTA owns a tile = (batch row, 2,048-column segment); it streams
// the tile with 128-bit loads, computes a = f(z), the target-is-zero loss and delta, then -- after
// a block barrier -- walks the row's target list and overwrites delta / corrects the loss at the
// non-zero targets that fall inside the tile.  The loss leaves the kernel as the reference's
// 2^30 fixed-point integer (one atomic per CTA), so its value does not depend on scheduling.
// The same kernel, with template switches, serves the stand-alone loss and delta entry points.
#include "common.cuh"
#include "launch.h"

namespace dsb {

constexpr int kOThreads = 256;
constexpr uint32_t kOSeg = 2048;

struct OArgs {
    dsb200_params P;
    dsb200_sparse S;
    int ef, act, ignoreZero;
    uint32_t position, batch, stride;
    const float* in;        // Z (DO_ACT) or activations
    float* unitOut;         // optional
    float* delta;           // optional
    unsigned long long* acc;// optional
    float slope, alpha, lambda;
    int fast;               // option "fast_math": ex2/lg2/rcp approximations (what the reference's -use_fast_math build runs)
};

// exp / log / reciprocal: accurate libdevice versions, or the MUFU approximations (relative error ~1e-6 on this
// path's value ranges, well inside the 1e-5 parity bound) which take the pass from instruction-bound to HBM-bound
__device__ __forceinline__ float dexp(float x, int fast) { return fast ? __expf(x) : expf(x); }
__device__ __forceinline__ float dlog(float x, int fast) { return fast ? __logf(x) : logf(x); }
// MUFU.RCP (1 ulp): __frcp_rn would expand to a correctly-rounded Newton sequence of ~10 instructions
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float drcp(float x, int fast) { return fast ? rcp_approx(x) : 1.0f / x; }

__device__ __forceinline__ float act_forward(int act, float z, float slope, float alpha, float lambda, int fast)
{
    switch (act) {
    case DSB200_ACT_SIGMOID: return drcp(1.0f + dexp(-z, fast), fast);            // E/kActivation.cu:53
    case DSB200_ACT_TANH:    return tanhf(z);
    case DSB200_ACT_RELU:    return fmaxf(0.0f, z);
    case DSB200_ACT_LRELU:   return fmaxf(z, z * slope);
    case DSB200_ACT_ELU:     return (z > 0.0f) ? z : alpha * (expf(z) - 1.0f);
    case DSB200_ACT_SELU:    return (z > 0.0f) ? lambda * z : lambda * alpha * (expf(z) - 1.0f);
    default:                 return z;
    }
}

// f'(x) through the activation value, as the L2 sparse delta kernels use it (E/kDelta.cu:2193-2482)
__device__ __forceinline__ float l2_deriv(int act, float a, float slope, float alpha, float lambda)
{
    switch (act) {
    case DSB200_ACT_SIGMOID: return a * (1.0f - a);
    case DSB200_ACT_TANH:    return 1.0f - a * a;
    case DSB200_ACT_RELU:    return (a > 0.0f) ? 1.0f : 0.0f;
    case DSB200_ACT_LRELU:   return (a > 0.0f) ? 1.0f : slope;
    case DSB200_ACT_ELU:     return (a > 0.0f) ? 1.0f : (a + alpha);
    case DSB200_ACT_SELU:    return (a > 0.0f) ? lambda : lambda * alpha * expf(a);
    default:                 return 1.0f;
    }
}

// target == 0 everywhere ("Raw" kernels)
template <int EF, int ACT>
__device__ __forceinline__ void raw_elem(const OArgs& a, int ef, int act, float x, float wd, float& loss, float& d)
{
    const int e = (EF >= 0) ? EF : ef, c = (ACT >= 0) ? ACT : act;
    if (e == DSB200_ERR_SMCE) {
        if (c == DSB200_ACT_SOFTMAX) {                                    // E/kDelta.cu:7229-7241 (unweighted)
            d = (x > a.P.SMCE_zeroTarget) ? a.P.SMCE_zeroScale * x : 0.0f;
        } else {                                                          // E/kLoss.cu:2215-2234, E/kDelta.cu:7184-7201
            const float w = a.P.SMCE_zeroScale * wd;
            if (x > a.P.SMCE_zeroTarget) { loss += -w * dlog(fmaxf(kMinError, 1.0f - x), a.fast); d = w * x; }
            else d = 0.0f;
        }
    } else if (e == DSB200_ERR_CROSS_ENTROPY) {
        if (c == DSB200_ACT_SOFTMAX) d = wd * x;                          // E/kDelta.cu:2484-2500
        else { loss += -wd * dlog(fmaxf(kMinError, 1.0f - x), a.fast); d = a.P.deltaBoost_zero * wd * x; }   // E/kLoss.cu:1751-1768, E/kDelta.cu:6535-6550
    } else {                                                              // L2
        loss += 0.5f * wd * x * x;                                        // E/kLoss.cu:597-615
        if (c == DSB200_ACT_SOFTMAX) d = wd * x;
        else if (c == DSB200_ACT_SIGMOID) d = a.P.deltaBoost_zero * wd * x * x * (1.0f - x);          // E/kDelta.cu:2195-2211
        else d = wd * x * l2_deriv(c, x, a.slope, a.alpha, a.lambda);
    }
}

// corrections at the non-zero targets ("NonZero" / "OnlyNonZero" kernels)
template <int EF, int ACT>
__device__ __forceinline__ void nz_elem(const OArgs& a, int ef, int act, float x, float t, float wd, float wrow,
                                        float& loss, float& d)
{
    const int e = (EF >= 0) ? EF : ef, c = (ACT >= 0) ? ACT : act;
    const bool iz = a.ignoreZero != 0;
    if (e == DSB200_ERR_SMCE) {
        if (c == DSB200_ACT_SOFTMAX) {                                    // E/kLoss.cu:2566-2599, E/kDelta.cu:7243-7267
            const float w = a.P.SMCE_oneScale * wrow;
            if (x < a.P.SMCE_oneTarget) { loss += -w * dlog(fmaxf(kMinError, x), a.fast); d = x - w; } else d = 0.0f;
        } else {                                                          // E/kLoss.cu:2236-2327, E/kDelta.cu:7203-7226
            if (!iz && x > a.P.SMCE_zeroTarget) loss += wd * a.P.SMCE_zeroScale * dlog(fmaxf(kMinError, 1.0f - x), a.fast);
            if (x < a.P.SMCE_oneTarget) { loss += -wd * a.P.SMCE_oneScale * dlog(fmaxf(kMinError, x), a.fast); d = a.P.SMCE_oneScale * wd * (x - 1.0f); }
            else d = 0.0f;
        }
    } else if (e == DSB200_ERR_CROSS_ENTROPY) {
        if (c == DSB200_ACT_SOFTMAX) { loss += -wrow * dlog(fmaxf(kMinError, x), a.fast); d = x - wrow; }     // E/kLoss.cu:1945-1967, E/kDelta.cu:2502-2521
        else {                                                            // E/kLoss.cu:1770-1839, E/kDelta.cu:6552-6571
            loss += iz ? -wd * dlog(fmaxf(kMinError, x), a.fast)
                       : wd * (-dlog(fmaxf(kMinError, x), a.fast) + dlog(fmaxf(kMinError, 1.0f - x), a.fast));
            d = a.P.deltaBoost_one * wd * (x - 1.0f);
        }
    } else {                                                              // L2: E/kLoss.cu:617-666, E/kDelta.cu:2213-2231
        const float w = 0.5f * wd;
        loss += iz ? w * ((x - t) * (x - t)) : w * ((x - t) * (x - t) - x * x);
        if (c == DSB200_ACT_SOFTMAX) d = x - wrow;
        else if (c == DSB200_ACT_SIGMOID) d = a.P.deltaBoost_one * wd * (x - t) * x * (1.0f - x);
        else d = wd * (x - t) * l2_deriv(c, x, a.slope, a.alpha, a.lambda);
    }
}

template <int EF, int ACT, bool DO_ACT, bool WRITE_UNIT, bool WRITE_DELTA, bool DO_LOSS>
__global__ void __launch_bounds__(kOThreads, 4)
output_tile_kernel(const OArgs a)
{
    __shared__ double sW[kOThreads / 32];
    const uint32_t tid = threadIdx.x;
    const uint32_t segs = (a.stride + kOSeg - 1) / kOSeg;
    const uint64_t tiles = (uint64_t)a.batch * segs;
    const bool analog = a.S.sparseData != nullptr;
    const bool raw = !a.ignoreZero;
    float loss = 0.0f;

    for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const uint32_t b = (uint32_t)(tile / segs), seg = (uint32_t)(tile % segs);
        const uint32_t c0 = seg * kOSeg, c1 = min(c0 + kOSeg, a.stride);
        const uint32_t ex = example_of(a.P, a.S.index, a.position, b);
        const float wd = a.S.dataWeight ? __ldg(a.S.dataWeight + ex) : 1.0f;
        const uint64_t g0 = (uint64_t)b * a.stride + c0, g1 = (uint64_t)b * a.stride + c1;
        const uint64_t a0u = (g0 + 3) & ~(uint64_t)3, a0 = a0u < g1 ? a0u : g1;
        const uint64_t a1d = g1 & ~(uint64_t)3, a1 = a1d > a0 ? a1d : a0;

        auto one = [&](uint64_t i) {
            float x = a.in[i];
            if (DO_ACT) x = act_forward((ACT >= 0) ? ACT : a.act, x, a.slope, a.alpha, a.lambda, a.fast);
            if (WRITE_UNIT) a.unitOut[i] = x;
            float d = 0.0f, l = 0.0f;
            if (raw) raw_elem<EF, ACT>(a, a.ef, a.act, x, wd, l, d);
            if (DO_LOSS) loss += l;
            if (WRITE_DELTA) a.delta[i] = d;
        };
        if (DO_ACT || WRITE_DELTA || (DO_LOSS && raw)) {
            for (uint64_t i = g0 + tid; i < a0; i += kOThreads) one(i);
            for (uint64_t i4 = (a0 >> 2) + tid; i4 < (a1 >> 2); i4 += kOThreads) {
                const float4 z4 = ldg_cs_f4(reinterpret_cast<const float4*>(a.in) + i4);
                float x[4] = {z4.x, z4.y, z4.z, z4.w}, d[4] = {0, 0, 0, 0};
#pragma unroll
                for (int v = 0; v < 4; v++) {
                    if (DO_ACT) x[v] = act_forward((ACT >= 0) ? ACT : a.act, x[v], a.slope, a.alpha, a.lambda, a.fast);
                    float l = 0.0f;
                    if (raw) raw_elem<EF, ACT>(a, a.ef, a.act, x[v], wd, l, d[v]);
                    if (DO_LOSS) loss += l;
                }
                if (WRITE_UNIT)  reinterpret_cast<float4*>(a.unitOut)[i4] = make_float4(x[0], x[1], x[2], x[3]);
                if (WRITE_DELTA) reinterpret_cast<float4*>(a.delta)[i4] = make_float4(d[0], d[1], d[2], d[3]);
            }
            for (uint64_t i = a1 + tid; i < g1; i += kOThreads) one(i);
        }
        if (WRITE_DELTA || WRITE_UNIT) __syncthreads();      // raw values of this tile are in place
        // non-zero targets of row b that fall into [c0, c1)
        const uint64_t rs = __ldg(a.S.sparseStart + ex), re = __ldg(a.S.sparseEnd + ex);
        const float wrow = a.S.dataWeight ? wd : 1.0f / (float)(re - rs);
        for (uint64_t j = rs + tid; j < re; j += kOThreads) {
            const uint32_t c = __ldg(a.S.sparseIndex + j);
            if (c < c0 || c >= c1) continue;
            const uint64_t i = (uint64_t)b * a.stride + c;
            float x;
            if (WRITE_UNIT) x = a.unitOut[i];
            else { x = a.in[i]; if (DO_ACT) x = act_forward((ACT >= 0) ? ACT : a.act, x, a.slope, a.alpha, a.lambda, a.fast); }
            const float t = analog ? load_value(a.S.sparseData, a.S.dataType, j) : 1.0f;
            float d = 0.0f, l = 0.0f;
            nz_elem<EF, ACT>(a, a.ef, a.act, x, t, wd, wrow, l, d);
            if (DO_LOSS) loss += l;
            if (WRITE_DELTA) a.delta[i] = d;
        }
        if (WRITE_DELTA || WRITE_UNIT) __syncthreads();      // next tile may alias nothing, but keep phases apart
    }
    if (DO_LOSS) {
        double e = warp_sum((double)loss);
        if ((tid & 31) == 0) sW[tid >> 5] = e;
        __syncthreads();
        if (tid == 0) {
            double tot = 0.0;
            for (int i = 0; i < kOThreads / 32; i++) tot += sW[i];
            if (tot != 0.0) atomicAdd(a.acc, (unsigned long long)llrint(tot * (double)kErrorScaleF));
        }
    }
}

template <bool DO_ACT, bool WRITE_UNIT, bool WRITE_DELTA, bool DO_LOSS>
static int launch_output(dsb200_ctx* ctx, const OArgs& a)
{
    const uint32_t segs = (a.stride + kOSeg - 1) / kOSeg;
    uint64_t tiles = (uint64_t)a.batch * segs;
    uint64_t grid = (uint64_t)ctx->numSMs * 8;
    if (grid > tiles) grid = tiles;
    if (grid < 1) grid = 1;
#define DSB_GO(EF, ACT) output_tile_kernel<EF, ACT, DO_ACT, WRITE_UNIT, WRITE_DELTA, DO_LOSS><<<(unsigned)grid, kOThreads, 0, ctx->stream>>>(a)
    if (a.act == DSB200_ACT_SIGMOID && a.ef == DSB200_ERR_SMCE) DSB_GO(DSB200_ERR_SMCE, DSB200_ACT_SIGMOID);
    else if (a.act == DSB200_ACT_SIGMOID && a.ef == DSB200_ERR_CROSS_ENTROPY) DSB_GO(DSB200_ERR_CROSS_ENTROPY, DSB200_ACT_SIGMOID);
    else if (a.act == DSB200_ACT_SIGMOID && a.ef == DSB200_ERR_L2) DSB_GO(DSB200_ERR_L2, DSB200_ACT_SIGMOID);
    else DSB_GO(-1, -1);
#undef DSB_GO
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    return 0;
}


// ---------------------------------------------------------------- one-pass row kernel (the training fast path)
// Boolean targets, sigmoid output, every unit counted (no SparseIgnoreZero): a CTA owns a (batch row, column
// segment) tile (<= 16,384 columns).  It first marks the row's non-zero targets of the segment in a shared-memory bitmap (4 KB), then
// streams the segment ONCE with four 128-bit loads in flight per thread: z -> a = sigmoid(z) -> loss and delta, taking
// the non-zero-target formulas where the bitmap says so.  Compared with the tile kernel above there is no second walk
// over the row (no re-read of z, no second write of delta: DRAM traffic = the 2 x 4 x batch x stride algorithmic bytes)
// and the per-tile latency chain (target list -> indices -> values) is paid once per ~109 KB instead of once per 8 KB.
constexpr uint32_t kRowSeg = 16384;
constexpr int kRowThreads = 512;

// straight-line sigmoid + target-is-zero loss / delta (same expressions as raw_elem, branches turned into selects)
template <int EF, bool FAST>
__device__ __forceinline__ float raw_sigmoid(const OArgs& a, float z, float wd, float wz, float& loss, float& d)
{
    const float x = FAST ? rcp_approx(1.0f + __expf(-z)) : 1.0f / (1.0f + expf(-z));
    if (EF == DSB200_ERR_SMCE) {                                          // wz = SMCE_zeroScale * wd
        const float lg = FAST ? __logf(fmaxf(kMinError, 1.0f - x)) : logf(fmaxf(kMinError, 1.0f - x));
        const bool on = x > a.P.SMCE_zeroTarget;
        loss += on ? -wz * lg : 0.0f;
        d = on ? wz * x : 0.0f;
    } else if (EF == DSB200_ERR_CROSS_ENTROPY) {                          // wz = deltaBoost_zero * wd
        const float lg = FAST ? __logf(fmaxf(kMinError, 1.0f - x)) : logf(fmaxf(kMinError, 1.0f - x));
        loss += -wd * lg;
        d = wz * x;
    } else {                                                              // L2, wz = deltaBoost_zero * wd
        loss += 0.5f * wd * x * x;
        d = wz * x * x * (1.0f - x);
    }
    return x;
}

template <int EF, bool WRITE_UNIT, bool DO_LOSS, bool FAST>
__global__ void __launch_bounds__(kRowThreads, 2)
output_row_kernel(const OArgs a)
{
    __shared__ uint32_t sBits[kRowSeg / 32 + 2];
    __shared__ double sW[kRowThreads / 32];
    const uint32_t tid = threadIdx.x;
    const uint32_t segs = (a.stride + kRowSeg - 1) / kRowSeg;
    const uint64_t tiles = (uint64_t)a.batch * segs;
    float loss = 0.0f;
    constexpr int ACT = DSB200_ACT_SIGMOID;

    for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const uint32_t b = (uint32_t)(tile / segs), seg = (uint32_t)(tile % segs);
        const uint32_t c0 = seg * kRowSeg, c1 = min(c0 + kRowSeg, a.stride);
        const uint32_t ex = example_of(a.P, a.S.index, a.position, b);
        const float wd = a.S.dataWeight ? __ldg(a.S.dataWeight + ex) : 1.0f;
        const float wz = (EF == DSB200_ERR_SMCE ? a.P.SMCE_zeroScale : a.P.deltaBoost_zero) * wd;
        const uint64_t rs = __ldg(a.S.sparseStart + ex), re = __ldg(a.S.sparseEnd + ex);
        for (uint32_t i = tid; i < (c1 - c0 + 31) / 32 + 2; i += kRowThreads) sBits[i] = 0u;
        __syncthreads();
        for (uint64_t j = rs + tid; j < re; j += kRowThreads) {
            const uint32_t c = __ldg(a.S.sparseIndex + j);
            if (c >= c0 && c < c1) atomicOr(&sBits[(c - c0) >> 5], 1u << ((c - c0) & 31));
        }
        __syncthreads();

        const uint64_t g0 = (uint64_t)b * a.stride + c0, g1 = (uint64_t)b * a.stride + c1;
        const uint64_t a0u = (g0 + 3) & ~(uint64_t)3, a0 = a0u < g1 ? a0u : g1;
        const uint64_t a1d = g1 & ~(uint64_t)3, a1 = a1d > a0 ? a1d : a0;
        // rare: an element whose target is non-zero (generic formulas, E/kLoss.cu / E/kDelta.cu "NonZero" kernels)
        auto nz_fix = [&](float x, float& d) {
            float l = 0.0f;
            nz_elem<EF, ACT>(a, EF, ACT, x, 1.0f, wd, 1.0f, l, d);
            if (DO_LOSS) loss += l;
        };
        auto one = [&](uint64_t i) {
            const uint32_t c = (uint32_t)(i - g0);
            float d, l = 0.0f;
            const float x = raw_sigmoid<EF, FAST>(a, a.in[i], wd, wz, l, d);
            if (DO_LOSS) loss += l;
            if ((sBits[c >> 5] >> (c & 31)) & 1u) nz_fix(x, d);
            if (WRITE_UNIT) a.unitOut[i] = x;
            a.delta[i] = d;
        };
        for (uint64_t i = g0 + tid; i < a0; i += kRowThreads) one(i);
        const float4* in4 = reinterpret_cast<const float4*>(a.in);
        const uint64_t q0 = a0 >> 2, q1 = a1 >> 2;
        for (uint64_t base = q0 + tid; base < q1; base += 4 * kRowThreads) {
            float4 z[4];
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (base + (uint64_t)u * kRowThreads < q1) z[u] = ldg_cs_f4(in4 + base + (uint64_t)u * kRowThreads);
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint64_t q = base + (uint64_t)u * kRowThreads;
                if (q >= q1) break;
                const uint32_t c = (uint32_t)((q << 2) - g0);
                const uint32_t bits = __funnelshift_r(sBits[c >> 5], sBits[(c >> 5) + 1], c & 31) & 0xFu;
                float l = 0.0f;
                float4 x, d;
                x.x = raw_sigmoid<EF, FAST>(a, z[u].x, wd, wz, l, d.x);
                x.y = raw_sigmoid<EF, FAST>(a, z[u].y, wd, wz, l, d.y);
                x.z = raw_sigmoid<EF, FAST>(a, z[u].z, wd, wz, l, d.z);
                x.w = raw_sigmoid<EF, FAST>(a, z[u].w, wd, wz, l, d.w);
                if (DO_LOSS) loss += l;
                if (bits) {
                    if (bits & 1u) nz_fix(x.x, d.x);
                    if (bits & 2u) nz_fix(x.y, d.y);
                    if (bits & 4u) nz_fix(x.z, d.z);
                    if (bits & 8u) nz_fix(x.w, d.w);
                }
                if (WRITE_UNIT) reinterpret_cast<float4*>(a.unitOut)[q] = x;
                stg_cs_f4(reinterpret_cast<float4*>(a.delta) + q, d);
            }
        }
        for (uint64_t i = a1 + tid; i < g1; i += kRowThreads) one(i);
        __syncthreads();                                     // the bitmap is rewritten by the next tile
    }
    if (DO_LOSS) {
        double e = warp_sum((double)loss);
        if ((tid & 31) == 0) sW[tid >> 5] = e;
        __syncthreads();
        if (tid == 0) {
            double tot = 0.0;
            for (int i = 0; i < kRowThreads / 32; i++) tot += sW[i];
            if (tot != 0.0) atomicAdd(a.acc, (unsigned long long)llrint(tot * (double)kErrorScaleF));
        }
    }
}

template <bool WRITE_UNIT, bool DO_LOSS>
static int launch_output_row(dsb200_ctx* ctx, const OArgs& a)
{
    const uint32_t segs = (a.stride + kRowSeg - 1) / kRowSeg;
    uint64_t tiles = (uint64_t)a.batch * segs;
    uint64_t grid = (uint64_t)ctx->numSMs * 2;
    if (grid > tiles) grid = tiles;
#define DSB_GO(EF)                                                                                                        \
    do {                                                                                                                  \
        if (a.fast) output_row_kernel<EF, WRITE_UNIT, DO_LOSS, true><<<(unsigned)grid, kRowThreads, 0, ctx->stream>>>(a);   \
        else        output_row_kernel<EF, WRITE_UNIT, DO_LOSS, false><<<(unsigned)grid, kRowThreads, 0, ctx->stream>>>(a);  \
    } while (0)
    if (a.ef == DSB200_ERR_SMCE) DSB_GO(DSB200_ERR_SMCE);
    else if (a.ef == DSB200_ERR_CROSS_ENTROPY) DSB_GO(DSB200_ERR_CROSS_ENTROPY);
    else DSB_GO(DSB200_ERR_L2);
#undef DSB_GO
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    return 0;
}

static int check_output_args(dsb200_ctx* ctx, const dsb200_sparse* s, int ef, int act, const float* in, const char* who)
{
    if (!ctx || !s || !in) return fail(ctx, DSB200_EINVAL, who);
    if (!s->sparseStart || !s->sparseEnd || !s->sparseIndex) return fail(ctx, DSB200_EINVAL, who);
    if (ef != DSB200_ERR_L2 && ef != DSB200_ERR_CROSS_ENTROPY && ef != DSB200_ERR_SMCE)
        return fail(ctx, DSB200_EUNSUPPORTED, "sparse output: error function outside the hot path (L2, CrossEntropy, ScaledMarginalCrossEntropy)");
    if (s->sparseData && ef != DSB200_ERR_L2)
        return fail(ctx, DSB200_EUNSUPPORTED, "sparse output: analog targets are supported for L2 only");
    (void)act;
    return 0;
}

static OArgs make_oargs(dsb200_ctx* ctx, const dsb200_sparse* s, int ef, int act, uint32_t position, uint32_t batch,
                        uint32_t stride, const float* in, int ignoreZero, float slope, float alpha, float lambda)
{
    OArgs a{};
    a.P = ctx->params; a.S = *s; a.ef = ef; a.act = act; a.ignoreZero = ignoreZero;
    a.position = position; a.batch = batch; a.stride = stride; a.in = in;
    a.slope = slope; a.alpha = alpha; a.lambda = lambda;
    a.fast = ctx->fastMath;
    return a;
}

}  // namespace dsb

extern "C" {

int dsb200_sparse_loss_async(dsb200_ctx* ctx, const dsb200_sparse* s, int ef, int act, uint32_t position, uint32_t batch,
                             uint32_t stride, const float* pUnit, int ignoreZero, unsigned long long* pDevAcc)
{
    DSB_PROFILE(ctx, "sparse_loss_async");
    using namespace dsb;
    int rc = check_output_args(ctx, s, ef, act, pUnit, "sparse_loss: null argument");
    if (rc) return rc;
    if (!pDevAcc) return fail(ctx, DSB200_EINVAL, "sparse_loss: null accumulator");
    if (!batch || !stride) return 0;
    // multinomial (softmax) variants only have the non-zero term (E/kLoss.cu:1945-1967, 2566-2599)
    const int iz = (act == DSB200_ACT_SOFTMAX && ef != DSB200_ERR_L2) ? 1 : ignoreZero;
    OArgs a = make_oargs(ctx, s, ef, act, position, batch, stride, pUnit, iz, 0.0f, 0.0f, 0.0f);
    a.acc = pDevAcc;
    return launch_output<false, false, false, true>(ctx, a);
}

int dsb200_sparse_loss(dsb200_ctx* ctx, const dsb200_sparse* s, int ef, int act, uint32_t position, uint32_t batch,
                       uint32_t stride, const float* pUnit, int ignoreZero, float* pLossOut)
{
    using namespace dsb;
    if (!ctx || !pLossOut) return fail(ctx, DSB200_EINVAL, "sparse_loss: null argument");
    DSB_CUDA_OK(cudaMemsetAsync(ctx->dAccumulator, 0, sizeof(unsigned long long), ctx->stream));   // E/kLoss.cu:2331
    int rc = dsb200_sparse_loss_async(ctx, s, ef, act, position, batch, stride, pUnit, ignoreZero, ctx->dAccumulator);
    if (rc) return rc;
    DSB_CUDA_OK(cudaMemcpyAsync(ctx->hAccumulator, ctx->dAccumulator, sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    DSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    *pLossOut = (float)((double)(long long)ctx->hAccumulator[0] * kOneOverErrorScale);             // E/kLoss.cu:2349-2351
    return 0;
}

int dsb200_sparse_output_delta(dsb200_ctx* ctx, const dsb200_sparse* s, int ef, int act, uint32_t position, uint32_t batch,
                               uint32_t stride, const float* pUnit, float* pDelta, int ignoreZero,
                               float slope, float alpha, float lambda)
{
    DSB_PROFILE(ctx, "sparse_output_delta");
    using namespace dsb;
    int rc = check_output_args(ctx, s, ef, act, pUnit, "sparse_output_delta: null argument");
    if (rc) return rc;
    if (!pDelta) return fail(ctx, DSB200_EINVAL, "sparse_output_delta: null delta");
    if (!batch || !stride) return 0;
    OArgs a = make_oargs(ctx, s, ef, act, position, batch, stride, pUnit, ignoreZero, slope, alpha, lambda);
    a.delta = pDelta;
    return launch_output<false, false, true, false>(ctx, a);
}

int dsb200_output_pass(dsb200_ctx* ctx, const dsb200_sparse* s, int ef, int act, uint32_t position, uint32_t batch,
                       uint32_t stride, const float* pZ, float* pUnitOut, float* pDelta, unsigned long long* pDevAcc)
{
    DSB_PROFILE(ctx, "output_pass");
    using namespace dsb;
    int rc = check_output_args(ctx, s, ef, act, pZ, "output_pass: null argument");
    if (rc) return rc;
    if (!pDelta) return fail(ctx, DSB200_EINVAL, "output_pass: null delta");
    if (act == DSB200_ACT_SOFTMAX)
        return fail(ctx, DSB200_EUNSUPPORTED, "output_pass: softmax needs dsb200_activation first, then dsb200_sparse_loss/_output_delta");
    if (!batch || !stride) return 0;
    OArgs a = make_oargs(ctx, s, ef, act, position, batch, stride, pZ, 0, 0.0f, 0.0f, 0.0f);
    a.unitOut = pUnitOut; a.delta = pDelta; a.acc = pDevAcc;
    if (act == DSB200_ACT_SIGMOID && !s->sparseData && !ctx->outputTileKernel) {          // Boolean targets: the one-pass row kernel
        if (pUnitOut) return pDevAcc ? launch_output_row<true, true>(ctx, a) : launch_output_row<true, false>(ctx, a);
        return pDevAcc ? launch_output_row<false, true>(ctx, a) : launch_output_row<false, false>(ctx, a);
    }
    if (pUnitOut) return pDevAcc ? launch_output<true, true, true, true>(ctx, a) : launch_output<true, true, true, false>(ctx, a);
    return pDevAcc ? launch_output<true, false, true, true>(ctx, a) : launch_output<true, false, true, false>(ctx, a);
}

}  // extern "C"
