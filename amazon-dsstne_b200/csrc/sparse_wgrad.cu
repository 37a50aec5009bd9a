// sparse_wgrad.cu -- sparse weight gradient of the input layer (hot-path row a6), optionally
// fused with the optimizer step (a12).
//
// Replaces kCalculateSparseTransposed[Analog]WeightGradient (E/kernels.cu:2537-2692):
//     dW[c,:] = beta*dW[c,:] + alpha*q * sum_{e in column c} (tdata_e *) delta[row_e,:]
// with the sum taken exactly as the reference takes it: every term converted with
// llrintf(2^30 * x) and added in int64, which makes the result independent of the order of the
// entries inside a column (so the transposed matrix need not be sorted) and bit-identical to
// the CPU oracle.
//
// Reference launch: one 256-thread block per input unit (27,278 / 1M tiny blocks), each lane
// summing one output column over all entries with a single dependent chain.  Here:
//  * persistent grid (multiple of the SM count); a CTA walks tiles of G columns, one column per
//    thread group (a warp for stride 128: 32 lanes x float4 = one 512-byte delta row per load);
//  * 8 independent delta-row gathers in flight per thread (ld.global.nc.v4, L1 bypass);
//  * columns with more than kHeavy entries (the Zipf head: up to `batch` entries) are not left
//    to one group: all G groups of the CTA split the entries and add their int64 partial sums
//    through shared memory (exact, so still order independent);
//  * |x| < 2 takes the 32-bit convert (F2I.S32) -- identical integer, 4x the throughput of the
//    64-bit convert;
//  * FUSED variant: the finished gradient row is fed straight into the optimizer rule and the
//    weight (and state) row is updated in place -- dW is never written or re-read.
#include "optimizer.cuh"
#include "launch.h"

namespace dsb {

constexpr int kGThreads = 256;
constexpr int kGUnroll  = 8;
constexpr uint32_t kHeavy = 96;
constexpr uint32_t kHeavy2 = 32;                // two-kernel scheme: columns above this go to the CTA-per-column kernel

struct GArgs {
    float alpha, beta;          // alpha already multiplied by q
    uint32_t m, n;
    const uint32_t* tStart; const uint32_t* tEnd; const uint32_t* tIndex; const float* tData;
    const float* delta;
    float* dW;
    // fused optimizer
    OptArgs opt;
    float* v; float* gv; float* w;
};

template <bool ANALOG>
__device__ __forceinline__ void accumulate_entries(const GArgs& a, uint32_t first, uint32_t last, uint32_t step,
                                                   uint32_t col, long long (&acc)[4])
{
    const float* dcol = a.delta + col;
    uint32_t e = first;
    for (; e + (kGUnroll - 1) * step < last && e + (kGUnroll - 1) * step >= e; e += kGUnroll * step) {
        float4 x[kGUnroll]; float tv[kGUnroll];
#pragma unroll
        for (int u = 0; u < kGUnroll; u++) {
            const uint32_t ee = e + u * step;
            const uint32_t row = __ldg(a.tIndex + ee);
            tv[u] = ANALOG ? __ldg(a.tData + ee) : 1.0f;
            x[u] = ldg_nc_f4(reinterpret_cast<const float4*>(dcol + (size_t)row * a.n));
        }
#pragma unroll
        for (int u = 0; u < kGUnroll; u++) {
            if (ANALOG) { x[u].x *= tv[u]; x[u].y *= tv[u]; x[u].z *= tv[u]; x[u].w *= tv[u]; }
            acc[0] += fix30(x[u].x); acc[1] += fix30(x[u].y); acc[2] += fix30(x[u].z); acc[3] += fix30(x[u].w);
        }
    }
    for (; e < last; e += step) {
        const uint32_t row = __ldg(a.tIndex + e);
        float4 x = ldg_nc_f4(reinterpret_cast<const float4*>(dcol + (size_t)row * a.n));
        if (ANALOG) { const float tv = __ldg(a.tData + e); x.x *= tv; x.y *= tv; x.z *= tv; x.w *= tv; }
        acc[0] += fix30(x.x); acc[1] += fix30(x.y); acc[2] += fix30(x.z); acc[3] += fix30(x.w);
    }
}

template <bool ANALOG, int FUSED_MODE>      // FUSED_MODE = -1: write dW
__device__ __forceinline__ void finish_row(const GArgs& a, uint32_t c, uint32_t col, const long long (&acc)[4])
{
    float g[4];
#pragma unroll
    for (int v = 0; v < 4; v++) g[v] = a.alpha * (__ll2float_rn(acc[v]) * 9.31322574615478515625e-10f);   // acc * 2^-30, one rounding: same value as
                                                                                                         // (float)((double)acc * 2^-30), E/kernels.cu:2585
    const size_t off = (size_t)c * a.n + col;
    if (FUSED_MODE < 0) {
        float4 out = make_float4(g[0], g[1], g[2], g[3]);
        if (a.beta != 0.0f) {
            const float4 old = *reinterpret_cast<const float4*>(a.dW + off);
            out.x += a.beta * old.x; out.y += a.beta * old.y; out.z += a.beta * old.z; out.w += a.beta * old.w;
        }
        *reinterpret_cast<float4*>(a.dW + off) = out;
    } else {
        constexpr int M = FUSED_MODE < 0 ? 0 : FUSED_MODE;
        float4 w4 = *reinterpret_cast<const float4*>(a.w + off);
        float4 v4 = make_float4(0, 0, 0, 0), s4 = make_float4(0, 0, 0, 0);
        if (opt_uses_v(M))  v4 = *reinterpret_cast<const float4*>(a.v + off);
        if (opt_uses_gv(M)) s4 = *reinterpret_cast<const float4*>(a.gv + off);
        w4.x = opt_weight<M>(a.opt, g[0], w4.x, v4.x, s4.x);
        w4.y = opt_weight<M>(a.opt, g[1], w4.y, v4.y, s4.y);
        w4.z = opt_weight<M>(a.opt, g[2], w4.z, v4.z, s4.z);
        w4.w = opt_weight<M>(a.opt, g[3], w4.w, v4.w, s4.w);
        *reinterpret_cast<float4*>(a.w + off) = w4;
        if (opt_uses_v(M))  *reinterpret_cast<float4*>(a.v + off) = v4;
        if (opt_uses_gv(M)) *reinterpret_cast<float4*>(a.gv + off) = s4;
    }
}

// n % 4 == 0.  lpr = lanes per delta row (power of two >= n/4, <= 256); G = 256/lpr groups.
template <bool ANALOG, int FUSED_MODE>
__global__ void __launch_bounds__(kGThreads, 3)
sparse_wgrad_kernel(const GArgs a, uint32_t lpr)
{
    extern __shared__ long long sRed[];            // [G][lpr*4] int64 partials for heavy columns
    __shared__ uint32_t sHeavy[kGThreads];         // heavy columns of the current tile (<= G)
    __shared__ uint32_t sNHeavy;
    const uint32_t tid = threadIdx.x;
    const uint32_t G = kGThreads / lpr, g = tid / lpr, lane = tid % lpr;
    const uint32_t n4 = a.n >> 2;
    const uint32_t colBlocks = (n4 + lpr - 1) / lpr;
    const uint32_t tiles = (a.m + G - 1) / G;

    for (uint32_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        if (tid == 0) sNHeavy = 0;
        __syncthreads();
        const uint32_t c = tile * G + g;
        uint32_t s = 0, e = 0;
        if (c < a.m) { s = __ldg(a.tStart + c); e = __ldg(a.tEnd + c); }
        const uint32_t cnt = e - s;
        const bool heavy = (G > 1) && cnt > kHeavy;
        if (heavy && lane == 0) sHeavy[atomicAdd(&sNHeavy, 1u)] = c;
        if (c < a.m && !heavy) {
            for (uint32_t cb = 0; cb < colBlocks; cb++) {
                const uint32_t col = (cb * lpr + lane) * 4;
                if (col < a.n) {
                    long long acc[4] = {0, 0, 0, 0};
                    accumulate_entries<ANALOG>(a, s, e, 1, col, acc);
                    finish_row<ANALOG, FUSED_MODE>(a, c, col, acc);
                }
            }
        }
        __syncthreads();
        const uint32_t nh = sNHeavy;
        for (uint32_t h = 0; h < nh; h++) {
            // heavy columns of the tile in the order their warps found them (the sums are integer, so the
            // result does not depend on it)
            const uint32_t hc = sHeavy[h];
            const uint32_t hs = __ldg(a.tStart + hc), he = __ldg(a.tEnd + hc);
            for (uint32_t cb = 0; cb < colBlocks; cb++) {
                const uint32_t col = (cb * lpr + lane) * 4;
                long long acc[4] = {0, 0, 0, 0};
                if (col < a.n) accumulate_entries<ANALOG>(a, hs + g, he, G, col, acc);
#pragma unroll
                for (int v = 0; v < 4; v++) sRed[((size_t)g * lpr + lane) * 4 + v] = acc[v];
                __syncthreads();
                if (g == 0 && col < a.n) {
                    long long tot[4] = {0, 0, 0, 0};
                    for (uint32_t gg = 0; gg < G; gg++)
#pragma unroll
                        for (int v = 0; v < 4; v++) tot[v] += sRed[((size_t)gg * lpr + lane) * 4 + v];
                    finish_row<ANALOG, FUSED_MODE>(a, hc, col, tot);
                }
                __syncthreads();
            }
        }
    }
}

// scalar fallback for strides that are not a multiple of 4 (or unaligned buffers): one thread
// per (column, output) pair, same fixed-point arithmetic.
template <bool ANALOG>
__global__ void __launch_bounds__(256)
sparse_wgrad_scalar_kernel(const GArgs a)
{
    const uint64_t total = (uint64_t)a.m * a.n;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t c = (uint32_t)(i / a.n), o = (uint32_t)(i % a.n);
        long long acc = 0;
        for (uint32_t e = __ldg(a.tStart + c); e < __ldg(a.tEnd + c); e++) {
            float x = __ldg(a.delta + (size_t)__ldg(a.tIndex + e) * a.n + o);
            if (ANALOG) x *= __ldg(a.tData + e);
            acc += fix30(x);
        }
        const float g = a.alpha * (float)((double)acc * kOneOverErrorScale);
        a.dW[i] = ((a.beta == 0.0f) ? 0.0f : a.beta * a.dW[i]) + g;
    }
}

static uint32_t lanes_per_row(uint32_t n)
{
    uint32_t n4 = n >> 2, lpr = 1;
    while (lpr < n4 && lpr < (uint32_t)kGThreads) lpr <<= 1;
    return lpr;
}

template <bool ANALOG, int FUSED_MODE>
static int launch_wgrad(dsb200_ctx* ctx, const GArgs& a)
{
    const uint32_t lpr = lanes_per_row(a.n);
    const uint32_t G = kGThreads / lpr;
    const uint32_t tiles = (a.m + G - 1) / G;
    int grid = ctx->numSMs * 3;
    if ((uint32_t)grid > tiles) grid = (int)tiles;
    if (grid < 1) grid = 1;
    const size_t smem = (size_t)kGThreads * 4 * sizeof(long long);
    sparse_wgrad_kernel<ANALOG, FUSED_MODE><<<grid, kGThreads, smem, ctx->stream>>>(a, lpr);
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    return 0;
}


// ---------------------------------------------------------------- two-kernel scheme (n % 128 == 0)
// The column population of a recommender batch is extremely skewed: ~5 entries for the typical column, up to `batch`
// for the Zipf head.  Kernel A gives every light column (<= kHeavy entries) to ONE WARP -- indices read coalesced into
// registers and broadcast with shuffles, each lane gathers its float4 of the delta rows, no shared memory and no block
// barrier, ~40 warps per SM hide the three dependent latencies (start/end -> indices -> delta rows) -- and appends the
// heavy columns to a list.  Kernel B takes one heavy column per CTA: its 8 warps split the entries and add their int64
// fixed-point partial sums through shared memory.  Both finish a row the same way (gradient or fused optimizer).
template <bool ANALOG, uint32_t step>
__device__ __forceinline__ void gather_warp(const GArgs& a, uint32_t s, uint32_t e, uint32_t first, uint32_t col, uint32_t lane, long long (&acc)[4])
{
    // entries s + first, s + first + step, ... < e; processed in blocks of 32 (one index per lane)
    const float* dcol = a.delta + col;
    for (uint32_t base = s + first; base < e; base += 32 * step) {
        const uint32_t mine = base + lane * step;
        uint32_t myRow = 0;
        float myVal = 1.0f;
        if (mine < e) {
            myRow = __ldg(a.tIndex + mine);
            if (ANALOG) myVal = __ldg(a.tData + mine);
        }
        const uint32_t n = min(32u, (e - base + step - 1) / step);
        uint32_t j = 0;
        for (; j + kGUnroll <= n; j += kGUnroll) {
            float4 x[kGUnroll]; float tv[kGUnroll];
#pragma unroll
            for (int u = 0; u < kGUnroll; u++) {
                const uint32_t row = __shfl_sync(0xffffffffu, myRow, j + u);
                if (ANALOG) tv[u] = __shfl_sync(0xffffffffu, myVal, j + u);
                x[u] = ldg_nc_f4(reinterpret_cast<const float4*>(dcol + (size_t)row * a.n));
            }
#pragma unroll
            for (int u = 0; u < kGUnroll; u++) {
                if (ANALOG) { x[u].x *= tv[u]; x[u].y *= tv[u]; x[u].z *= tv[u]; x[u].w *= tv[u]; }
                acc[0] += fix30(x[u].x); acc[1] += fix30(x[u].y); acc[2] += fix30(x[u].z); acc[3] += fix30(x[u].w);
            }
        }
        for (; j < n; j++) {
            const uint32_t row = __shfl_sync(0xffffffffu, myRow, j);
            float4 x = ldg_nc_f4(reinterpret_cast<const float4*>(dcol + (size_t)row * a.n));
            if (ANALOG) { const float tv = __shfl_sync(0xffffffffu, myVal, j); x.x *= tv; x.y *= tv; x.z *= tv; x.w *= tv; }
            acc[0] += fix30(x.x); acc[1] += fix30(x.y); acc[2] += fix30(x.z); acc[3] += fix30(x.w);
        }
    }
}

// Light columns hold <= 32 entries: one index (and value) per lane.  The warp software-pipelines its columns so that only
// ONE memory latency per column stays on the critical path: the column range is fetched two columns ahead, the indices
// one column ahead, and the weight / state row of the current column is requested before its delta rows are gathered.
template <bool ANALOG, int FUSED_MODE>
__global__ void __launch_bounds__(kGThreads, 3)
sparse_wgrad_light_kernel(const GArgs a, uint32_t* __restrict__ heavyList, uint32_t* __restrict__ heavyCount)
{
    constexpr int M = FUSED_MODE < 0 ? 0 : FUSED_MODE;
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * kGThreads + threadIdx.x) >> 5, nw = (gridDim.x * kGThreads) >> 5;
    // pipeline registers: range of column c + 2 nw, range + indices of column c + nw
    uint32_t s2 = 0, e2 = 0, s1 = 0, e1 = 0, row1 = 0;
    float val1 = 1.0f;
    auto load_range = [&](uint32_t c, uint32_t& s, uint32_t& e) { s = 0; e = 0; if (c < a.m) { s = __ldg(a.tStart + c); e = __ldg(a.tEnd + c); } };
    auto load_entries = [&](uint32_t s, uint32_t e, uint32_t& row, float& val) {
        row = 0; val = 1.0f;
        if (e - s <= kHeavy2 && s + lane < e) { row = __ldg(a.tIndex + s + lane); if (ANALOG) val = __ldg(a.tData + s + lane); }
    };
    load_range(gw, s1, e1);
    load_range(gw + nw, s2, e2);
    load_entries(s1, e1, row1, val1);
    for (uint32_t c = gw; c < a.m; c += nw) {
        const uint32_t s = s1, e = e1, myRow = row1;
        const float myVal = val1;
        s1 = s2; e1 = e2;
        load_entries(s1, e1, row1, val1);                                     // indices of the next column
        load_range(c + 2 * nw, s2, e2);                                       // range of the one after
        if (e - s > kHeavy2) {
            if (lane == 0) heavyList[atomicAdd(heavyCount, 1u)] = c;
            continue;
        }
        const uint32_t cnt = e - s;
        for (uint32_t col = lane * 4; col < a.n; col += 128) {            // n % 128 == 0: every lane is active
            const size_t off = (size_t)c * a.n + col;
            // request the row this column updates before the gathers (its latency overlaps theirs)
            float4 w4 = make_float4(0, 0, 0, 0), v4 = make_float4(0, 0, 0, 0), g4 = make_float4(0, 0, 0, 0);
            if (FUSED_MODE >= 0) {
                w4 = *reinterpret_cast<const float4*>(a.w + off);
                if (opt_uses_v(M))  v4 = *reinterpret_cast<const float4*>(a.v + off);
                if (opt_uses_gv(M)) g4 = *reinterpret_cast<const float4*>(a.gv + off);
            } else if (a.beta != 0.0f) w4 = *reinterpret_cast<const float4*>(a.dW + off);
            long long acc[4] = {0, 0, 0, 0};
            const float* dcol = a.delta + col;
            uint32_t j = 0;
            for (; j + kGUnroll <= cnt; j += kGUnroll) {
                float4 x[kGUnroll]; float tv[kGUnroll];
#pragma unroll
                for (int u = 0; u < kGUnroll; u++) {
                    const uint32_t row = __shfl_sync(0xffffffffu, myRow, j + u);
                    if (ANALOG) tv[u] = __shfl_sync(0xffffffffu, myVal, j + u);
                    x[u] = ldg_nc_f4(reinterpret_cast<const float4*>(dcol + (size_t)row * a.n));
                }
#pragma unroll
                for (int u = 0; u < kGUnroll; u++) {
                    if (ANALOG) { x[u].x *= tv[u]; x[u].y *= tv[u]; x[u].z *= tv[u]; x[u].w *= tv[u]; }
                    acc[0] += fix30(x[u].x); acc[1] += fix30(x[u].y); acc[2] += fix30(x[u].z); acc[3] += fix30(x[u].w);
                }
            }
            if (j < cnt) {                                                     // up to 7 more: issue them together
                float4 x[kGUnroll]; float tv[kGUnroll];
#pragma unroll
                for (int u = 0; u < kGUnroll - 1; u++) {
                    const uint32_t row = __shfl_sync(0xffffffffu, myRow, min(j + u, 31u));
                    if (ANALOG) tv[u] = __shfl_sync(0xffffffffu, myVal, min(j + u, 31u));
                    if (j + u < cnt) x[u] = ldg_nc_f4(reinterpret_cast<const float4*>(dcol + (size_t)row * a.n));
                }
#pragma unroll
                for (int u = 0; u < kGUnroll - 1; u++) {
                    if (j + u < cnt) {
                        if (ANALOG) { x[u].x *= tv[u]; x[u].y *= tv[u]; x[u].z *= tv[u]; x[u].w *= tv[u]; }
                        acc[0] += fix30(x[u].x); acc[1] += fix30(x[u].y); acc[2] += fix30(x[u].z); acc[3] += fix30(x[u].w);
                    }
                }
            }
            float g[4];
#pragma unroll
            for (int v = 0; v < 4; v++) g[v] = a.alpha * (__ll2float_rn(acc[v]) * 9.31322574615478515625e-10f);
            if (FUSED_MODE < 0) {
                float4 out = make_float4(g[0], g[1], g[2], g[3]);
                if (a.beta != 0.0f) { out.x += a.beta * w4.x; out.y += a.beta * w4.y; out.z += a.beta * w4.z; out.w += a.beta * w4.w; }
                *reinterpret_cast<float4*>(a.dW + off) = out;
            } else {
                w4.x = opt_weight<M>(a.opt, g[0], w4.x, v4.x, g4.x);
                w4.y = opt_weight<M>(a.opt, g[1], w4.y, v4.y, g4.y);
                w4.z = opt_weight<M>(a.opt, g[2], w4.z, v4.z, g4.z);
                w4.w = opt_weight<M>(a.opt, g[3], w4.w, v4.w, g4.w);
                *reinterpret_cast<float4*>(a.w + off) = w4;
                if (opt_uses_v(M))  *reinterpret_cast<float4*>(a.v + off) = v4;
                if (opt_uses_gv(M)) *reinterpret_cast<float4*>(a.gv + off) = g4;
            }
        }
    }
}

template <bool ANALOG, int FUSED_MODE>
__global__ void __launch_bounds__(kGThreads, 2)
sparse_wgrad_heavy_kernel(const GArgs a, const uint32_t* __restrict__ heavyList, uint32_t* __restrict__ heavyCount)
{
    constexpr int M = FUSED_MODE < 0 ? 0 : FUSED_MODE;
    __shared__ long long sRed[kGThreads / 32][128];                          // one 128-column block of int64 partials per warp
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr uint32_t W = kGThreads / 32;
    const uint32_t nh = *heavyCount;
    // column id and range of the next heavy column are fetched while the current one is reduced
    uint32_t cN = 0, sN = 0, eN = 0;
    if (blockIdx.x < nh) { cN = heavyList[blockIdx.x]; sN = __ldg(a.tStart + cN); eN = __ldg(a.tEnd + cN); }
    for (uint32_t h = blockIdx.x; h < nh; h += gridDim.x) {
        const uint32_t c = cN, s = sN, e = eN;
        if (h + gridDim.x < nh) { cN = heavyList[h + gridDim.x]; sN = __ldg(a.tStart + cN); eN = __ldg(a.tEnd + cN); }
        for (uint32_t col = lane * 4; col < a.n; col += 128) {
            const size_t off = (size_t)c * a.n + col;
            float4 w4 = make_float4(0, 0, 0, 0), v4 = make_float4(0, 0, 0, 0), g4 = make_float4(0, 0, 0, 0);
            if (warp == 0) {                                                  // the finishing warp requests its row up front
                if (FUSED_MODE >= 0) {
                    w4 = *reinterpret_cast<const float4*>(a.w + off);
                    if (opt_uses_v(M))  v4 = *reinterpret_cast<const float4*>(a.v + off);
                    if (opt_uses_gv(M)) g4 = *reinterpret_cast<const float4*>(a.gv + off);
                } else if (a.beta != 0.0f) w4 = *reinterpret_cast<const float4*>(a.dW + off);
            }
            long long acc[4] = {0, 0, 0, 0};
            gather_warp<ANALOG, W>(a, s, e, warp, col, lane, acc);            // warp w takes entries w, w + 8, ...
#pragma unroll
            for (int v = 0; v < 4; v++) sRed[warp][lane * 4 + v] = acc[v];
            __syncthreads();
            if (warp == 0) {
                long long tot[4] = {0, 0, 0, 0};
#pragma unroll
                for (uint32_t w = 0; w < W; w++)
#pragma unroll
                    for (int v = 0; v < 4; v++) tot[v] += sRed[w][lane * 4 + v];
                float g[4];
#pragma unroll
                for (int v = 0; v < 4; v++) g[v] = a.alpha * (__ll2float_rn(tot[v]) * 9.31322574615478515625e-10f);
                if (FUSED_MODE < 0) {
                    float4 out = make_float4(g[0], g[1], g[2], g[3]);
                    if (a.beta != 0.0f) { out.x += a.beta * w4.x; out.y += a.beta * w4.y; out.z += a.beta * w4.z; out.w += a.beta * w4.w; }
                    *reinterpret_cast<float4*>(a.dW + off) = out;
                } else {
                    w4.x = opt_weight<M>(a.opt, g[0], w4.x, v4.x, g4.x);
                    w4.y = opt_weight<M>(a.opt, g[1], w4.y, v4.y, g4.y);
                    w4.z = opt_weight<M>(a.opt, g[2], w4.z, v4.z, g4.z);
                    w4.w = opt_weight<M>(a.opt, g[3], w4.w, v4.w, g4.w);
                    *reinterpret_cast<float4*>(a.w + off) = w4;
                    if (opt_uses_v(M))  *reinterpret_cast<float4*>(a.v + off) = v4;
                    if (opt_uses_gv(M)) *reinterpret_cast<float4*>(a.gv + off) = g4;
                }
            }
            __syncthreads();
        }
    }
}

// the heavy kernel reads the count the light kernel produced; the count is cleared by a memset node ahead of the pair
template <bool ANALOG, int FUSED_MODE>
static int launch_wgrad2(dsb200_ctx* ctx, const GArgs& a)
{
    if (a.m + 1 > ctx->heavyCap) {
        DSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->dHeavy);
        ctx->dHeavy = nullptr; ctx->heavyCap = 0;
        DSB_CUDA_OK(cudaMalloc(&ctx->dHeavy, ((size_t)a.m + 1) * sizeof(uint32_t)));
        ctx->heavyCap = a.m + 1;
    }
    uint32_t* count = ctx->dHeavy;                                            // [0] = count, [1..] = list
    DSB_CUDA_OK(cudaMemsetAsync(count, 0, sizeof(uint32_t), ctx->stream));
    const uint32_t warps = kGThreads / 32;
    // One resident wave: the kernel is grid-strided over warps, so blocks beyond what fits on the GPU at once only form a second,
    // thinly occupied wave (ncu, round 1: 592 blocks launched, 3 per SM resident at 80 registers -> 25 % of the warp slots
    // active on average).  Option "wgrad_light_blocks" overrides the blocks per SM (0 = ask the occupancy calculator).
    static int blocksPerSM = 0;                                               // per instantiation
    if (!blocksPerSM) {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sparse_wgrad_light_kernel<ANALOG, FUSED_MODE>, kGThreads, 0) != cudaSuccess || occ < 1) {
            cudaGetLastError();
            occ = 4;                                                          // the former fixed choice
        }
        blocksPerSM = occ;
    }
    int grid = ctx->numSMs * (ctx->wgradLightBlocks > 0 ? ctx->wgradLightBlocks : blocksPerSM);
    if ((uint32_t)grid > (a.m + warps - 1) / warps) grid = (int)((a.m + warps - 1) / warps);
    if (grid < 1) grid = 1;
    sparse_wgrad_light_kernel<ANALOG, FUSED_MODE><<<grid, kGThreads, 0, ctx->stream>>>(a, count + 1, count);
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    sparse_wgrad_heavy_kernel<ANALOG, FUSED_MODE><<<ctx->numSMs * 2, kGThreads, 0, ctx->stream>>>(a, count + 1, count);
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------- unified scheme (n % 128 == 0), round 2
// Round 1's pair above ran the light columns (30 us on BASELINE config 2) and then the heavy ones (21 us) although the two sets touch
// disjoint rows, and the light kernel spilled its look-ahead registers.  Here both populations run in ONE grid:
//   wgrad_classify_kernel   one pass over the column ranges: every column with more than 32 entries is cut into items of 32 entries
//                           (column, first entry) and gets a slot in a zeroed int64 accumulator [slot][n]
//   sparse_wgrad_unified    every warp first takes heavy items (list index = warp id, + warps in the grid, ...): gathers its <= 32
//                           delta rows, adds its int64 partial sums into the column's accumulator with 64-bit atomics (exact: the
//                           result does not depend on their order) and counts itself in; the LAST item of a column reads the totals
//                           back, finishes the row (gradient or fused optimizer rule) and re-zeroes the slot.  Then the warp walks its
//                           share of the light columns (<= 32 entries: one index per lane); the row it updates is requested first
//                           (HBM, the longest latency), and the gathers run in exact-count batches -- round-2 ncu of the first version:
//                           380 warp instructions per column at ~5 entries per column, most of them predicated-off slots of a
//                           padded batch of 7, issue slots 40 % busy: the kernel was instruction bound.
// A variant with one QUARTER-warp per column (four columns per warp instruction) was measured at 64 us against 36 us and dropped.
struct UArgs {
    uint32_t* items;            // [2 * i] = column, [2 * i + 1] = first entry
    uint32_t* itemCount;        // [0] items, [1] heavy columns
    uint32_t* slotOf;           // [m] accumulator slot of a heavy column
    uint32_t* arrived;          // [slots] items of the column finished so far (zero between calls)
    long long* acc;             // [slots][n] (zero between calls)
    uint32_t maxItems, maxSlots;
    uint32_t* status;           // sticky device status word of the context
};

__global__ void __launch_bounds__(256)
wgrad_classify_kernel(uint32_t m, const uint32_t* __restrict__ tStart, const uint32_t* __restrict__ tEnd, const UArgs u)
{
    pdl_launch_dependents();
    pdl_wait();
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < m; c += gridDim.x * blockDim.x) {
        const uint32_t s = __ldg(tStart + c), e = __ldg(tEnd + c);
        if (e - s <= kHeavy2) continue;
        const uint32_t n = (e - s + 31) / 32;
        const uint32_t slot = atomicAdd(u.itemCount + 1, 1u);
        const uint32_t first = atomicAdd(u.itemCount, n);
        if (slot >= u.maxSlots || first + n > u.maxItems) { atomicOr(u.status, DSB200_STATUS_G_CAPACITY); continue; }   // reported by dsb200_ctx_sync
        u.slotOf[c] = slot;
        for (uint32_t i = 0; i < n; i++) { u.items[2 * (first + i)] = c; u.items[2 * (first + i) + 1] = s + 32 * i; }
    }
}

// sum of <= 32 entries whose delta-row indices sit one per lane: full batches of 8 gathers, then the remainder with EXACT counts
// (cnt is warp-uniform: the switch is a jump, not predication)
template <bool ANALOG>
__device__ __forceinline__ void gather32(const GArgs& a, const float* dcol, uint32_t cnt, uint32_t myRow, float myVal, long long (&acc)[4])
{
    auto ld = [&](uint32_t j, float4& x, float& tv) {
        const uint32_t row = __shfl_sync(0xffffffffu, myRow, j);
        if (ANALOG) tv = __shfl_sync(0xffffffffu, myVal, j);
        x = ldg_nc_f4(reinterpret_cast<const float4*>(dcol + (size_t)row * a.n));
    };
    auto add = [&](float4 x, float tv) {
        if (ANALOG) { x.x *= tv; x.y *= tv; x.z *= tv; x.w *= tv; }
        acc[0] += fix30(x.x); acc[1] += fix30(x.y); acc[2] += fix30(x.z); acc[3] += fix30(x.w);
    };
    uint32_t j = 0;
    for (; j + kGUnroll <= cnt; j += kGUnroll) {
        float4 x[kGUnroll]; float tv[kGUnroll];
#pragma unroll
        for (int q = 0; q < kGUnroll; q++) { tv[q] = 1.0f; ld(j + q, x[q], tv[q]); }
#pragma unroll
        for (int q = 0; q < kGUnroll; q++) add(x[q], tv[q]);
    }
    float4 x0, x1, x2, x3; float t0 = 1.f, t1 = 1.f, t2 = 1.f, t3 = 1.f;
    uint32_t rem = cnt - j;
    if (rem >= 4) { ld(j, x0, t0); ld(j + 1, x1, t1); ld(j + 2, x2, t2); ld(j + 3, x3, t3); add(x0, t0); add(x1, t1); add(x2, t2); add(x3, t3); j += 4; rem -= 4; }
    switch (rem) {
    case 3: ld(j, x0, t0); ld(j + 1, x1, t1); ld(j + 2, x2, t2); add(x0, t0); add(x1, t1); add(x2, t2); break;
    case 2: ld(j, x0, t0); ld(j + 1, x1, t1); add(x0, t0); add(x1, t1); break;
    case 1: ld(j, x0, t0); add(x0, t0); break;
    default: break;
    }
}

struct RowRegs { float4 w, v, g; };
template <int FUSED_MODE>
__device__ __forceinline__ void load_row(const GArgs& a, size_t off, RowRegs& r)
{
    constexpr int M = FUSED_MODE < 0 ? 0 : FUSED_MODE;
    r.w = make_float4(0, 0, 0, 0); r.v = r.w; r.g = r.w;
    if (FUSED_MODE >= 0) {
        r.w = *reinterpret_cast<const float4*>(a.w + off);
        if (opt_uses_v(M))  r.v = *reinterpret_cast<const float4*>(a.v + off);
        if (opt_uses_gv(M)) r.g = *reinterpret_cast<const float4*>(a.gv + off);
    } else if (a.beta != 0.0f) r.w = *reinterpret_cast<const float4*>(a.dW + off);
}
template <int FUSED_MODE>
__device__ __forceinline__ void finish_row_regs(const GArgs& a, size_t off, RowRegs& r, const long long (&acc)[4])
{
    constexpr int M = FUSED_MODE < 0 ? 0 : FUSED_MODE;
    float g[4];
#pragma unroll
    for (int v = 0; v < 4; v++) g[v] = a.alpha * (__ll2float_rn(acc[v]) * 9.31322574615478515625e-10f);
    if (FUSED_MODE < 0) {
        float4 out = make_float4(g[0], g[1], g[2], g[3]);
        if (a.beta != 0.0f) { out.x += a.beta * r.w.x; out.y += a.beta * r.w.y; out.z += a.beta * r.w.z; out.w += a.beta * r.w.w; }
        *reinterpret_cast<float4*>(a.dW + off) = out;
    } else {
        r.w.x = opt_weight<M>(a.opt, g[0], r.w.x, r.v.x, r.g.x);
        r.w.y = opt_weight<M>(a.opt, g[1], r.w.y, r.v.y, r.g.y);
        r.w.z = opt_weight<M>(a.opt, g[2], r.w.z, r.v.z, r.g.z);
        r.w.w = opt_weight<M>(a.opt, g[3], r.w.w, r.v.w, r.g.w);
        *reinterpret_cast<float4*>(a.w + off) = r.w;
        if (opt_uses_v(M))  *reinterpret_cast<float4*>(a.v + off) = r.v;
        if (opt_uses_gv(M)) *reinterpret_cast<float4*>(a.gv + off) = r.g;
    }
}

template <bool ANALOG, int FUSED_MODE>
__global__ void __launch_bounds__(kGThreads, 2)
sparse_wgrad_unified_kernel(const GArgs a, const UArgs u)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * kGThreads + threadIdx.x) >> 5, nw = (gridDim.x * kGThreads) >> 5;
    pdl_launch_dependents();
    pdl_wait();
    // ---- heavy columns, 32 entries per item
    const uint32_t nItems = min(*u.itemCount, u.maxItems);
    for (uint32_t it = gw; it < nItems; it += nw) {
        const uint32_t c = __ldg(u.items + 2 * it), e0 = __ldg(u.items + 2 * it + 1);
        const uint32_t s = __ldg(a.tStart + c), e = __ldg(a.tEnd + c);
        const uint32_t cnt = min(32u, e - e0), slot = __ldg(u.slotOf + c), nItemsCol = (e - s + 31) / 32;
        uint32_t myRow = 0; float myVal = 1.0f;
        if (lane < cnt) { myRow = __ldg(a.tIndex + e0 + lane); if (ANALOG) myVal = __ldg(a.tData + e0 + lane); }
        long long* accRow = u.acc + (size_t)slot * a.n;
        for (uint32_t col = lane * 4; col < a.n; col += 128) {
            long long acc[4] = {0, 0, 0, 0};
            gather32<ANALOG>(a, a.delta + col, cnt, myRow, myVal, acc);
#pragma unroll
            for (int v = 0; v < 4; v++) atomicAdd(reinterpret_cast<unsigned long long*>(accRow + col + v), (unsigned long long)acc[v]);
        }
        __threadfence();
        __syncwarp();
        uint32_t last = 0;
        if (lane == 0) last = (atomicAdd(u.arrived + slot, 1u) == nItemsCol - 1) ? 1u : 0u;
        last = __shfl_sync(0xffffffffu, last, 0);
        if (!last) continue;
        __threadfence();
        if (lane == 0) u.arrived[slot] = 0;                                   // zero again for the next call
        for (uint32_t col = lane * 4; col < a.n; col += 128) {
            long long tot[4];
#pragma unroll
            for (int v = 0; v < 4; v++) {
                tot[v] = (long long)__ldcg(reinterpret_cast<const unsigned long long*>(accRow + col + v));
                accRow[col + v] = 0;
            }
            finish_row<ANALOG, FUSED_MODE>(a, c, col, tot);
        }
    }
    // ---- light columns: one warp each
    for (uint32_t c = gw; c < a.m; c += nw) {
        const size_t rowOff = (size_t)c * a.n + lane * 4;
        RowRegs r;
        load_row<FUSED_MODE>(a, rowOff, r);                                    // HBM: the longest latency of the column, asked for first
        const uint32_t s = __ldg(a.tStart + c), e = __ldg(a.tEnd + c), cnt = e - s;
        if (cnt > kHeavy2) continue;
        uint32_t myRow = 0; float myVal = 1.0f;
        if (lane < cnt) { myRow = __ldg(a.tIndex + s + lane); if (ANALOG) myVal = __ldg(a.tData + s + lane); }
        for (uint32_t col = lane * 4; col < a.n; col += 128) {
            RowRegs rn;
            if (col + 128 < a.n) load_row<FUSED_MODE>(a, rowOff + (col - lane * 4) + 128, rn);   // next 128-column block of the row
            long long acc[4] = {0, 0, 0, 0};
            gather32<ANALOG>(a, a.delta + col, cnt, myRow, myVal, acc);
            finish_row_regs<FUSED_MODE>(a, (size_t)c * a.n + col, r, acc);
            r = rn;
        }
    }
}

template <bool ANALOG, int FUSED_MODE>
static int launch_wgrad3(dsb200_ctx* ctx, const GArgs& a, uint32_t maxEntries)
{
    // capacities: a heavy column holds > 32 entries -> at most maxEntries / 33 of them, and at most maxEntries / 32 + that many items
    const uint32_t maxSlots = maxEntries / 33 + 1, maxItems = maxEntries / 32 + maxSlots + 1;
    const size_t need = 2 * sizeof(uint32_t) * 2 + (size_t)a.m * 4 + (size_t)maxSlots * 4 + (size_t)maxItems * 8 + (size_t)maxSlots * a.n * 8 + 64;
    if (need > ctx->heavy3Bytes) {
        DSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->dHeavy3);
        ctx->dHeavy3 = nullptr; ctx->heavy3Bytes = 0;
        DSB_CUDA_OK(cudaMalloc(&ctx->dHeavy3, need));
        DSB_CUDA_OK(cudaMemsetAsync(ctx->dHeavy3, 0, need, ctx->stream));          // accumulators and arrival counters start (and stay) zero
        ctx->heavy3Bytes = need;
        ctx->heavySlots = maxSlots; ctx->heavyN = a.n; ctx->heavyM = a.m;
    } else if (ctx->heavySlots != maxSlots || ctx->heavyN != a.n || ctx->heavyM != a.m) {
        DSB_CUDA_OK(cudaMemsetAsync(ctx->dHeavy3, 0, ctx->heavy3Bytes, ctx->stream));   // the layout below changes: nothing of the old one may remain
        ctx->heavySlots = maxSlots; ctx->heavyN = a.n; ctx->heavyM = a.m;
    }
    uint8_t* p = reinterpret_cast<uint8_t*>(ctx->dHeavy3);
    UArgs u{};
    u.status = ctx->dStatus;
    u.acc = reinterpret_cast<long long*>(p); p += (size_t)maxSlots * a.n * 8;
    u.items = reinterpret_cast<uint32_t*>(p); p += (size_t)maxItems * 8;
    u.itemCount = reinterpret_cast<uint32_t*>(p); p += 16;
    u.slotOf = reinterpret_cast<uint32_t*>(p); p += (size_t)a.m * 4;
    u.arrived = reinterpret_cast<uint32_t*>(p);
    u.maxItems = maxItems; u.maxSlots = maxSlots;
    DSB_CUDA_OK(cudaMemsetAsync(u.itemCount, 0, 8, ctx->stream));
    DSB_CUDA_OK(launch_pdl(wgrad_classify_kernel, dim3(std::min<uint32_t>((a.m + 255) / 256, (uint32_t)ctx->numSMs * 4)), dim3(256), 0, ctx->stream, a.m, a.tStart, a.tEnd, u));
    count_launch();
    static int blocksPerSM = 0;                                               // per instantiation
    if (!blocksPerSM) {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sparse_wgrad_unified_kernel<ANALOG, FUSED_MODE>, kGThreads, 0) != cudaSuccess || occ < 1) {
            cudaGetLastError();
            occ = 2;
        }
        blocksPerSM = occ;
    }
    int grid = ctx->numSMs * (ctx->wgradLightBlocks > 0 ? ctx->wgradLightBlocks : blocksPerSM);
    const uint32_t warps = kGThreads / 32;
    if ((uint32_t)grid > (a.m + warps - 1) / warps) grid = (int)((a.m + warps - 1) / warps);
    if (grid < 1) grid = 1;
    DSB_CUDA_OK(launch_pdl(sparse_wgrad_unified_kernel<ANALOG, FUSED_MODE>, dim3(grid), dim3(kGThreads), 0, ctx->stream, a, u));
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    return 0;
}

static int check_wgrad_args(dsb200_ctx* ctx, const uint32_t* tStart, const uint32_t* tEnd, const uint32_t* tIndex, const float* delta)
{
    if (!ctx || !tStart || !tEnd || !tIndex || !delta) return fail(ctx, DSB200_EINVAL, "sparse_wgrad: null argument");
    return 0;
}

}  // namespace dsb

extern "C" {

int dsb200_sparse_wgrad(dsb200_ctx* ctx, float alpha, float beta, uint32_t m, uint32_t n,
                        const uint32_t* tStart, const uint32_t* tEnd, const uint32_t* tIndex, const float* tData,
                        const float* delta, float* dW)
{
    DSB_PROFILE(ctx, "sparse_wgrad");
    using namespace dsb;
    int rc = check_wgrad_args(ctx, tStart, tEnd, tIndex, delta);
    if (rc) return rc;
    if (!dW) return fail(ctx, DSB200_EINVAL, "sparse_wgrad: null gradient buffer");
    if (!m || !n) return 0;
    GArgs a{};
    a.alpha = alpha * ctx->params.denoising_q;              // E/kernels.cu:2547
    a.beta = beta; a.m = m; a.n = n;
    a.tStart = tStart; a.tEnd = tEnd; a.tIndex = tIndex; a.tData = tData; a.delta = delta; a.dW = dW;
    const bool vec = (n % 4 == 0) && ((((uintptr_t)delta | (uintptr_t)dW) % 16) == 0);
    if (vec && (n % 128 == 0) && !ctx->wgradTileKernel && !ctx->wgradTwoKernel) return tData ? launch_wgrad3<true, -1>(ctx, a, ctx->wgradMaxEntries) : launch_wgrad3<false, -1>(ctx, a, ctx->wgradMaxEntries);
    if (vec && (n % 128 == 0) && !ctx->wgradTileKernel) return tData ? launch_wgrad2<true, -1>(ctx, a) : launch_wgrad2<false, -1>(ctx, a);
    if (vec) return tData ? launch_wgrad<true, -1>(ctx, a) : launch_wgrad<false, -1>(ctx, a);
    uint64_t blocks = ((uint64_t)m * n + 255) / 256;
    if (blocks > (uint64_t)ctx->numSMs * 8) blocks = (uint64_t)ctx->numSMs * 8;
    if (tData) sparse_wgrad_scalar_kernel<true><<<(unsigned)blocks, 256, 0, ctx->stream>>>(a);
    else       sparse_wgrad_scalar_kernel<false><<<(unsigned)blocks, 256, 0, ctx->stream>>>(a);
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    return 0;
}

int dsb200_sparse_wgrad_update(dsb200_ctx* ctx, int mode, float galpha, uint32_t m, uint32_t n,
                               const uint32_t* tStart, const uint32_t* tEnd, const uint32_t* tIndex, const float* tData,
                               const float* delta, float alpha, float lambda, float lambda1, float mu, float mu1, float t,
                               float* v, float* gv, float* w)
{
    DSB_PROFILE(ctx, "sparse_wgrad_update");
    using namespace dsb;
    int rc = check_wgrad_args(ctx, tStart, tEnd, tIndex, delta);
    if (rc) return rc;
    if (!w) return fail(ctx, DSB200_EINVAL, "sparse_wgrad_update: null weight buffer");
    if (mode < 0 || mode > DSB200_ADAM) return fail(ctx, DSB200_EINVAL, "sparse_wgrad_update: bad mode");
    if (opt_uses_v(mode) && !v) return fail(ctx, DSB200_EINVAL, "sparse_wgrad_update: velocity buffer missing");
    if (opt_uses_gv(mode) && !gv) return fail(ctx, DSB200_EINVAL, "sparse_wgrad_update: gradient-velocity buffer missing");
    if ((n % 4) || ((((uintptr_t)delta | (uintptr_t)w | (uintptr_t)v | (uintptr_t)gv) % 16) != 0))
        return fail(ctx, DSB200_EUNSUPPORTED, "sparse_wgrad_update: needs stride % 4 == 0 and 16-byte aligned buffers");
    if (!m || !n) return 0;
    GArgs a{};
    a.alpha = galpha * ctx->params.denoising_q;
    a.beta = 0.0f; a.m = m; a.n = n;
    a.tStart = tStart; a.tEnd = tEnd; a.tIndex = tIndex; a.tData = tData; a.delta = delta; a.dW = nullptr;
    a.opt = make_opt(mode, alpha, lambda, lambda1, mu, mu1, t);
    a.v = v; a.gv = gv; a.w = w;
    const bool two = (n % 128 == 0) && !ctx->wgradTileKernel;
    const bool uni = two && !ctx->wgradTwoKernel;
#define DSB_FUSED(M) (uni ? (tData ? launch_wgrad3<true, M>(ctx, a, ctx->wgradMaxEntries) : launch_wgrad3<false, M>(ctx, a, ctx->wgradMaxEntries)) \
                    : two ? (tData ? launch_wgrad2<true, M>(ctx, a) : launch_wgrad2<false, M>(ctx, a)) \
                          : (tData ? launch_wgrad<true, M>(ctx, a) : launch_wgrad<false, M>(ctx, a)))
    switch (mode) {
    case DSB200_SGD:      return DSB_FUSED(DSB200_SGD);
    case DSB200_MOMENTUM: return DSB_FUSED(DSB200_MOMENTUM);
    case DSB200_ADAGRAD:  return DSB_FUSED(DSB200_ADAGRAD);
    case DSB200_NESTEROV: return DSB_FUSED(DSB200_NESTEROV);
    case DSB200_RMSPROP:  return DSB_FUSED(DSB200_RMSPROP);
    case DSB200_ADADELTA: return DSB_FUSED(DSB200_ADADELTA);
    default:              return DSB_FUSED(DSB200_ADAM);
    }
#undef DSB_FUSED
}

}  // extern "C"
