// sparse_z.cu -- sparse-input layer forward (hot-path rows a1-a3, a14, fused a9).
//
// Replaces kCalculate[Indexed]Sparse[Analog][Denoised]Z (E/kernels.cu:662-1977):
//     Z[b,:] = beta*Z[b,:] + w_b * sum_j v_j * W[idx_j,:]
//
// B200 design (not the reference's one-block-per-row / 32-columns-per-warp scheme):
//  * persistent CTAs (a multiple of the SM count); work item = (row, chunk of <= C nnz) so a
//    9,254-nnz row is spread over many SMs instead of serialising one block;
//  * every CTA derives the same item list from a block-wide scan of the row lengths kept in
//    shared memory (no extra launch, no device-wide sync);
//  * a dedicated producer warp stages each item's index / value / random runs into shared
//    memory with the TMA engine (cp.async.bulk + mbarrier, SASS UBLKCP), running ahead of the
//    8 consumer warps through a 3-stage full/empty mbarrier ring;
//  * consumers gather weight rows with 128-bit ld.global.nc loads, 8 independent rows in
//    flight per thread, one float4 column per lane; partial sums of the nnz-groups are
//    combined through shared memory;
//  * rows that span several chunks are combined deterministically: chunks write partial sums
//    to a context workspace and the last-arriving CTA adds them in chunk order (no float
//    atomics, bit-reproducible run to run).
#include "common.cuh"
#include "launch.h"

namespace dsb {

constexpr int kZConsumers   = 256;                 // 8 consumer warps
constexpr int kZThreads     = kZConsumers + 32;    // + 1 producer warp
constexpr int kZStages      = 3;
constexpr int kZMaxChunk    = 512;                 // nnz staged per item
constexpr int kZStageElems  = kZMaxChunk + 8;      // + alignment slack
constexpr int kZMaxRows     = 4096;                // rows planned per launch
constexpr int kZUnroll      = 8;

struct ZArgs {
    dsb200_params P;
    dsb200_sparse S;
    uint32_t position, batch, rowBase, stride;
    const float* W;
    const float* bias;      // fused variant: Zold = bias, beta = 1
    float*       Z;
    float        beta;
    int          activation; // -1: none
    int          chunk;      // C
    int          fused;      // empty rows still produce act(bias)
    int          useTma;
    uint32_t*    rowCounters;
    float*       partials;
    unsigned long long partialsCap;   // floats
    volatile uint32_t* status;
};

struct ZMeta {
    uint32_t row, k, nChunks, count, smemOff;
    float    w;
};

struct __align__(16) ZSmem {
    uint32_t idx[kZStages][kZStageElems];
    float    val[kZStages][kZStageElems];
    float    rnd[kZStages][kZStageElems];
    uint64_t full[kZStages];
    uint64_t empty[kZStages];
    ZMeta    meta[kZStages];
    uint32_t prefix[kZMaxRows + 1];
    uint32_t scan[16];
    uint32_t chunkUsed;
    uint32_t isLast;
};

__device__ __forceinline__ void consumer_sync()
{
    asm volatile("bar.sync 1, %0;" :: "n"(kZConsumers) : "memory");
}

__device__ __forceinline__ float apply_act(int act, float z)
{
    switch (act) {
    case DSB200_ACT_SIGMOID: return 1.0f / (1.0f + expf(-z));
    case DSB200_ACT_TANH:    return tanhf(z);
    case DSB200_ACT_RELU:    return fmaxf(0.0f, z);
    default:                 return z;
    }
}

// VEC = 4: stride % 4 == 0 and 16-byte aligned W/Z (float4 lanes); VEC = 1: scalar fallback.
template <int VEC, bool ANALOG, bool DENOISED>
__global__ void __launch_bounds__(kZThreads, 2)
sparse_z_kernel(const ZArgs a)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    ZSmem& sm = *reinterpret_cast<ZSmem*>(smemRaw);
    float* sRed = reinterpret_cast<float*>(smemRaw + sizeof(ZSmem));   // [G][lanesPerRow*VEC]

    const int tid = threadIdx.x;
    const uint32_t batch = a.batch;

    // ---------------------------------------------------------------- plan
    if (tid == 0) {
        for (int s = 0; s < kZStages; s++) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], kZConsumers / 32); }
        mbar_fence_init();
    }
    // row lengths -> chunk counts (coalesced), retry with a larger chunk if the split-row
    // workspace would overflow (every CTA takes the same decision)
    uint32_t C = (uint32_t)a.chunk;
    uint32_t T = 0;
    for (;;) {
        for (uint32_t r = tid; r < batch; r += kZThreads) {
            uint32_t ex = example_of(a.P, a.S.index, a.position, a.rowBase + r);
            uint64_t len = __ldg(a.S.sparseEnd + ex) - __ldg(a.S.sparseStart + ex);
            uint32_t c = (uint32_t)((len + C - 1) / C);
            if (a.fused && c == 0) c = 1;
            sm.prefix[r] = c;
        }
        __syncthreads();
        // block exclusive scan over prefix[0..batch)
        const uint32_t per = (batch + kZThreads - 1) / kZThreads;
        uint32_t lo = min((uint32_t)tid * per, batch), hi = min(lo + per, batch);
        uint32_t local = 0;
        for (uint32_t r = lo; r < hi; r++) local += sm.prefix[r];
        uint32_t incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
            if ((tid & 31) >= o) incl += n;
        }
        if ((tid & 31) == 31) sm.scan[tid >> 5] = incl;
        __syncthreads();
        uint32_t warpBase = 0;
        for (int w = 0; w < (tid >> 5); w++) warpBase += sm.scan[w];
        uint32_t run = warpBase + incl - local;
        __syncthreads();
        for (uint32_t r = lo; r < hi; r++) { uint32_t c = sm.prefix[r]; sm.prefix[r] = run; run += c; }
        if (tid == kZThreads - 1) sm.prefix[batch] = run;
        __syncthreads();
        T = sm.prefix[batch];
        // split rows need T*stride partial floats in the worst case
        if ((unsigned long long)T * a.stride <= a.partialsCap) break;
        if (C >= (uint32_t)kZMaxChunk) {
            // split-row workspace too small even at the largest chunk: refuse loudly (sticky
            // status word in mapped host memory, reported by the next dsb200 call / sync)
            if (blockIdx.x == 0 && tid == 0) *a.status = DSB200_STATUS_Z_WORKSPACE;
            return;
        }
        C = min(C * 2, (uint32_t)kZMaxChunk);
        __syncthreads();
    }

    const uint32_t stride = a.stride;
    const bool producer = tid >= kZConsumers;

    if (producer) {
        // ------------------------------------------------------------ producer warp
        const int lane = tid & 31;
        uint32_t it = 0;
        for (uint32_t t = blockIdx.x; t < T; t += gridDim.x, it++) {
            const int s = it % kZStages;
            const uint32_t ph = (it / kZStages) & 1;
            // locate the row: last r with prefix[r] <= t
            uint32_t lo = 0, hi = batch;
            while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (sm.prefix[mid] <= t) lo = mid; else hi = mid; }
            const uint32_t row = lo, k = t - sm.prefix[row], nChunks = sm.prefix[row + 1] - sm.prefix[row];
            const uint32_t ex = example_of(a.P, a.S.index, a.position, a.rowBase + row);
            const uint64_t rs = __ldg(a.S.sparseStart + ex), re = __ldg(a.S.sparseEnd + ex);
            const uint64_t e0 = rs + (uint64_t)k * C;
            const uint64_t e1 = (re < e0 + C) ? re : e0 + C;
            const uint32_t count = (e1 > e0) ? (uint32_t)(e1 - e0) : 0u;
            const uint64_t floorE = e0 & ~(uint64_t)3;                    // smem slot 0 <-> element floorE
            const uint32_t off = (uint32_t)(e0 - floorE);
            mbar_wait(&sm.empty[s], ph ^ 1);
            uint64_t a4 = (e0 + 3) & ~(uint64_t)3, b4 = e1 & ~(uint64_t)3;
            const bool bulk = a.useTma && b4 > a4;
            const bool valTma = ANALOG && a.S.dataType == DSB200_DT_FLOAT;
            // head [e0,a4) and tail [b4,e1) -- or the whole run -- with plain loads
            auto stage_plain = [&](uint64_t from, uint64_t to) {
                for (uint64_t e = from + lane; e < to; e += 32) {
                    sm.idx[s][e - floorE] = __ldg(a.S.sparseIndex + e);
                    if (DENOISED) sm.rnd[s][e - floorE] = __ldg(a.S.denoisingRandom + e);
                    if (ANALOG && valTma) sm.val[s][e - floorE] = __ldg((const float*)a.S.sparseData + e);
                }
            };
            if (bulk) { stage_plain(e0, a4); stage_plain(b4, e1); }
            else      { stage_plain(e0, e1); }
            if (ANALOG && !valTma)
                for (uint64_t e = e0 + lane; e < e1; e += 32) sm.val[s][e - floorE] = load_value(a.S.sparseData, a.S.dataType, e);
            __syncwarp();
            if (lane == 0) {
                ZMeta m; m.row = row; m.k = k; m.nChunks = nChunks; m.count = count; m.smemOff = off;
                m.w = a.S.dataWeight ? __ldg(a.S.dataWeight + ex) : 1.0f;
                sm.meta[s] = m;
                if (bulk) {
                    const uint32_t bytes = (uint32_t)(b4 - a4) * 4u;
                    const uint32_t narr = 1u + (DENOISED ? 1u : 0u) + (valTma ? 1u : 0u);
                    mbar_arrive_expect_tx(&sm.full[s], bytes * narr);
                    bulk_g2s(&sm.idx[s][a4 - floorE], a.S.sparseIndex + a4, bytes, &sm.full[s]);
                    if (DENOISED) bulk_g2s(&sm.rnd[s][a4 - floorE], a.S.denoisingRandom + a4, bytes, &sm.full[s]);
                    if (valTma)   bulk_g2s(&sm.val[s][a4 - floorE], (const float*)a.S.sparseData + a4, bytes, &sm.full[s]);
                } else {
                    mbar_arrive(&sm.full[s]);
                }
            }
            __syncwarp();
        }
        return;
    }

    // ---------------------------------------------------------------- consumers
    // lanesPerRow = power of two >= ceil(stride/VEC), capped at 256; G groups split the nnz
    const uint32_t cols = (stride + VEC - 1) / VEC;
    uint32_t lpr = 1; while (lpr < cols && lpr < (uint32_t)kZConsumers) lpr <<= 1;
    const uint32_t G = kZConsumers / lpr;
    const uint32_t g = tid / lpr, lane = tid % lpr;
    const uint32_t colBlocks = (cols + lpr - 1) / lpr;
    const float q = a.P.denoising_q, dp = a.P.denoising_p;

    uint32_t it = 0;
    for (uint32_t t = blockIdx.x; t < T; t += gridDim.x, it++) {
        const int s = it % kZStages;
        const uint32_t ph = (it / kZStages) & 1;
        mbar_wait(&sm.full[s], ph);
        const ZMeta m = sm.meta[s];
        const uint32_t* sIdx = &sm.idx[s][m.smemOff];
        const float*    sVal = &sm.val[s][m.smemOff];
        const float*    sRnd = &sm.rnd[s][m.smemOff];
        const float scale = DENOISED ? q * m.w : m.w;
        float* zrow = a.Z + (size_t)(a.rowBase + m.row) * stride;
        const bool multi = m.nChunks > 1;
        float* prow = a.partials + (size_t)t * stride;

        for (uint32_t cb = 0; cb < colBlocks; cb++) {
            const uint32_t col = (cb * lpr + lane) * VEC;          // first column of this lane
            const bool active = col < stride;
            float acc[VEC];
#pragma unroll
            for (int v = 0; v < VEC; v++) acc[v] = 0.0f;
            if (active) {
                const float* wcol = a.W + col;
                uint32_t j = g;
                // 8 independent row gathers in flight per thread
                for (; j + (kZUnroll - 1) * G < m.count; j += kZUnroll * G) {
                    float mult[kZUnroll];
                    float4 x4[kZUnroll]; float x1[kZUnroll];
#pragma unroll
                    for (int u = 0; u < kZUnroll; u++) {
                        const uint32_t jj = j + u * G;
                        const size_t base = (size_t)sIdx[jj] * stride;
                        mult[u] = ANALOG ? sVal[jj] : 1.0f;
                        if (DENOISED && sRnd[jj] < dp) mult[u] = 0.0f;
                        if (VEC == 4) x4[u] = ldg_nc_f4(reinterpret_cast<const float4*>(wcol + base));
                        else          x1[u] = __ldg(wcol + base);
                    }
#pragma unroll
                    for (int u = 0; u < kZUnroll; u++) {
                        if (VEC == 4) {
                            acc[0] = fmaf(x4[u].x, mult[u], acc[0]); acc[1 % VEC] = fmaf(x4[u].y, mult[u], acc[1 % VEC]);
                            acc[2 % VEC] = fmaf(x4[u].z, mult[u], acc[2 % VEC]); acc[3 % VEC] = fmaf(x4[u].w, mult[u], acc[3 % VEC]);
                        } else acc[0] = fmaf(x1[u], mult[u], acc[0]);
                    }
                }
                for (; j < m.count; j += G) {
                    const size_t base = (size_t)sIdx[j] * stride;
                    float mult = ANALOG ? sVal[j] : 1.0f;
                    if (DENOISED && sRnd[j] < dp) mult = 0.0f;
                    if (VEC == 4) {
                        const float4 x = ldg_nc_f4(reinterpret_cast<const float4*>(wcol + base));
                        acc[0] = fmaf(x.x, mult, acc[0]); acc[1 % VEC] = fmaf(x.y, mult, acc[1 % VEC]);
                        acc[2 % VEC] = fmaf(x.z, mult, acc[2 % VEC]); acc[3 % VEC] = fmaf(x.w, mult, acc[3 % VEC]);
                    } else acc[0] = fmaf(__ldg(wcol + base), mult, acc[0]);
                }
            }
            // combine the G nnz-groups in fixed order through shared memory
            if (G > 1) {
                consumer_sync();                                   // previous use of sRed finished
#pragma unroll
                for (int v = 0; v < VEC; v++) sRed[(g * lpr + lane) * VEC + v] = acc[v];
                consumer_sync();
                if (g == 0) {
#pragma unroll
                    for (int v = 0; v < VEC; v++) {
                        float sum = 0.0f;
                        for (uint32_t gg = 0; gg < G; gg++) sum += sRed[(gg * lpr + lane) * VEC + v];
                        acc[v] = sum;
                    }
                }
            }
            if (cb == colBlocks - 1) {                              // staged runs no longer needed
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(&sm.empty[s]);
            }
            if (g == 0 && active) {
                if (!multi) {
#pragma unroll
                    for (int v = 0; v < VEC; v++) {
                        if (col + v < stride) {
                            float zold = a.bias ? __ldg(a.bias + col + v) : ((a.beta == 0.0f) ? 0.0f : a.beta * zrow[col + v]);
                            float z = DENOISED && !ANALOG ? scale * (zold + acc[v]) : fmaf(scale, acc[v], zold);
                            acc[v] = (a.activation >= 0) ? apply_act(a.activation, z) : z;
                        }
                    }
                    if (VEC == 4) *reinterpret_cast<float4*>(zrow + col) = make_float4(acc[0], acc[1 % VEC], acc[2 % VEC], acc[3 % VEC]);
                    else zrow[col] = acc[0];
                } else {
                    if (VEC == 4) *reinterpret_cast<float4*>(prow + col) = make_float4(acc[0], acc[1 % VEC], acc[2 % VEC], acc[3 % VEC]);
                    else prow[col] = acc[0];
                }
            }
        }
        if (multi) {
            // deterministic split-row combine: the last chunk to arrive adds all partials in chunk order
            __threadfence();
            consumer_sync();
            if (tid == 0) {
                const uint32_t old = atomicAdd(a.rowCounters + a.rowBase + m.row, 1u);
                const bool last = (old == m.nChunks - 1);
                if (last) a.rowCounters[a.rowBase + m.row] = 0;    // self-reset for the next launch
                sm.isLast = last ? 1u : 0u;
            }
            consumer_sync();
            if (sm.isLast) {
                __threadfence();
                const float* pbase = a.partials + (size_t)sm.prefix[m.row] * stride;
                for (uint32_t c0 = tid * VEC; c0 < stride; c0 += kZConsumers * VEC) {
                    float sum[VEC];
#pragma unroll
                    for (int v = 0; v < VEC; v++) sum[v] = 0.0f;
                    for (uint32_t kk = 0; kk < m.nChunks; kk++) {
                        if (VEC == 4) {
                            const float4 x = ldg_cg_f4(reinterpret_cast<const float4*>(pbase + (size_t)kk * stride + c0));
                            sum[0] += x.x; sum[1 % VEC] += x.y; sum[2 % VEC] += x.z; sum[3 % VEC] += x.w;
                        } else sum[0] += ldg_cg_f(pbase + (size_t)kk * stride + c0);
                    }
#pragma unroll
                    for (int v = 0; v < VEC; v++) {
                        if (c0 + v < stride) {
                            float zold = a.bias ? __ldg(a.bias + c0 + v) : ((a.beta == 0.0f) ? 0.0f : a.beta * zrow[c0 + v]);
                            float z = DENOISED && !ANALOG ? scale * (zold + sum[v]) : fmaf(scale, sum[v], zold);
                            sum[v] = (a.activation >= 0) ? apply_act(a.activation, z) : z;
                        }
                    }
                    if (VEC == 4) *reinterpret_cast<float4*>(zrow + c0) = make_float4(sum[0], sum[1 % VEC], sum[2 % VEC], sum[3 % VEC]);
                    else zrow[c0] = sum[0];
                }
            }
            consumer_sync();                                       // sm.isLast may be rewritten next item
        }
    }
}

// bias broadcast (kClearUnit, E/kernels.cu:60-80) and bias add (kAddBias, E/kernels.cu:564-584)

// ---------------------------------------------------------------- warp-autonomous kernel (stride % 4 == 0)
// Work item = (row, chunk of <= C nnz), taken by ONE WARP: the item's indices (and values / randoms) are read straight
// into registers with coalesced loads, broadcast lane to lane with shuffles, and every lane gathers its float4 column of
// the weight rows, 8 independent 512-byte row reads in flight per warp-instruction slot.  No shared-memory staging, no
// block barrier per item, up to 64 warps per SM: the whole batch (~2,800 items for BASELINE config 2) is one wave.
// Rows spanning several items are combined exactly as before: partial sums to the context workspace, the last warp to
// arrive (atomic counter per row) adds them in chunk order -- deterministic, no float atomics.
constexpr int kZWThreads = 256;
constexpr int kZWUnroll  = 8;

struct ZWSmem {
    uint32_t prefix[kZMaxRows + 1];
    uint32_t scan[kZWThreads / 32];
};

template <bool ANALOG, bool DENOISED>
__global__ void __launch_bounds__(kZWThreads, 3)
sparse_z_warp_kernel(const ZArgs a)
{
    __shared__ ZWSmem sm;
    const int tid = threadIdx.x;
    const uint32_t batch = a.batch, lane = tid & 31;
    pdl_launch_dependents();
    pdl_wait();

    // ---- plan: chunk counts per row -> exclusive prefix (identical in every CTA); larger chunks if the split-row
    // workspace would overflow
    uint32_t C = (uint32_t)a.chunk;
    uint32_t T = 0;
    for (;;) {
        for (uint32_t r = tid; r < batch; r += kZWThreads) {
            const uint32_t ex = example_of(a.P, a.S.index, a.position, a.rowBase + r);
            const uint64_t len = __ldg(a.S.sparseEnd + ex) - __ldg(a.S.sparseStart + ex);
            uint32_t c = (uint32_t)((len + C - 1) / C);
            if (a.fused && c == 0) c = 1;
            sm.prefix[r] = c;
        }
        __syncthreads();
        const uint32_t per = (batch + kZWThreads - 1) / kZWThreads;
        const uint32_t lo = min((uint32_t)tid * per, batch), hi = min(lo + per, batch);
        uint32_t local = 0;
        for (uint32_t r = lo; r < hi; r++) local += sm.prefix[r];
        uint32_t incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += n;
        }
        if (lane == 31) sm.scan[tid >> 5] = incl;
        __syncthreads();
        uint32_t warpBase = 0;
        for (int w = 0; w < (tid >> 5); w++) warpBase += sm.scan[w];
        uint32_t run = warpBase + incl - local;
        __syncthreads();
        for (uint32_t r = lo; r < hi; r++) { const uint32_t c = sm.prefix[r]; sm.prefix[r] = run; run += c; }
        if (tid == kZWThreads - 1) sm.prefix[batch] = run;
        __syncthreads();
        T = sm.prefix[batch];
        if ((unsigned long long)T * a.stride <= a.partialsCap) break;
        if (C >= 65536u) {
            if (blockIdx.x == 0 && tid == 0) *a.status = DSB200_STATUS_Z_WORKSPACE;
            return;
        }
        C *= 2;
        __syncthreads();
    }

    const uint32_t stride = a.stride;
    const float dp = a.P.denoising_p;
    const uint32_t warpsPerCta = kZWThreads / 32, gw = blockIdx.x * warpsPerCta + (tid >> 5), nw = gridDim.x * warpsPerCta;
    for (uint32_t t = gw; t < T; t += nw) {
        uint32_t lo = 0, hi = batch;                                      // last row r with prefix[r] <= t
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (sm.prefix[mid] <= t) lo = mid; else hi = mid; }
        const uint32_t row = lo, k = t - sm.prefix[row], nChunks = sm.prefix[row + 1] - sm.prefix[row];
        const uint32_t ex = example_of(a.P, a.S.index, a.position, a.rowBase + row);
        const uint64_t rs = __ldg(a.S.sparseStart + ex), re = __ldg(a.S.sparseEnd + ex);
        const uint64_t e0 = rs + (uint64_t)k * C;
        const uint64_t e1 = (re < e0 + C) ? re : e0 + C;
        const float w = a.S.dataWeight ? __ldg(a.S.dataWeight + ex) : 1.0f;
        const float scale = DENOISED ? a.P.denoising_q * w : w;
        float* zrow = a.Z + (size_t)(a.rowBase + row) * stride;
        float* prow = a.partials + (size_t)t * stride;
        const bool multi = nChunks > 1;

        for (uint32_t colBase = 0; colBase < stride; colBase += 128) {   // every lane runs the loop: the shuffles need the whole warp
            const uint32_t col = colBase + lane * 4;
            const bool active = col < stride;
            const float* wcol = a.W + (active ? col : 0u);
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (uint64_t e = e0; e < e1; e += 32) {
                // 32 entries of the item: one per lane
                const uint32_t n = (uint32_t)min((uint64_t)32, e1 - e);
                uint32_t myIdx = 0;
                float myMul = 0.0f;
                if (lane < n) {
                    myIdx = __ldg(a.S.sparseIndex + e + lane);
                    myMul = ANALOG ? load_value(a.S.sparseData, a.S.dataType, e + lane) : 1.0f;
                    if (DENOISED && __ldg(a.S.denoisingRandom + e + lane) < dp) myMul = 0.0f;
                }
                uint32_t j = 0;
                for (; j + kZWUnroll <= n; j += kZWUnroll) {
                    float4 x[kZWUnroll];
                    float m[kZWUnroll];
#pragma unroll
                    for (int u = 0; u < kZWUnroll; u++) {
                        const uint32_t idx = __shfl_sync(0xffffffffu, myIdx, j + u);
                        if (ANALOG || DENOISED) m[u] = __shfl_sync(0xffffffffu, myMul, j + u);
                        x[u] = ldg_nc_f4(reinterpret_cast<const float4*>(wcol + (size_t)idx * stride));
                    }
#pragma unroll
                    for (int u = 0; u < kZWUnroll; u++) {
                        const float mm = (ANALOG || DENOISED) ? m[u] : 1.0f;
                        acc.x = fmaf(x[u].x, mm, acc.x); acc.y = fmaf(x[u].y, mm, acc.y);
                        acc.z = fmaf(x[u].z, mm, acc.z); acc.w = fmaf(x[u].w, mm, acc.w);
                    }
                }
                for (; j < n; j++) {
                    const uint32_t idx = __shfl_sync(0xffffffffu, myIdx, j);
                    const float mm = (ANALOG || DENOISED) ? __shfl_sync(0xffffffffu, myMul, j) : 1.0f;
                    const float4 x = ldg_nc_f4(reinterpret_cast<const float4*>(wcol + (size_t)idx * stride));
                    acc.x = fmaf(x.x, mm, acc.x); acc.y = fmaf(x.y, mm, acc.y);
                    acc.z = fmaf(x.z, mm, acc.z); acc.w = fmaf(x.w, mm, acc.w);
                }
            }
            if (!active) continue;
            if (!multi) {
                float v[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const float zold = a.bias ? __ldg(a.bias + col + q) : ((a.beta == 0.0f) ? 0.0f : a.beta * zrow[col + q]);
                    const float z = DENOISED && !ANALOG ? scale * (zold + v[q]) : fmaf(scale, v[q], zold);
                    v[q] = (a.activation >= 0) ? apply_act(a.activation, z) : z;
                }
                *reinterpret_cast<float4*>(zrow + col) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
                *reinterpret_cast<float4*>(prow + col) = acc;
            }
        }
        if (multi) {
            // deterministic split-row combine by the last warp to arrive
            __threadfence();
            __syncwarp();
            uint32_t last = 0;
            if (lane == 0) {
                const uint32_t old = atomicAdd(a.rowCounters + a.rowBase + row, 1u);
                last = (old == nChunks - 1) ? 1u : 0u;
                if (last) a.rowCounters[a.rowBase + row] = 0;              // self-reset for the next launch
            }
            last = __shfl_sync(0xffffffffu, last, 0);
            if (last) {
                __threadfence();
                const float* pbase = a.partials + (size_t)sm.prefix[row] * stride;
                for (uint32_t col = lane * 4; col < stride; col += 128) {
                    float4 sum = make_float4(0.f, 0.f, 0.f, 0.f);
                    for (uint32_t kk = 0; kk < nChunks; kk++) {
                        const float4 x = ldg_cg_f4(reinterpret_cast<const float4*>(pbase + (size_t)kk * stride + col));
                        sum.x += x.x; sum.y += x.y; sum.z += x.z; sum.w += x.w;
                    }
                    float v[4] = {sum.x, sum.y, sum.z, sum.w};
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const float zold = a.bias ? __ldg(a.bias + col + q) : ((a.beta == 0.0f) ? 0.0f : a.beta * zrow[col + q]);
                        const float z = DENOISED && !ANALOG ? scale * (zold + v[q]) : fmaf(scale, v[q], zold);
                        v[q] = (a.activation >= 0) ? apply_act(a.activation, z) : z;
                    }
                    *reinterpret_cast<float4*>(zrow + col) = make_float4(v[0], v[1], v[2], v[3]);
                }
            }
        }
    }
}

template <bool ANALOG, bool DENOISED>
static int launch_zw(dsb200_ctx* ctx, const ZArgs& a)
{
    int grid = ctx->numSMs * 3;
    const uint32_t want = (a.batch * 3u + 7u) / 8u;                        // ~3 items per row at the ML-20M row lengths
    if ((uint32_t)grid > want) grid = (int)want;
    if (grid < 1) grid = 1;
    DSB_CUDA_OK(launch_pdl(sparse_z_warp_kernel<ANALOG, DENOISED>, dim3(grid), dim3(kZWThreads), 0, ctx->stream, a));
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    return 0;
}

template <bool ADD>
__global__ void __launch_bounds__(256) bias_kernel(float* __restrict__ unit, const float* __restrict__ bias, uint32_t stride, uint64_t size)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < size; i += (uint64_t)gridDim.x * blockDim.x) {
        const float b = __ldg(bias + (uint32_t)(i % stride));
        unit[i] = ADD ? unit[i] + b : b;
    }
}

static size_t z_smem_bytes(uint32_t stride, int vec)
{
    uint32_t cols = (stride + vec - 1) / vec;
    uint32_t lpr = 1; while (lpr < cols && lpr < (uint32_t)kZConsumers) lpr <<= 1;
    uint32_t G = kZConsumers / lpr;
    size_t red = (G > 1) ? (size_t)kZConsumers * vec * sizeof(float) : 0;
    return sizeof(ZSmem) + red;
}

template <int VEC, bool ANALOG, bool DENOISED>
static int launch_z(dsb200_ctx* ctx, const ZArgs& a, int grid)
{
    auto kern = sparse_z_kernel<VEC, ANALOG, DENOISED>;
    size_t smem = z_smem_bytes(a.stride, VEC);
    DSB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kZThreads, smem, ctx->stream>>>(a);
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    return 0;
}

int sparse_z_impl(dsb200_ctx* ctx, const dsb200_sparse* s, uint32_t position, uint32_t batch, uint32_t stride,
                  const float* W, const float* bias, int activation, float* Z, float beta, int denoised, int fused)
{
    if (!ctx || !s || !W || !Z || stride == 0) return fail(ctx, DSB200_EINVAL, "sparse_z: null argument");
    if (!s->sparseStart || !s->sparseEnd || !s->sparseIndex) return fail(ctx, DSB200_EINVAL, "sparse_z: CSR arrays missing");
    if (denoised && !s->denoisingRandom) return fail(ctx, DSB200_EINVAL, "sparse_z: denoised without random buffer");
    if (batch == 0) return 0;
    int rc = dsb200_ctx_reserve(ctx, batch, 0);
    if (rc) return rc;

    const bool analog = s->sparseData != nullptr;
    const bool vec4 = (stride % 4 == 0) && (((uintptr_t)W | (uintptr_t)Z) % 16 == 0);
    const int vec = vec4 ? 4 : 1;
    uint32_t cols = (stride + vec - 1) / vec;
    uint32_t lpr = 1; while (lpr < cols && lpr < (uint32_t)kZConsumers) lpr <<= 1;
    const uint32_t G = kZConsumers / lpr;
    int chunk = (int)(32 * G);
    if (chunk < 64) chunk = 64;
    if (chunk > kZMaxChunk) chunk = kZMaxChunk;

    for (uint32_t base = 0; base < batch; base += kZMaxRows) {
        ZArgs a;
        a.P = ctx->params; a.S = *s;
        a.position = position; a.batch = (batch - base < (uint32_t)kZMaxRows) ? batch - base : (uint32_t)kZMaxRows;
        a.rowBase = base; a.stride = stride;
        a.W = W; a.bias = bias; a.Z = Z; a.beta = beta; a.activation = activation; a.chunk = chunk; a.fused = fused;
        a.useTma = !ctx->noTma && (((uintptr_t)s->sparseIndex % 16) == 0)
                   && (!denoised || ((uintptr_t)s->denoisingRandom % 16) == 0)
                   && (!analog || s->dataType != DSB200_DT_FLOAT || ((uintptr_t)s->sparseData % 16) == 0);
        a.rowCounters = ctx->dRowCounters; a.partials = ctx->dPartials; a.partialsCap = ctx->partialsCap;
        a.status = ctx->dStatus;
        int grid = ctx->numSMs * 2;
        if ((uint32_t)grid > a.batch * 4u) grid = (int)(a.batch * 4u);
        if (grid < 1) grid = 1;
        if (vec4 && !ctx->zStagedKernel) {                                  // warp-autonomous kernel, 64-nnz items
            a.chunk = 64;
            if (analog) rc = denoised ? launch_zw<true, true>(ctx, a) : launch_zw<true, false>(ctx, a);
            else        rc = denoised ? launch_zw<false, true>(ctx, a) : launch_zw<false, false>(ctx, a);
        } else if (vec4) {
            if (analog) rc = denoised ? launch_z<4, true, true>(ctx, a, grid) : launch_z<4, true, false>(ctx, a, grid);
            else        rc = denoised ? launch_z<4, false, true>(ctx, a, grid) : launch_z<4, false, false>(ctx, a, grid);
        } else {
            if (analog) rc = denoised ? launch_z<1, true, true>(ctx, a, grid) : launch_z<1, true, false>(ctx, a, grid);
            else        rc = denoised ? launch_z<1, false, true>(ctx, a, grid) : launch_z<1, false, false>(ctx, a, grid);
        }
        if (rc) return rc;
    }
    return 0;
}

}  // namespace dsb

extern "C" {

int dsb200_sparse_z(dsb200_ctx* ctx, const dsb200_sparse* s, uint32_t position, uint32_t batch, uint32_t stride,
                    const float* pWeight, float* pUnit, float beta, int denoised)
{
    DSB_PROFILE(ctx, "sparse_z");
    return dsb::sparse_z_impl(ctx, s, position, batch, stride, pWeight, nullptr, -1, pUnit, beta, denoised, 0);
}

int dsb200_sparse_z_bias_act(dsb200_ctx* ctx, const dsb200_sparse* s, uint32_t position, uint32_t batch, uint32_t stride,
                             const float* pWeight, const float* pBias, int activation, float* pUnit, int denoised)
{
    DSB_PROFILE(ctx, "sparse_z_bias_act");
    if (!pBias) return dsb::fail(ctx, DSB200_EINVAL, "sparse_z_bias_act: bias missing");
    if (activation != DSB200_ACT_SIGMOID && activation != DSB200_ACT_TANH && activation != DSB200_ACT_RELU &&
        activation != DSB200_ACT_LINEAR)
        return dsb::fail(ctx, DSB200_EUNSUPPORTED, "sparse_z_bias_act: activation not fusable");
    return dsb::sparse_z_impl(ctx, s, position, batch, stride, pWeight, pBias, activation, pUnit, 1.0f, denoised, 1);
}

int dsb200_clear_unit(dsb200_ctx* ctx, float* pUnit, const float* pBias, uint32_t stride, uint32_t batch)
{
    DSB_PROFILE(ctx, "clear_unit");
    if (!ctx || !pUnit || !pBias) return dsb::fail(ctx, DSB200_EINVAL, "clear_unit: null argument");
    uint64_t size = (uint64_t)stride * batch;
    if (!size) return 0;
    int grid = (int)((size + 255) / 256); if (grid > ctx->numSMs * 8) grid = ctx->numSMs * 8;
    dsb::bias_kernel<false><<<grid, 256, 0, ctx->stream>>>(pUnit, pBias, stride, size);
    dsb::count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    return 0;
}

int dsb200_add_bias(dsb200_ctx* ctx, float* pUnit, const float* pBias, uint32_t stride, uint32_t batch)
{
    DSB_PROFILE(ctx, "add_bias");
    if (!ctx || !pUnit || !pBias) return dsb::fail(ctx, DSB200_EINVAL, "add_bias: null argument");
    uint64_t size = (uint64_t)stride * batch;
    if (!size) return 0;
    int grid = (int)((size + 255) / 256); if (grid > ctx->numSMs * 8) grid = ctx->numSMs * 8;
    dsb::bias_kernel<true><<<grid, 256, 0, ctx->stream>>>(pUnit, pBias, stride, size);
    dsb::count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // extern "C"
