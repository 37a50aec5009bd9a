// topk.cu -- prediction top-K with in-kernel exclusion filtering (hot-path row a13).
//
// Replaces kCalculateTopK 3-arg / 4-arg (E/kernels.cu:3201-4385, E/bitonic.h) and the host-side
// filter round trip of the recommendation generator (U/NNRecsGenerator.cpp:132-150: D2H of the
// whole [batch][N] score matrix, score *= 0 on the host via U/Filters.cpp:49-67, H2D, top-K).
//
// Reference: ONE warp streams a whole row (4 MB for N = 1M) with 128-byte loads and a 2k-wide
// register bitonic network; 4,096 rows = 4,096 warps on the whole GPU, latency bound.
// Here a row is cut into segments so that the grid holds >= 8 CTAs per SM:
//   pass 1  CTA = (row, segment): streams the segment with 128-bit loads (four in flight per thread), zeroes the
//           scores named by the row's exclusion list through a shared-memory bitmap, keeps candidates strictly
//           above the running k-th key in a 4,096-entry shared buffer that is cut back to the k best with an
//           O(n) radix select whenever it fills, and emits the segment's top-k (only those k are sorted);
//   pass 2  CTA = row: the same selection over the segs*k candidates.
// Rule (fixed, unlike the reference's arrival-order ties): descending key, ties by ascending
// column; a candidate must be > -MAX_VALUE; unused slots hold (-MAX_VALUE, 0) (E/NNTypes.h:49).
#include <algorithm>

#include "common.cuh"
#include "launch.h"

namespace dsb {

constexpr int      kKThreads = 256;
constexpr uint32_t kKCap     = 4096;            // candidate buffer
constexpr uint32_t kKChunk   = 1024;            // elements per block step (4 per thread)
constexpr uint32_t kKMaxSeg  = 131072;          // bitmap = 16 KB
constexpr uint32_t kKMaxK    = 1024;

struct KArgs {
    const float* key; const uint32_t* value;       // value == NULL: payload = position
    uint32_t batch, width, k, segs, segLen;
    const uint64_t* fStart; const uint64_t* fEnd; const uint32_t* fIndex;
    float* outKey; uint32_t* outValue;             // [batch][segs][k] (pass 1) or [batch][k]
    uint32_t* outPos;                              // pass 1 with value: not needed (payload carried)
};

__device__ __forceinline__ bool before(float ka, uint32_t pa, float kb, uint32_t pb)
{
    return (ka > kb) || (ka == kb && pa < pb);
}

// sorts sKey/sPos/sVal[0..n) (n <= kKCap) so that "before" elements come first; pads to a power of two
__device__ void block_sort(float* sKey, uint32_t* sPos, uint32_t* sVal, uint32_t n, bool hasVal)
{
    const uint32_t tid = threadIdx.x;
    uint32_t p2 = 2; while (p2 < n) p2 <<= 1;
    for (uint32_t i = n + tid; i < p2; i += kKThreads) { sKey[i] = -INFINITY; sPos[i] = 0xffffffffu; if (hasVal) sVal[i] = 0; }
    __syncthreads();
    for (uint32_t ksz = 2; ksz <= p2; ksz <<= 1) {
        for (uint32_t j = ksz >> 1; j > 0; j >>= 1) {
            for (uint32_t i = tid; i < p2; i += kKThreads) {
                const uint32_t x = i ^ j;
                if (x > i) {
                    const bool asc = (i & ksz) == 0;          // "asc" block: best first
                    const bool wrong = asc ? before(sKey[x], sPos[x], sKey[i], sPos[i]) : before(sKey[i], sPos[i], sKey[x], sPos[x]);
                    if (wrong) {
                        const float tk = sKey[i]; sKey[i] = sKey[x]; sKey[x] = tk;
                        const uint32_t tp = sPos[i]; sPos[i] = sPos[x]; sPos[x] = tp;
                        if (hasVal) { const uint32_t tv = sVal[i]; sVal[i] = sVal[x]; sVal[x] = tv; }
                    }
                }
            }
            __syncthreads();
        }
    }
}

// Keeps the k best of sKey/sPos/sVal[0..n) (k < n <= kKCap) in slots [0, k), unsorted, and returns the k-th best key.
// An O(n) radix select instead of sorting the whole buffer (the bitonic network over 4,096 entries costs ~30k
// instructions per thread and ran ~3 times per 512 KB tile -- it, not the streaming, bounded the kernel): the order
// "key descending, position ascending" is the descending order of the 64-bit composite (sortable key bits, ~position),
// whose k-th largest value is found byte by byte from the top with a 256-bin shared histogram; the entries at or above
// it that sit beyond slot k then move into the holes left below slot k.
__device__ __forceinline__ unsigned long long composite(float key, uint32_t pos)
{
    uint32_t u = __float_as_uint(key);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)u << 32) | (unsigned long long)(0xFFFFFFFFu - pos);
}

__device__ float block_select(float* sKey, uint32_t* sPos, uint32_t* sVal, uint32_t n, uint32_t k, bool hasVal, uint32_t* sHist /*256 + 4*/,
                              uint32_t* sMove /*2 * k*/)
{
    const uint32_t tid = threadIdx.x;
    unsigned long long prefix = 0, maskHigh = 0;
    uint32_t want = k;
    for (int b = 7; b >= 0; b--) {
        for (uint32_t i = tid; i < 256; i += kKThreads) sHist[i] = 0;
        __syncthreads();
        for (uint32_t i = tid; i < n; i += kKThreads) {
            const unsigned long long c = composite(sKey[i], sPos[i]);
            if ((c & maskHigh) == prefix) atomicAdd(&sHist[(uint32_t)(c >> (8 * b)) & 255u], 1u);
        }
        __syncthreads();
        if (tid < 32) {
            // lane l owns digits 255 - 8l .. 248 - 8l (descending); find the digit where the running count reaches `want`
            uint32_t local[8], sum = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) { local[q] = sHist[255 - (tid * 8 + q)]; sum += local[q]; }
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t nb = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= (uint32_t)o) incl += nb; }
            const uint32_t excl = incl - sum;
            if (excl < want && want <= incl) {
                uint32_t run = excl;
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    if (run < want && want <= run + local[q]) { sHist[256] = 255 - (tid * 8 + q); sHist[257] = want - run; }
                    run += local[q];
                }
            }
        }
        __syncthreads();
        prefix |= (unsigned long long)sHist[256] << (8 * b);
        maskHigh |= 0xFFull << (8 * b);
        want = sHist[257];
        __syncthreads();
    }
    // prefix = the k-th best composite; exactly k entries are >= prefix
    if (tid == 0) { sHist[258] = 0; sHist[259] = 0; }
    __syncthreads();
    for (uint32_t i = tid; i < n; i += kKThreads) {
        const bool keep = composite(sKey[i], sPos[i]) >= prefix;
        if (i < k && !keep) sMove[atomicAdd(&sHist[258], 1u)] = i;             // hole
        if (i >= k && keep) sMove[k + atomicAdd(&sHist[259], 1u)] = i;         // mover
    }
    __syncthreads();
    const uint32_t moves = sHist[258];                                          // == sHist[259]
    for (uint32_t j = tid; j < moves; j += kKThreads) {
        const uint32_t dst = sMove[j], src = sMove[k + j];
        sKey[dst] = sKey[src]; sPos[dst] = sPos[src];
        if (hasVal) sVal[dst] = sVal[src];
    }
    __syncthreads();
    uint32_t u = (uint32_t)(prefix >> 32);
    u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u;
    return __uint_as_float(u);
}

template <bool HAS_VALUE, bool HAS_FILTER>
__global__ void __launch_bounds__(kKThreads)
topk_kernel(const KArgs a)
{
    extern __shared__ __align__(16) unsigned char smemRaw[];
    float*    sKey = reinterpret_cast<float*>(smemRaw);
    uint32_t* sPos = reinterpret_cast<uint32_t*>(sKey + kKCap);
    uint32_t* sVal = sPos + kKCap;                                    // only when HAS_VALUE
    uint32_t* sHist = HAS_VALUE ? sVal + kKCap : sVal;                // 260 words: radix-select histogram + scratch
    uint32_t* sMove = sHist + 260;                                    // 2 * k words
    uint32_t* sBits = sMove + 2 * a.k;                                // only when HAS_FILTER
    __shared__ uint32_t sCount;
    __shared__ float sThr;

    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint64_t tiles = (uint64_t)a.batch * a.segs;
    for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const uint32_t b = (uint32_t)(tile / a.segs), seg = (uint32_t)(tile % a.segs);
        const uint32_t c0 = seg * a.segLen, c1 = min(c0 + a.segLen, a.width);
        const float* row = a.key + (size_t)b * a.width;
        const uint32_t* vrow = HAS_VALUE ? a.value + (size_t)b * a.width : nullptr;
        if (tid == 0) { sCount = 0; sThr = -kMaxValue; }
        if (HAS_FILTER) {
            const uint32_t words = (c1 - c0 + 31) / 32;
            for (uint32_t i = tid; i < words; i += kKThreads) sBits[i] = 0;
            __syncthreads();
            const uint64_t fs = __ldg(a.fStart + b), fe = __ldg(a.fEnd + b);
            for (uint64_t j = fs + tid; j < fe; j += kKThreads) {
                const uint32_t c = __ldg(a.fIndex + j);
                if (c >= c0 && c < c1) atomicOr(&sBits[(c - c0) >> 5], 1u << ((c - c0) & 31));
            }
        }
        __syncthreads();

        // element i of the row is 16-byte aligned in global memory iff (rowBase + i) % 4 == 0
        const uint32_t mis = (uint32_t)((((uintptr_t)row) >> 2) & 3);
        uint32_t budget = kKCap;                                       // free slots guaranteed before next check
        constexpr int kDepth = 4;                                      // chunks whose loads are issued together
        for (uint32_t base0 = c0; base0 < c1; base0 += kDepth * kKChunk) {
            // issue the loads of up to four chunks first: four independent 128-bit loads in flight per thread
            float kx[kDepth][4];
#pragma unroll
            for (int d = 0; d < kDepth; d++) {
                const uint32_t p = base0 + d * kKChunk + tid * 4;
                if (p + 3 < c1 && ((p + mis) & 3) == 0) {
                    const float4 x = ldg_cs_f4(reinterpret_cast<const float4*>(row + p));
                    kx[d][0] = x.x; kx[d][1] = x.y; kx[d][2] = x.z; kx[d][3] = x.w;
                } else {
#pragma unroll
                    for (int v = 0; v < 4; v++) kx[d][v] = (p + v < c1) ? __ldg(row + p + v) : -INFINITY;
                }
            }
#pragma unroll
            for (int d = 0; d < kDepth; d++) {
                const uint32_t base = base0 + d * kKChunk;
                if (base >= c1) break;
                if (budget < kKChunk) {
                    __syncthreads();
                    const uint32_t n = sCount;
                    if (n > kKCap - kKChunk) {                        // n > k here: cut back to the k best, new threshold = k-th key
                        const float kth = block_select(sKey, sPos, sVal, n, a.k, HAS_VALUE, sHist, sMove);
                        if (tid == 0) { sCount = a.k; sThr = kth; }
                        __syncthreads();
                    }
                    budget = kKCap - sCount;
                    __syncthreads();
                }
                budget -= kKChunk;
                const float thr = sThr;
                const uint32_t p = base + tid * 4;
                if (HAS_FILTER && p < c1) {
                    // c0 and p are multiples of 4: the four exclusion bits of this float4 sit in one bitmap word
                    const uint32_t r = p - c0;
                    const uint32_t bits = (sBits[r >> 5] >> (r & 31)) & 0xFu;
                    if (bits) {
#pragma unroll
                        for (int v = 0; v < 4; v++)
                            if ((bits >> v) & 1u) kx[d][v] *= 0.0f;                       // U/Filters.cpp:49-67: score *= 0
                    }
                }
                uint32_t mask = 0;
#pragma unroll
                for (int v = 0; v < 4; v++)
                    if (p + v < c1 && kx[d][v] > thr) mask |= 1u << v;
                if (!__any_sync(0xffffffffu, mask != 0)) continue;       // common case once the threshold has settled
                // warp-aggregated append
                const uint32_t cnt = __popc(mask);
                uint32_t incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t nb = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += nb; }
                const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
                uint32_t wbase = 0;
                if (lane == 31) wbase = atomicAdd(&sCount, total);
                wbase = __shfl_sync(0xffffffffu, wbase, 31);
                uint32_t o = wbase + incl - cnt;
#pragma unroll
                for (int v = 0; v < 4; v++) {
                    if (mask & (1u << v)) {
                        sKey[o] = kx[d][v]; sPos[o] = p + v;
                        if (HAS_VALUE) sVal[o] = __ldg(vrow + p + v);
                        o++;
                    }
                }
            }
        }
        __syncthreads();
        uint32_t n = sCount;
        if (n > a.k) { block_select(sKey, sPos, sVal, n, a.k, HAS_VALUE, sHist, sMove); n = a.k; }
        block_sort(sKey, sPos, sVal, n, HAS_VALUE);                       // <= k (<= 1,024) entries
        float* ok = a.outKey + ((size_t)b * a.segs + seg) * a.k;
        uint32_t* ov = a.outValue + ((size_t)b * a.segs + seg) * a.k;
        for (uint32_t i = tid; i < a.k; i += kKThreads) {
            if (i < n) { ok[i] = sKey[i]; ov[i] = HAS_VALUE ? sVal[i] : sPos[i]; }
            else       { ok[i] = -kMaxValue; ov[i] = 0; }
        }
        __syncthreads();
    }
}

static int topk_impl(dsb200_ctx* ctx, const float* key, const uint32_t* value, uint32_t batch, uint32_t width, uint32_t k,
                     const uint64_t* fs, const uint64_t* fe, const uint32_t* fi, float* outKey, uint32_t* outValue)
{
    if (!ctx || !key || !outKey || !outValue) return fail(ctx, DSB200_EINVAL, "topk: null argument");
    if (k == 0 || k > kKMaxK) return fail(ctx, DSB200_EINVAL, "topk: k must be in [1, 1024]");
    if ((fs || fe || fi) && !(fs && fe && fi)) return fail(ctx, DSB200_EINVAL, "topk: incomplete filter");
    if (!batch || !width) return 0;
    // segments: enough tiles to fill the GPU, each >= 16K elements, <= 128K (bitmap), multiple of the chunk
    uint32_t segs = 1;
    const uint64_t wantTiles = (uint64_t)ctx->numSMs * 8;
    if (batch < wantTiles) segs = (uint32_t)((wantTiles + batch - 1) / batch);
    const uint32_t maxSegs = (width + 16383) / 16384;
    if (segs > maxSegs) segs = maxSegs;
    const uint32_t minSegs = (width + kKMaxSeg - 1) / kKMaxSeg;
    if (segs < minSegs) segs = minSegs;
    if (segs < 1) segs = 1;
    uint32_t segLen = (width + segs - 1) / segs;
    segLen = ((segLen + kKChunk - 1) / kKChunk) * kKChunk;
    segs = (width + segLen - 1) / segLen;

    KArgs a{};
    a.key = key; a.value = value; a.batch = batch; a.width = width; a.k = k; a.segs = segs; a.segLen = segLen;
    a.fStart = fs; a.fEnd = fe; a.fIndex = fi;
    float* candKey = outKey; uint32_t* candVal = outValue;
    if (segs > 1) {
        // candidates live in the context workspace: [batch][segs][k] keys then values
        const size_t need = (size_t)batch * segs * k * 2;
        int rc = dsb200_ctx_reserve(ctx, 0, need);
        if (rc) return rc;
        candKey = ctx->dPartials;
        candVal = reinterpret_cast<uint32_t*>(ctx->dPartials + (size_t)batch * segs * k);
    }
    a.outKey = candKey; a.outValue = candVal;
    const bool hasVal = value != nullptr, hasFilter = fs != nullptr;
    size_t smem = (size_t)kKCap * 8 + (hasVal ? (size_t)kKCap * 4 : 0) + (260 + 2 * (size_t)k) * 4 + (hasFilter ? (size_t)(segLen / 32 + 1) * 4 : 0);
    uint64_t grid = (uint64_t)batch * segs;
    const uint64_t cap = (uint64_t)ctx->numSMs * 6;
    if (grid > cap) grid = cap;
#define DSB_TOPK(V, F)                                                                                           \
    do {                                                                                                         \
        DSB_CUDA_OK(cudaFuncSetAttribute(topk_kernel<V, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        topk_kernel<V, F><<<(unsigned)grid, kKThreads, smem, ctx->stream>>>(a);                                  \
    } while (0)
    if (hasVal) { if (hasFilter) DSB_TOPK(true, true); else DSB_TOPK(true, false); }
    else        { if (hasFilter) DSB_TOPK(false, true); else DSB_TOPK(false, false); }
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    if (segs > 1) {
        KArgs m{};
        m.key = candKey; m.value = candVal; m.batch = batch; m.width = segs * k; m.k = k; m.segs = 1;
        m.segLen = ((segs * k + kKChunk - 1) / kKChunk) * kKChunk;
        m.outKey = outKey; m.outValue = outValue;
        smem = (size_t)kKCap * 12 + (260 + 2 * (size_t)k) * 4;
        grid = batch; if (grid > cap) grid = cap;
        DSB_CUDA_OK(cudaFuncSetAttribute(topk_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        topk_kernel<true, false><<<(unsigned)grid, kKThreads, smem, ctx->stream>>>(m);
        count_launch();
        DSB_CUDA_OK(cudaGetLastError());
    }
#undef DSB_TOPK
    return 0;
}

// value[i] += offset: turns the column ids of a rank's local top-K into global ids before the cross-rank merge
__global__ void __launch_bounds__(256)
topk_offset_kernel(uint32_t* __restrict__ value, uint64_t n, uint32_t offset)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) value[i] += offset;
}

}  // namespace dsb

extern "C" {

int dsb200_topk_offset(dsb200_ctx* ctx, uint32_t* pValue, uint64_t n, uint32_t offset)
{
    DSB_PROFILE(ctx, "topk_offset");
    if (!ctx || !pValue) return dsb::fail(ctx, DSB200_EINVAL, "topk_offset: null argument");
    if (!n || !offset) return 0;
    const unsigned blocks = (unsigned)std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->numSMs * 8);
    dsb::topk_offset_kernel<<<blocks, 256, 0, ctx->stream>>>(pValue, n, offset);
    dsb::count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    return 0;
}

int dsb200_topk(dsb200_ctx* ctx, const float* pScores, uint32_t batch, uint32_t width, uint32_t k,
                const uint64_t* fs, const uint64_t* fe, const uint32_t* fi, float* pOutKey, uint32_t* pOutValue)
{
    DSB_PROFILE(ctx, "topk");
    return dsb::topk_impl(ctx, pScores, nullptr, batch, width, k, fs, fe, fi, pOutKey, pOutValue);
}

int dsb200_topk_kv(dsb200_ctx* ctx, const float* pKey, const uint32_t* pValue, uint32_t batch, uint32_t width, uint32_t k,
                   float* pOutKey, uint32_t* pOutValue)
{
    DSB_PROFILE(ctx, "topk_kv");
    if (!pValue) return dsb::fail(ctx, DSB200_EINVAL, "topk_kv: null value array");
    return dsb::topk_impl(ctx, pKey, pValue, batch, width, k, nullptr, nullptr, nullptr, pOutKey, pOutValue);
}

}  // extern "C"
