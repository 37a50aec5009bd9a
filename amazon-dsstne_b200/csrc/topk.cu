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
constexpr uint32_t kKHistWords = 520;           // block_select: two histograms + five scratch words

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
    if (n <= (uint32_t)kKThreads) {
        // the usual case (k = 100): every thread counts the entries that come before its own and moves it there -- two barriers
        // instead of the 28 of the bitonic network below (barriers, not bytes, bounded this kernel: profiles/r2_notes.md)
        float key = 0.0f; uint32_t pos = 0, val = 0, rank = 0;
        if (tid < n) {
            key = sKey[tid]; pos = sPos[tid]; if (hasVal) val = sVal[tid];
            for (uint32_t j = 0; j < n; j++) rank += before(sKey[j], sPos[j], key, pos) ? 1u : 0u;
        }
        __syncthreads();
        if (tid < n) { sKey[rank] = key; sPos[rank] = pos; if (hasVal) sVal[rank] = val; }
        __syncthreads();
        return;
    }
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

__device__ float block_select(float* sKey, uint32_t* sPos, uint32_t* sVal, uint32_t n, uint32_t k, bool hasVal, uint32_t* sHist /*kKHistWords*/,
                              uint32_t* sMove /*2 * k*/)
{
    // sHist: two 256-bin histograms used alternately (the next pass's is zeroed while warp 0 scans the current one: two barriers
    // per pass), then [512] digit, [513] remaining rank, [514] size of the digit's bin, [515] holes, [516] movers
    const uint32_t tid = threadIdx.x;
    unsigned long long prefix = 0, maskHigh = 0;
    uint32_t want = k;
    for (uint32_t i = tid; i < 256; i += kKThreads) sHist[i] = 0;
    if (tid == 0) { sHist[515] = 0; sHist[516] = 0; }
    __syncthreads();
    for (int b = 7; b >= 0; b--) {
        uint32_t* cur = sHist + (((7 - b) & 1) ? 256 : 0);
        uint32_t* nxt = sHist + (((7 - b) & 1) ? 0 : 256);
        // The survivors of a settled threshold share their leading bytes (keys in [0.97, 1) of uniform scores: the top two or three
        // digits are the same for all of them), and thousands of plain atomics on ONE shared-memory word serialise: a warp whose
        // live lanes all hold the same digit adds its count with a single atomic.
        for (uint32_t i0 = 0; i0 < n; i0 += kKThreads) {
            const uint32_t i = i0 + tid;
            bool in = false; uint32_t digit = 0;
            if (i < n) {
                const unsigned long long c = composite(sKey[i], sPos[i]);
                in = (c & maskHigh) == prefix;
                digit = (uint32_t)(c >> (8 * b)) & 255u;
            }
            const unsigned act = __ballot_sync(0xffffffffu, in);
            if (act == 0) continue;
            const uint32_t d0 = __shfl_sync(0xffffffffu, digit, __ffs(act) - 1);
            if (__all_sync(0xffffffffu, !in || digit == d0)) {
                if ((tid & 31u) == (uint32_t)(__ffs(act) - 1)) atomicAdd(&cur[d0], (uint32_t)__popc(act));
            } else if (in) atomicAdd(&cur[digit], 1u);
        }
        __syncthreads();
        if (tid < 32) {
            // lane l owns digits 255 - 8l .. 248 - 8l (descending); find the digit where the running count reaches `want`
            uint32_t local[8], sum = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) { local[q] = cur[255 - (tid * 8 + q)]; sum += local[q]; }
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t nb = __shfl_up_sync(0xffffffffu, incl, o); if (tid >= (uint32_t)o) incl += nb; }
            const uint32_t excl = incl - sum;
            if (excl < want && want <= incl) {
                uint32_t run = excl;
#pragma unroll
                for (int q = 0; q < 8; q++) {
                    if (run < want && want <= run + local[q]) { sHist[512] = 255 - (tid * 8 + q); sHist[513] = want - run; sHist[514] = local[q]; }
                    run += local[q];
                }
            }
        }
        for (uint32_t i = tid; i < 256; i += kKThreads) nxt[i] = 0;               // last read by warp 0 one pass ago, before a barrier
        __syncthreads();
        prefix |= (unsigned long long)sHist[512] << (8 * b);
        maskHigh |= 0xFFull << (8 * b);
        want = sHist[513];
        // the digit's whole bin is wanted: the remaining bytes cannot change the selection (with distinct keys: after the key bytes
        // at the latest -- the position bytes only break ties).  sHist[512..514] are rewritten after the next pass's first barrier.
        if (want == sHist[514]) break;
    }
    // prefix = the resolved leading bytes of the k-th best composite; exactly k entries have (composite & maskHigh) >= prefix
    for (uint32_t i = tid; i < n; i += kKThreads) {
        const bool keep = (composite(sKey[i], sPos[i]) & maskHigh) >= prefix;
        if (i < k && !keep) sMove[atomicAdd(&sHist[515], 1u)] = i;             // hole
        if (i >= k && keep) sMove[k + atomicAdd(&sHist[516], 1u)] = i;         // mover
    }
    __syncthreads();
    const uint32_t moves = sHist[515];                                          // == sHist[516]
    for (uint32_t j = tid; j < moves; j += kKThreads) {
        const uint32_t dst = sMove[j], src = sMove[k + j];
        sKey[dst] = sKey[src]; sPos[dst] = sPos[src];
        if (hasVal) sVal[dst] = sVal[src];
    }
    __syncthreads();
    // the key of the prefix (unresolved low bytes zero: a lower bound of the k-th key, which is all a threshold has to be)
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
    uint32_t* sHist = HAS_VALUE ? sVal + kKCap : sVal;                // kKHistWords: radix-select histograms + scratch
    uint32_t* sMove = sHist + kKHistWords;                            // 2 * k words
    uint32_t* sBits = sMove + 2 * a.k;                                // only when HAS_FILTER
    __shared__ uint32_t sCnt[3];          // appended in period p: sCnt[p % 3] (see the barrier in `consume`)
    __shared__ float sThr;

    const uint32_t tid = threadIdx.x, lane = tid & 31;
    const uint64_t tiles = (uint64_t)a.batch * a.segs;
    for (uint64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const uint32_t b = (uint32_t)(tile / a.segs), seg = (uint32_t)(tile % a.segs);
        const uint32_t c0 = seg * a.segLen, c1 = min(c0 + a.segLen, a.width);
        const float* row = a.key + (size_t)b * a.width;
        const uint32_t* vrow = HAS_VALUE ? a.value + (size_t)b * a.width : nullptr;
        if (tid == 0) { sCnt[0] = 0; sCnt[1] = 0; sCnt[2] = 0; sThr = -kMaxValue; }
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
        uint32_t held = 0, period = 0;                                 // entries in the buffer at the start of the period (uniform)
        constexpr int kDepth = 4;                                      // chunks whose loads are issued together
        constexpr uint32_t kStep = kDepth * kKChunk;
        // the loads of a step: four independent 128-bit loads per thread; elements past the segment read as -inf (never candidates)
        auto fetch = [&](float (&kx)[kDepth][4], uint32_t base0) {
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
        };
        // the exclusion filter on the values of a step (U/Filters.cpp:49-67: score *= 0)
        auto exclude = [&](float (&kx)[kDepth][4], uint32_t base0) {
            if (!HAS_FILTER) return;
#pragma unroll
            for (int d = 0; d < kDepth; d++) {
                const uint32_t p = base0 + d * kKChunk + tid * 4;
                if (p >= c1) break;
                // c0 and p are multiples of 4: the four exclusion bits of this float4 sit in one bitmap word
                const uint32_t r = p - c0;
                const uint32_t bits = (sBits[r >> 5] >> (r & 31)) & 0xFu;
                if (bits) {
#pragma unroll
                    for (int v = 0; v < 4; v++)
                        if ((bits >> v) & 1u) kx[d][v] *= 0.0f;
                }
            }
        };
        auto consume = [&](float (&kx)[kDepth][4], uint32_t base0) {
#pragma unroll
            for (int d = 0; d < kDepth; d++) {
                const uint32_t base = base0 + d * kKChunk;
                if (base >= c1) break;
                if (budget < kKChunk) {
                    // ONE barrier per check.  The appends of a period go to sCnt[period % 3]; after the barrier nobody adds to that
                    // counter any more, so every thread reads the same value without a second barrier; thread 0 clears the counter of
                    // the period after next (last read one barrier ago, first written one barrier from now).
                    __syncthreads();
                    held += sCnt[period % 3];
                    period++;
                    if (tid == 0) sCnt[(period + 1) % 3] = 0;
                    if (held > kKCap - kKChunk) {                     // held > k here: cut back to the k best, new threshold = k-th key
                        const float kth = block_select(sKey, sPos, sVal, held, a.k, HAS_VALUE, sHist, sMove);
                        held = a.k;
                        if (tid == 0) sThr = kth;
                        __syncthreads();
                    }
                    budget = kKCap - held;
                }
                budget -= kKChunk;
                const float thr = sThr;
                const uint32_t p = base + tid * 4;
                // Append by ballots.  With k = 100 the threshold sits at the 97.6th percentile after the first select, so 19 out of 20
                // warps DO hold a candidate among their 128 elements of a chunk: the append is the common path (round-2 ncu: a prefix
                // scan by shuffles here was most of the kernel's 1.2 G warp instructions) -- four votes, one atomic, popc offsets.
                const unsigned b0 = __ballot_sync(0xffffffffu, kx[d][0] > thr), b1 = __ballot_sync(0xffffffffu, kx[d][1] > thr);
                const unsigned b2 = __ballot_sync(0xffffffffu, kx[d][2] > thr), b3 = __ballot_sync(0xffffffffu, kx[d][3] > thr);
                const uint32_t n0 = __popc(b0), n1 = n0 + __popc(b1), n2 = n1 + __popc(b2), total = n2 + __popc(b3);
                if (total == 0) continue;
                uint32_t wbase = 0;
                if (lane == 0) wbase = held + atomicAdd(&sCnt[period % 3], total);
                wbase = __shfl_sync(0xffffffffu, wbase, 0);
                const unsigned lt = (1u << lane) - 1u;
                auto put = [&](unsigned bv, uint32_t first, int v) {
                    if ((bv >> lane) & 1u) {
                        const uint32_t o = wbase + first + __popc(bv & lt);
                        sKey[o] = kx[d][v]; sPos[o] = p + v;
                        if (HAS_VALUE) sVal[o] = __ldg(vrow + p + v);
                    }
                };
                put(b0, 0, 0); put(b1, n0, 1); put(b2, n1, 2); put(b3, n2, 3);
            }
        };
        // two register sets: the loads of step i + 1 are in flight while step i is compared, appended and -- every fourth chunk --
        // the block meets at its barriers
        // A first threshold without a select.  The k-th largest of ANY subset is a lower bound of the tile's k-th key: take the 256
        // per-thread maxima of the first step (16 values each) and rank them by counting (two barriers).  For uniform scores the
        // 100th largest of 256 such maxima is the 97.0th percentile, against 97.6 for the exact 100th of those 4,096 values.  The
        // exact alternative -- filling the buffer with the first 4,096 values and running the radix select on them -- was HALF of
        // the kernel at the config-5 shape (tools/topk_floor.py: 3.57 ms with that one select per tile, 1.45 ms without any).
        // Re-ranking running maxima after 4 and 16 steps as well was measured and dropped: each ranking cost more than the smaller
        // select at the end of the tile saved.
        float ka[kDepth][4], kb[kDepth][4];
        fetch(ka, c0);
        exclude(ka, c0);
        if (a.k <= (uint32_t)kKThreads) {
            float mx = -INFINITY;
#pragma unroll
            for (int d = 0; d < kDepth; d++)
#pragma unroll
                for (int v = 0; v < 4; v++) mx = fmaxf(mx, ka[d][v]);
            sKey[tid] = mx;                                                    // (the candidate buffer is still empty)
            __syncthreads();
            uint32_t rank = 0;
            for (uint32_t j = 0; j < (uint32_t)kKThreads; j++) { const float o = sKey[j]; rank += (o > mx || (o == mx && j < tid)) ? 1u : 0u; }
            // candidates must be strictly above the threshold, and an element EQUAL to the bound can still belong to the top k
            if (rank == a.k - 1 && mx > -kMaxValue) sThr = fmaxf(nextafterf(mx, -INFINITY), -kMaxValue);
            __syncthreads();
        }
        for (uint32_t base0 = c0; base0 < c1; base0 += 2 * kStep) {
            if (base0 + kStep < c1) { fetch(kb, base0 + kStep); exclude(kb, base0 + kStep); }
            consume(ka, base0);
            if (base0 + kStep >= c1) break;
            if (base0 + 2 * kStep < c1) { fetch(ka, base0 + 2 * kStep); exclude(ka, base0 + 2 * kStep); }
            consume(kb, base0 + kStep);
        }
        __syncthreads();
        uint32_t n = held + sCnt[period % 3];
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
    size_t smem = (size_t)kKCap * 8 + (hasVal ? (size_t)kKCap * 4 : 0) + (kKHistWords + 2 * (size_t)k) * 4 + (hasFilter ? (size_t)(segLen / 32 + 1) * 4 : 0);
    uint64_t grid = (uint64_t)batch * segs;
    uint64_t cap = (uint64_t)ctx->numSMs * 6;
    // the tile loop is persistent: the grid must be ONE resident wave (a block that starts after the others have finished their share
    // runs its whole share alone; ncu of round 1's 6 blocks per SM: 1.5 waves) -- as many blocks per SM as this call's shared memory allows
#define DSB_TOPK(V, F)                                                                                           \
    do {                                                                                                         \
        DSB_CUDA_OK(cudaFuncSetAttribute(topk_kernel<V, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        int occ = 0;                                                                                             \
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, topk_kernel<V, F>, kKThreads, smem) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 1; } \
        cap = (uint64_t)ctx->numSMs * (uint64_t)occ;                                                             \
        if (grid > cap) grid = cap;                                                                              \
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
        smem = (size_t)kKCap * 12 + (kKHistWords + 2 * (size_t)k) * 4;
        grid = batch;
        DSB_CUDA_OK(cudaFuncSetAttribute(topk_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        {
            int occ = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, topk_kernel<true, false>, kKThreads, smem) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 1; }
            cap = (uint64_t)ctx->numSMs * (uint64_t)occ;
        }
        if (grid > cap) grid = cap;
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
