// sparse_transpose.cu -- batch CSR -> CSC build for the sparse backward pass (hot-path row a5).
//
// Replaces kCalculate[Indexed]SparseTransposed[Analog][Denoised]Matrix (E/kernels.cu:1980-2534)
// plus the End<-Start device copy the dataset issues before it (E/NNTypes.h:576).
//
// Reference: one warp per example walks its row 32 nnz at a time; a 9,254-nnz row is 290
// dependent (load -> atomic -> store) rounds on one warp, and the order inside a column is
// whatever the atomics produced.  Here:
//   1. `transpose_init`  End[c] = Start[c]                       (fused memcpy, 8 B per column)
//   2. `transpose_scatter` work item = 128 consecutive nnz of one row (item list derived per
//      CTA from a shared-memory scan of the row lengths, as in sparse_z.cu); a warp owns an
//      item, every lane issues 4 independent index loads -> 4 atomics -> 4(+4) stores;
//   3. `transpose_sort`  (option "transpose_sort", default on) puts every column segment into
//      ascending batch-row order: warp bitonic network for <= 32 entries, shared-memory bitonic
//      for <= 4096.  The weight-gradient kernel does not need the order (it sums in fixed
//      point) -- this pass only makes the integer output canonical / reproducible.
#include "common.cuh"
#include "launch.h"

namespace dsb {

constexpr int kTThreads  = 256;
constexpr int kTMaxRows  = 4096;
constexpr int kTItem     = 128;           // nnz per warp item
constexpr int kSortTile  = 128;           // columns per CTA in the sort pass
constexpr int kSortMax   = 4096;          // largest column segment sorted in shared memory

__global__ void __launch_bounds__(256)
transpose_init_kernel(uint32_t N, const uint32_t* __restrict__ tStart, uint32_t* __restrict__ tEnd)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) tEnd[i] = __ldg(tStart + i);
}

struct TArgs {
    dsb200_params P;
    dsb200_sparse S;
    uint32_t position, batch, rowBase;
    int denoised;
    uint32_t* tEnd; uint32_t* tIndex; float* tData;
};

template <bool ANALOG, bool DENOISED>
__global__ void __launch_bounds__(kTThreads, 4)
transpose_scatter_kernel(const TArgs a)
{
    __shared__ uint32_t prefix[kTMaxRows + 1];
    __shared__ uint32_t scan[8];
    const int tid = threadIdx.x;
    const uint32_t batch = a.batch;

    for (uint32_t r = tid; r < batch; r += kTThreads) {
        const uint32_t ex = example_of(a.P, a.S.index, a.position, a.rowBase + r);
        const uint64_t len = __ldg(a.S.sparseEnd + ex) - __ldg(a.S.sparseStart + ex);
        prefix[r] = (uint32_t)((len + kTItem - 1) / kTItem);
    }
    __syncthreads();
    const uint32_t per = (batch + kTThreads - 1) / kTThreads;
    const uint32_t lo = min((uint32_t)tid * per, batch), hi = min(lo + per, batch);
    uint32_t local = 0;
    for (uint32_t r = lo; r < hi; r++) local += prefix[r];
    uint32_t incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += n;
    }
    if ((tid & 31) == 31) scan[tid >> 5] = incl;
    __syncthreads();
    uint32_t run = incl - local;
    for (int w = 0; w < (tid >> 5); w++) run += scan[w];
    __syncthreads();
    for (uint32_t r = lo; r < hi; r++) { const uint32_t c = prefix[r]; prefix[r] = run; run += c; }
    if (tid == kTThreads - 1) prefix[batch] = run;
    __syncthreads();
    const uint32_t T = prefix[batch];

    const uint32_t lane = tid & 31;
    const uint32_t warpsPerGrid = gridDim.x * (kTThreads / 32);
    for (uint32_t t = blockIdx.x * (kTThreads / 32) + (tid >> 5); t < T; t += warpsPerGrid) {
        uint32_t l = 0, h = batch;
        while (h - l > 1) { const uint32_t mid = (l + h) >> 1; if (prefix[mid] <= t) l = mid; else h = mid; }
        const uint32_t row = l, k = t - prefix[row];
        const uint32_t ex = example_of(a.P, a.S.index, a.position, a.rowBase + row);
        const uint64_t rs = __ldg(a.S.sparseStart + ex), re = __ldg(a.S.sparseEnd + ex);
        const uint64_t e0 = rs + (uint64_t)k * kTItem;
        const uint64_t e1 = (re < e0 + kTItem) ? re : e0 + kTItem;
        float w = a.S.dataWeight ? __ldg(a.S.dataWeight + ex) : 1.0f;
        // only the weighted Boolean denoised kernel folds q into the payload (E/kernels.cu:2152)
        if (DENOISED && !ANALOG && a.S.dataWeight) w *= a.P.denoising_q;
        const uint32_t b = a.rowBase + row;
        uint32_t col[4]; float val[4]; bool keep[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const uint64_t e = e0 + lane + 32 * u;
            keep[u] = e < e1;
            col[u] = 0; val[u] = w;
            if (keep[u]) {
                col[u] = __ldg(a.S.sparseIndex + e);
                if (DENOISED && __ldg(a.S.denoisingRandom + e) < a.P.denoising_p) keep[u] = false;
                if (ANALOG) val[u] = w * load_value(a.S.sparseData, a.S.dataType, e);
            }
        }
        uint32_t pos[4];
#pragma unroll
        for (int u = 0; u < 4; u++) pos[u] = keep[u] ? atomicAdd(a.tEnd + col[u], 1u) : 0u;
#pragma unroll
        for (int u = 0; u < 4; u++) {
            if (keep[u]) {
                a.tIndex[pos[u]] = b;
                if (a.tData) a.tData[pos[u]] = val[u];
            }
        }
    }
}

// ---- canonical order pass -------------------------------------------------------------
__device__ __forceinline__ void cmpswap(uint32_t& ka, float& va, uint32_t& kb, float& vb, bool up)
{
    if ((ka > kb) == up) { const uint32_t tk = ka; ka = kb; kb = tk; const float tv = va; va = vb; vb = tv; }
}

__global__ void __launch_bounds__(kTThreads)
transpose_sort_kernel(uint32_t N, const uint32_t* __restrict__ tStart, const uint32_t* __restrict__ tEnd,
                      uint32_t* __restrict__ tIndex, float* __restrict__ tData)
{
    __shared__ uint32_t sKey[kSortMax];
    __shared__ float    sVal[kSortMax];
    __shared__ uint32_t heavy[kSortTile];
    __shared__ uint32_t nHeavy;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    for (uint32_t tile = blockIdx.x; tile * kSortTile < N; tile += gridDim.x) {
        if (tid == 0) nHeavy = 0;
        __syncthreads();
        // each warp looks at kSortTile/8 = 16 columns of the tile... one lane per column, two rounds
        for (uint32_t cbase = tile * kSortTile + warp * 32; cbase < min((tile + 1) * kSortTile, N); cbase += 8 * 32) {
            const uint32_t c = cbase + lane;
            uint32_t s = 0, n = 0;
            if (c < N && c < (tile + 1) * kSortTile) { s = __ldg(tStart + c); n = tEnd[c] - s; }
            if (n > 32 && n <= (uint32_t)kSortMax) heavy[atomicAdd(&nHeavy, 1u)] = c;
            uint32_t light = __ballot_sync(0xffffffffu, n >= 2 && n <= 32);
            while (light) {
                const int src = __ffs(light) - 1;
                light &= light - 1;
                const uint32_t cs = __shfl_sync(0xffffffffu, s, src), cn = __shfl_sync(0xffffffffu, n, src);
                uint32_t key = (lane < cn) ? tIndex[cs + lane] : 0xffffffffu;
                float val = (tData && lane < cn) ? tData[cs + lane] : 0.0f;
                // 32-wide bitonic network on (key,val), ascending
#pragma unroll
                for (int ksz = 2; ksz <= 32; ksz <<= 1) {
#pragma unroll
                    for (int j = ksz >> 1; j > 0; j >>= 1) {
                        const uint32_t ok = __shfl_xor_sync(0xffffffffu, key, j);
                        const float ov = __shfl_xor_sync(0xffffffffu, val, j);
                        const bool up = ((lane & ksz) == 0);
                        const bool lower = ((lane & j) == 0);
                        const bool takeMin = (lower == up);
                        const bool swap = takeMin ? (ok < key) : (ok > key);
                        if (swap) { key = ok; val = ov; }
                    }
                }
                if (lane < cn) { tIndex[cs + lane] = key; if (tData) tData[cs + lane] = val; }
            }
        }
        __syncthreads();
        // heavy columns of this tile: shared-memory bitonic sort by the whole CTA
        const uint32_t nh = nHeavy;
        for (uint32_t hI = 0; hI < nh; hI++) {
            const uint32_t c = heavy[hI];
            const uint32_t s = __ldg(tStart + c), n = tEnd[c] - s;
            uint32_t p2 = 64; while (p2 < n) p2 <<= 1;
            for (uint32_t i = tid; i < p2; i += kTThreads) {
                sKey[i] = (i < n) ? tIndex[s + i] : 0xffffffffu;
                sVal[i] = (tData && i < n) ? tData[s + i] : 0.0f;
            }
            __syncthreads();
            for (uint32_t ksz = 2; ksz <= p2; ksz <<= 1) {
                for (uint32_t j = ksz >> 1; j > 0; j >>= 1) {
                    for (uint32_t i = tid; i < p2; i += kTThreads) {
                        const uint32_t ixj = i ^ j;
                        if (ixj > i) cmpswap(sKey[i], sVal[i], sKey[ixj], sVal[ixj], (i & ksz) == 0);
                    }
                    __syncthreads();
                }
            }
            for (uint32_t i = tid; i < n; i += kTThreads) { tIndex[s + i] = sKey[i]; if (tData) tData[s + i] = sVal[i]; }
            __syncthreads();
        }
        __syncthreads();
    }
}

template <bool ANALOG, bool DENOISED>
static int launch_scatter(dsb200_ctx* ctx, const TArgs& a)
{
    int grid = ctx->numSMs * 2;
    transpose_scatter_kernel<ANALOG, DENOISED><<<grid, kTThreads, 0, ctx->stream>>>(a);
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    return 0;
}


// ---------------------------------------------------------------- capacity table on the device (a4)
// NNDataSet<T>::GenerateSparseTransposedMatrix (E/NNTypes.cpp:1631-1735) sizes every transposed column on the host from
// per-column datapoint counts of the whole dataset.  For a dataset that is replaced every step (a serving / streaming
// caller loading one batch at a time) that host pass is the dominant cost, so the same table is built here: count the
// entries of every column over the first `rows` examples, then tStart = exclusive prefix of the counts rounded up to a
// multiple of 32 (the reference's alignment).  Exact per-column counts are an upper bound for any batch window.
__global__ void __launch_bounds__(256)
column_count_kernel(const uint64_t* __restrict__ start, const uint64_t* __restrict__ end, const uint32_t* __restrict__ index, uint32_t rows,
                    uint32_t N, uint32_t* __restrict__ count, volatile uint32_t* status)
{
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t r = gw; r < rows; r += nw) {
        const uint64_t s = __ldg(start + r), e = __ldg(end + r);
        for (uint64_t j = s + lane; j < e; j += 32) {
            const uint32_t c = __ldg(index + j);
            if (c < N) atomicAdd(count + c, 1u);
            else *status = DSB200_STATUS_T_OVERFLOW;
        }
    }
}

// single CTA: exclusive scan of align32(count) -> tStart; total capacity -> *pTotal (may be NULL).  Tiles of 1024 x 8 columns:
// every thread takes eight consecutive counts (two 16-byte loads when the table is aligned), so a warp reads 1 KB in one go and
// the eight loads of a thread do not wait for each other; the running total is carried from tile to tile.
constexpr int kScanItems = 8;

__global__ void __launch_bounds__(1024)
capacity_scan_kernel(const uint32_t* __restrict__ count, uint32_t N, uint32_t* __restrict__ tStart, uint32_t* __restrict__ pTotal)
{
    __shared__ uint32_t sWarp[32];
    __shared__ uint32_t sTile;
    const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool vec = ((reinterpret_cast<uintptr_t>(count) | reinterpret_cast<uintptr_t>(tStart)) & 15) == 0;
    uint32_t carry = 0;
    for (uint64_t base = 0; base < N; base += 1024ull * kScanItems) {
        const uint64_t first = base + (uint64_t)tid * kScanItems;
        uint32_t v[kScanItems];
        const bool whole = first + kScanItems <= N;
        if (whole && vec) {
            const uint4 a = __ldg(reinterpret_cast<const uint4*>(count + first)), b = __ldg(reinterpret_cast<const uint4*>(count + first) + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else {
#pragma unroll
            for (int k = 0; k < kScanItems; k++) v[k] = (first + k < N) ? __ldg(count + first + k) : 0u;
        }
        uint32_t local = 0;
#pragma unroll
        for (int k = 0; k < kScanItems; k++) { v[k] = (v[k] + 31u) & ~31u; local += v[k]; }
        uint32_t incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= (uint32_t)o) incl += n;
        }
        if (lane == 31) sWarp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const uint32_t w = sWarp[lane];
            uint32_t wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t n = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= (uint32_t)o) wi += n;
            }
            sWarp[lane] = wi - w;                                         // exclusive warp offsets
            if (lane == 31) sTile = wi;
        }
        __syncthreads();
        uint32_t run = carry + sWarp[warp] + incl - local;
        carry += sTile;
        if (whole && vec) {
            uint4 a, b;
            a.x = run; run += v[0]; a.y = run; run += v[1]; a.z = run; run += v[2]; a.w = run; run += v[3];
            b.x = run; run += v[4]; b.y = run; run += v[5]; b.z = run; run += v[6]; b.w = run;
            reinterpret_cast<uint4*>(tStart + first)[0] = a;
            reinterpret_cast<uint4*>(tStart + first)[1] = b;
        } else {
#pragma unroll
            for (int k = 0; k < kScanItems; k++) { if (first + k < N) tStart[first + k] = run; run += v[k]; }
        }
        __syncthreads();                                                  // sWarp / sTile are rewritten by the next tile
    }
    if (tid == 0 && pTotal) *pTotal = carry;
}

}  // namespace dsb

extern "C" int dsb200_sparse_transpose(dsb200_ctx* ctx, const dsb200_sparse* s, uint32_t position, uint32_t batch, int denoised,
                                       uint32_t N, const uint32_t* tStart, uint32_t* tEnd, uint32_t* tIndex, float* tData)
{
    DSB_PROFILE(ctx, "sparse_transpose");
    using namespace dsb;
    if (!ctx || !s || !tEnd || !tIndex) return fail(ctx, DSB200_EINVAL, "sparse_transpose: null argument");
    if (!s->sparseStart || !s->sparseEnd || !s->sparseIndex) return fail(ctx, DSB200_EINVAL, "sparse_transpose: CSR arrays missing");
    if (denoised && !s->denoisingRandom) return fail(ctx, DSB200_EINVAL, "sparse_transpose: denoised without random buffer");
    const bool analog = s->sparseData != nullptr;
    if ((analog || s->dataWeight) && !tData) return fail(ctx, DSB200_EINVAL, "sparse_transpose: transposed data buffer missing");
    if (tStart) {
        if (!N) return fail(ctx, DSB200_EINVAL, "sparse_transpose: N == 0 with a start table");
        int grid = (int)((N + 255) / 256); if (grid > ctx->numSMs * 8) grid = ctx->numSMs * 8;
        transpose_init_kernel<<<grid, 256, 0, ctx->stream>>>(N, tStart, tEnd);
        count_launch();
        DSB_CUDA_OK(cudaGetLastError());
    }
    for (uint32_t base = 0; base < batch; base += kTMaxRows) {
        TArgs a;
        a.P = ctx->params; a.S = *s; a.position = position; a.rowBase = base;
        a.batch = (batch - base < (uint32_t)kTMaxRows) ? batch - base : (uint32_t)kTMaxRows;
        a.denoised = denoised; a.tEnd = tEnd; a.tIndex = tIndex; a.tData = tData;
        int rc;
        if (analog) rc = denoised ? launch_scatter<true, true>(ctx, a) : launch_scatter<true, false>(ctx, a);
        else        rc = denoised ? launch_scatter<false, true>(ctx, a) : launch_scatter<false, false>(ctx, a);
        if (rc) return rc;
    }
    if (ctx->transposeSort && tStart && N) {
        int grid = (int)((N + kSortTile - 1) / kSortTile); if (grid > ctx->numSMs * 4) grid = ctx->numSMs * 4;
        transpose_sort_kernel<<<grid, kTThreads, 0, ctx->stream>>>(N, tStart, tEnd, tIndex, tData);
        count_launch();
        DSB_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

extern "C" int dsb200_transposed_capacity(dsb200_ctx* ctx, const dsb200_sparse* s, uint32_t rows, uint32_t N, uint32_t* pCountScratch,
                                          uint32_t* pTransposedStart, uint32_t* pDevTotal)
{
    DSB_PROFILE(ctx, "transposed_capacity");
    using namespace dsb;
    if (!ctx || !s || !pCountScratch || !pTransposedStart || !N) return fail(ctx, DSB200_EINVAL, "transposed_capacity: null argument");
    if (!s->sparseStart || !s->sparseEnd || !s->sparseIndex) return fail(ctx, DSB200_EINVAL, "transposed_capacity: CSR arrays missing");
    if (s->index) return fail(ctx, DSB200_EUNSUPPORTED, "transposed_capacity: indexed datasets size their table on the host");
    DSB_CUDA_OK(cudaMemsetAsync(pCountScratch, 0, (size_t)N * sizeof(uint32_t), ctx->stream));
    if (rows) {
        int grid = (int)((rows + 7) / 8);
        if (grid > ctx->numSMs * 8) grid = ctx->numSMs * 8;
        column_count_kernel<<<grid, 256, 0, ctx->stream>>>(s->sparseStart, s->sparseEnd, s->sparseIndex, rows, N, pCountScratch, ctx->dStatus);
        count_launch();
        DSB_CUDA_OK(cudaGetLastError());
    }
    capacity_scan_kernel<<<1, 1024, 0, ctx->stream>>>(pCountScratch, N, pTransposedStart, pDevTotal);
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    return 0;
}
