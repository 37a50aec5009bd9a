// comm.cu -- model-parallel exchange steps (hot-path row a15) as NCCL collectives over NVLink.
//
// Replaces NNLayer::Reduce / NNLayer::Gather (E/NNLayer.cpp:2702-2826: P-1 ring stages of
// kCopy2D/kAddBuffers2D, each followed by cudaDeviceSynchronize + MPI_Barrier) and
// NNNetwork::P2P_Allreduce (E/NNNetwork.cpp:4127-4197).  One stream-ordered NCCL call per
// collective, no host synchronisation.  Units of a layer are split [S*r/P, S*(r+1)/P)
// (E/NNLayer.cpp:108-112); NCCL wants contiguous equal chunks, so the [batch][S] row-major
// activations are (un)packed to rank-major [P][batch][S/P] by a small 128-bit copy kernel on
// either side of the collective (uneven splits fall back to all-reduce / per-rank broadcast).
//
// NCCL is resolved with dlopen at first use: a single-GPU process never needs it, and inside a
// torch process the already-loaded bundled libnccl.so.2 is the one that gets used.
#include "common.cuh"
#include "launch.h"

#include <dlfcn.h>
#include <algorithm>
#include <cstring>
#include <vector>

namespace dsb {

// minimal NCCL ABI (stable since 2.x); values from nccl.h
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclInt8 = 0, ncclUint8 = 1, ncclInt32 = 2, ncclUint32 = 3, ncclInt64 = 4, ncclUint64 = 5, ncclFloat16 = 6, ncclFloat32 = 7, ncclFloat64 = 8 };
enum { ncclSum = 0 };

struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(ncclUniqueId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*ReduceScatter)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static bool nccl_load()
{
    if (g_nccl.lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
    if (!lib) return false;
#define DSB_SYM(field, name) *(void**)(&g_nccl.field) = dlsym(lib, name); if (!g_nccl.field) return false
    DSB_SYM(GetUniqueId, "ncclGetUniqueId");
    DSB_SYM(CommInitRank, "ncclCommInitRank");
    DSB_SYM(CommDestroy, "ncclCommDestroy");
    DSB_SYM(AllReduce, "ncclAllReduce");
    DSB_SYM(ReduceScatter, "ncclReduceScatter");
    DSB_SYM(AllGather, "ncclAllGather");
    DSB_SYM(Broadcast, "ncclBroadcast");
    DSB_SYM(GroupStart, "ncclGroupStart");
    DSB_SYM(GroupEnd, "ncclGroupEnd");
    DSB_SYM(GetErrorString, "ncclGetErrorString");
#undef DSB_SYM
    g_nccl.lib = lib;
    return true;
}

#define DSB_NCCL_OK(expr)                                                        \
    do {                                                                         \
        int _r = (expr);                                                         \
        if (_r != ncclSuccess) return fail(ctx, DSB200_ENCCL, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : #expr); \
    } while (0)

// [batch][stride] row-major  <->  rank-major [P][batch][span] (span = stride / P), 128-bit when possible
template <bool PACK>
__global__ void __launch_bounds__(256)
repack_kernel(const float* __restrict__ src, float* __restrict__ dst, uint32_t batch, uint32_t stride, uint32_t P)
{
    const uint32_t span = stride / P;
    const uint64_t total = (uint64_t)batch * stride;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t b = (uint32_t)(i / stride), c = (uint32_t)(i % stride);
        const uint32_t r = c / span, cc = c % span;
        const uint64_t j = ((uint64_t)r * batch + b) * span + cc;
        if (PACK) dst[j] = src[i]; else dst[i] = src[j];
    }
}

__global__ void __launch_bounds__(256)
slice_kernel(const float* __restrict__ full, float* __restrict__ out, uint32_t batch, uint32_t stride, uint32_t lo, uint32_t span)
{
    const uint64_t total = (uint64_t)batch * span;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t b = (uint32_t)(i / span), c = (uint32_t)(i % span);
        out[i] = full[(size_t)b * stride + lo + c];
    }
}

__global__ void __launch_bounds__(256)
place_kernel(const float* __restrict__ local, float* __restrict__ full, uint32_t batch, uint32_t stride, uint32_t lo, uint32_t span)
{
    const uint64_t total = (uint64_t)batch * span;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t b = (uint32_t)(i / span), c = (uint32_t)(i % span);
        full[(size_t)b * stride + lo + c] = local[i];
    }
}

static unsigned grid_of(dsb200_ctx* ctx, uint64_t n)
{
    uint64_t g = (n + 255) / 256; const uint64_t cap = (uint64_t)ctx->numSMs * 8;
    if (g > cap) g = cap; if (g < 1) g = 1;
    return (unsigned)g;
}

// =====================================================================================================================
// Exchange steps as ONE kernel over peer memory (option "p2p_exchange" = 1; EXPERIMENTAL: written at the end of round 1
// after the GPU budget was spent, not yet run -- the NCCL path above stays the default).
//
// Why: on BASELINE config 2 the exchanged blocks are 512 KB and a training step has 11 of them; an NCCL collective plus the
// repack kernel next to it costs 15-20 us, as much as the compute between two exchanges.  Over
// NVSwitch every peer is one hop away, so a collective of this size is a few microseconds of loads / stores plus one flag
// round trip.  Every rank owns an exchange buffer (two regions, used alternately) and a flag array, both exported with
// cudaIpcGetMemHandle and mapped by every peer.  A collective is one kernel per rank:
//   stage    copy my contribution into my own exchange region (local stores)
//   publish  last block to finish: st.release.sys of the collective's epoch into slot `rank` of EVERY rank's flag array
//   wait     thread 0 of every block polls its OWN flag array until all P slots carry the epoch (ld.acquire.sys)
//   gather   all-gather: read every peer's region over NVLink straight into the caller's [batch][stride] rows
//            reduce-scatter: sum the P regions' columns [lo, hi) in rank order (deterministic) into the caller's slice
// No repack pass (the slices are addressed in place, uneven unit ranges included) and no second barrier: a rank can reuse
// a region only two collectives later, and it cannot get there before every peer has published the collective in between,
// which a peer does only after it has finished reading the older one.
struct P2PState {
    int       P = 0, rank = 0;
    float*    xbuf = nullptr;                       // own exchange buffer: 2 regions of `cap` floats
    size_t    cap = 0;
    unsigned long long* flags = nullptr;            // own flag array: P epochs + the block-arrival counter
    float**   dPeerX = nullptr;                     // device tables [P] of the mapped peer buffers (own pointers at [rank])
    unsigned long long** dPeerFlags = nullptr;
    std::vector<void*> opened;                      // cudaIpcOpenMemHandle mappings
    unsigned long long epoch = 0;
    bool      failed = false;                       // set-up failed on some rank: every rank stays on NCCL
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

struct P2PArgs {
    float* const* peerX; unsigned long long* const* peerFlags; unsigned long long* myFlags;
    uint32_t rank, P; unsigned long long epoch; size_t region;        // region = offset of this epoch's region in floats
    uint32_t batch, stride;
};

// every block: my stores are done -> (last block) publish the epoch to all ranks -> wait until all ranks have published
__device__ __forceinline__ void p2p_publish_and_wait(const P2PArgs& a)
{
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long* counter = a.myFlags + a.P;
        const unsigned long long prev = atomicAdd(counter, 1ull);
        if (prev == (unsigned long long)gridDim.x - 1) {
            *counter = 0;                                                          // for the next collective on this stream
            __threadfence_system();
            for (uint32_t r = 0; r < a.P; r++) st_release_sys(a.peerFlags[r] + a.rank, a.epoch);
        }
        for (uint32_t r = 0; r < a.P; r++)
            while (ld_acquire_sys(a.myFlags + r) < a.epoch) __nanosleep(32);
    }
    __syncthreads();
}

__device__ __forceinline__ void p2p_range(uint32_t stride, uint32_t r, uint32_t P, uint32_t& lo, uint32_t& hi)
{
    lo = (uint32_t)((uint64_t)stride * r / P); hi = (uint32_t)((uint64_t)stride * (r + 1) / P);      // E/NNLayer.cpp:108-112
}

// local [batch][span_rank] of every rank -> full [batch][stride] on every rank
__global__ void __launch_bounds__(256)
p2p_all_gather_kernel(const P2PArgs a, const float* __restrict__ pLocal, float* __restrict__ pFull)
{
    uint32_t lo, hi;
    p2p_range(a.stride, a.rank, a.P, lo, hi);
    const uint64_t mine = (uint64_t)a.batch * (hi - lo), tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (uint64_t)gridDim.x * blockDim.x;
    float* region = a.peerX[a.rank] + a.region;
    for (uint64_t i = tid; i < mine; i += nth) region[i] = pLocal[i];
    p2p_publish_and_wait(a);
    for (uint32_t q = 0; q < a.P; q++) {
        const uint32_t r = (a.rank + q) % a.P;                                     // start with the local slice, spread the peers
        uint32_t rlo, rhi;
        p2p_range(a.stride, r, a.P, rlo, rhi);
        const uint32_t span = rhi - rlo;
        if (!span) continue;
        const float* src = (r == a.rank) ? pLocal : a.peerX[r] + a.region;
        const uint64_t n = (uint64_t)a.batch * span;
        for (uint64_t i = tid; i < n; i += nth) {
            const uint32_t b = (uint32_t)(i / span), c = (uint32_t)(i % span);
            pFull[(size_t)b * a.stride + rlo + c] = (r == a.rank) ? src[i] : __ldcg(src + i);
        }
    }
}

// [batch][stride] of every rank summed; this rank keeps columns [lo, hi) as [batch][span]
__global__ void __launch_bounds__(256)
p2p_reduce_scatter_kernel(const P2PArgs a, const float* __restrict__ pIn, float* __restrict__ pOut)
{
    const uint64_t total = (uint64_t)a.batch * a.stride, tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (uint64_t)gridDim.x * blockDim.x;
    float* region = a.peerX[a.rank] + a.region;
    for (uint64_t i = tid; i < total; i += nth) region[i] = pIn[i];
    p2p_publish_and_wait(a);
    uint32_t lo, hi;
    p2p_range(a.stride, a.rank, a.P, lo, hi);
    const uint32_t span = hi - lo;
    const uint64_t n = (uint64_t)a.batch * span;
    for (uint64_t i = tid; i < n; i += nth) {
        const uint32_t b = (uint32_t)(i / span), c = (uint32_t)(i % span);
        const size_t off = a.region + (size_t)b * a.stride + lo + c;
        float acc = 0.0f;
        for (uint32_t r = 0; r < a.P; r++) acc += (r == a.rank) ? pIn[(size_t)b * a.stride + lo + c] : __ldcg(a.peerX[r] + off);   // rank order: the same bits on every run
        pOut[i] = acc;
    }
}

static void p2p_release(dsb200_ctx* ctx)
{
    P2PState* st = static_cast<P2PState*>(ctx->p2p);
    if (!st) return;
    for (void* p : st->opened) cudaIpcCloseMemHandle(p);
    cudaFree(st->xbuf); cudaFree(st->flags); cudaFree(st->dPeerX); cudaFree(st->dPeerFlags);
    delete st;
    ctx->p2p = nullptr;
}

// Collective (every rank calls it at the same point with the same `need`): true when the peer-memory path can take a message
// of `need` floats.  The first call maps the peers; a later, larger message than the mapped regions stays on NCCL.
static bool p2p_ready(dsb200_ctx* ctx, size_t need)
{
    P2PState* st = static_cast<P2PState*>(ctx->p2p);
    if (st) return !st->failed && st->cap >= need;
    st = new P2PState;
    ctx->p2p = st;
    st->P = ctx->nranks; st->rank = ctx->rank;
    const int P = st->P;
    struct Handles { cudaIpcMemHandle_t x, f; };
    bool ok = true;
    st->cap = std::max(need, (size_t)8 << 20);                                     // 2 x 32 MB by default
    ok = ok && cudaMalloc(&st->xbuf, 2 * st->cap * sizeof(float)) == cudaSuccess;
    ok = ok && cudaMalloc(&st->flags, (size_t)(P + 1) * sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaMemsetAsync(st->flags, 0, (size_t)(P + 1) * sizeof(unsigned long long), ctx->stream) == cudaSuccess;
    Handles mine;
    memset(&mine, 0, sizeof(mine));
    ok = ok && cudaIpcGetMemHandle(&mine.x, st->xbuf) == cudaSuccess && cudaIpcGetMemHandle(&mine.f, st->flags) == cudaSuccess;
    // the handle exchange is itself a collective, so it runs even on a rank whose allocation failed (that rank votes "no" below)
    std::vector<Handles> all((size_t)P);
    unsigned char* dH = nullptr;
    bool exchanged = cudaMalloc(&dH, (size_t)P * sizeof(Handles)) == cudaSuccess;
    if (exchanged) {
        exchanged = cudaMemcpyAsync(dH + (size_t)st->rank * sizeof(Handles), &mine, sizeof(Handles), cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess &&
                    g_nccl.AllGather(dH + (size_t)st->rank * sizeof(Handles), dH, sizeof(Handles), ncclUint8, (ncclComm_t)ctx->comm, ctx->stream) == ncclSuccess &&
                    cudaMemcpyAsync(all.data(), dH, (size_t)P * sizeof(Handles), cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess &&
                    cudaStreamSynchronize(ctx->stream) == cudaSuccess;
    }
    ok = ok && exchanged;
    std::vector<float*> px((size_t)P, nullptr);
    std::vector<unsigned long long*> pf((size_t)P, nullptr);
    for (int r = 0; ok && r < P; r++) {
        if (r == st->rank) { px[r] = st->xbuf; pf[r] = st->flags; continue; }
        void* x = nullptr; void* f = nullptr;
        ok = cudaIpcOpenMemHandle(&x, all[r].x, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
        if (ok) { st->opened.push_back(x); ok = cudaIpcOpenMemHandle(&f, all[r].f, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess; }
        if (ok) st->opened.push_back(f);
        px[r] = static_cast<float*>(x); pf[r] = static_cast<unsigned long long*>(f);
    }
    ok = ok && cudaMalloc(&st->dPeerX, (size_t)P * sizeof(float*)) == cudaSuccess && cudaMalloc(&st->dPeerFlags, (size_t)P * sizeof(unsigned long long*)) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(st->dPeerX, px.data(), (size_t)P * sizeof(float*), cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess &&
         cudaMemcpyAsync(st->dPeerFlags, pf.data(), (size_t)P * sizeof(unsigned long long*), cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess;
    // all ranks must agree: the vote is a sum of failures
    float* dVote = reinterpret_cast<float*>(dH);
    float vote = ok ? 0.0f : 1.0f;
    bool voted = dH && cudaMemcpyAsync(dVote, &vote, sizeof(float), cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess &&
                 g_nccl.AllReduce(dVote, dVote, 1, ncclFloat32, ncclSum, (ncclComm_t)ctx->comm, ctx->stream) == ncclSuccess &&
                 cudaMemcpyAsync(&vote, dVote, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess &&
                 cudaStreamSynchronize(ctx->stream) == cudaSuccess;
    cudaFree(dH);
    cudaGetLastError();                                                            // a failed probe must not poison later calls
    st->failed = !(voted && vote == 0.0f);
    return !st->failed && st->cap >= need;
}

static P2PArgs p2p_args(dsb200_ctx* ctx, uint32_t batch, uint32_t stride)
{
    P2PState* st = static_cast<P2PState*>(ctx->p2p);
    st->epoch++;
    P2PArgs a;
    a.peerX = st->dPeerX; a.peerFlags = st->dPeerFlags; a.myFlags = st->flags;
    a.rank = (uint32_t)st->rank; a.P = (uint32_t)st->P; a.epoch = st->epoch; a.region = (st->epoch & 1ull) ? st->cap : 0;
    a.batch = batch; a.stride = stride;
    return a;
}

static unsigned p2p_grid(dsb200_ctx* ctx, uint64_t n)
{
    // every block spins in p2p_publish_and_wait, so the whole grid must be resident: at most one block per SM
    uint64_t g = (n + 1023) / 1024;
    if (g > (uint64_t)ctx->numSMs) g = (uint64_t)ctx->numSMs;
    return (unsigned)std::max<uint64_t>(g, 1);
}

}  // namespace dsb

extern "C" {

int dsb200_comm_unique_id(void* uniqueId128)
{
    using namespace dsb;
    if (!uniqueId128) return DSB200_EINVAL;
    if (!nccl_load()) return DSB200_ENCCL;
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) return DSB200_ENCCL;
    memcpy(uniqueId128, &id, sizeof(id));
    return 0;
}

int dsb200_comm_init(dsb200_ctx* ctx, const void* uniqueId128, int rank, int nranks)
{
    using namespace dsb;
    if (!ctx || !uniqueId128 || rank < 0 || nranks < 1 || rank >= nranks) return fail(ctx, DSB200_EINVAL, "comm_init: bad argument");
    if (!nccl_load()) return fail(ctx, DSB200_ENCCL, "comm_init: libnccl.so.2 not found");
    DSB_CUDA_OK(cudaSetDevice(ctx->device));
    ncclUniqueId id; memcpy(&id, uniqueId128, sizeof(id));
    ncclComm_t comm = nullptr;
    DSB_NCCL_OK(g_nccl.CommInitRank(&comm, nranks, id, rank));
    ctx->comm = comm; ctx->rank = rank; ctx->nranks = nranks;
    return 0;
}

int dsb200_comm_destroy(dsb200_ctx* ctx)
{
    using namespace dsb;
    if (ctx) p2p_release(ctx);
    if (ctx && ctx->comm && g_nccl.CommDestroy) { g_nccl.CommDestroy((ncclComm_t)ctx->comm); ctx->comm = nullptr; }
    if (ctx) { ctx->rank = 0; ctx->nranks = 1; }
    return 0;
}

int dsb200_reduce_scatter(dsb200_ctx* ctx, uint32_t batch, uint32_t stride, const float* pIn, float* pOut)
{
    DSB_PROFILE(ctx, "reduce_scatter");
    using namespace dsb;
    if (!ctx || !pIn || !pOut) return fail(ctx, DSB200_EINVAL, "reduce_scatter: null argument");
    const uint32_t P = (uint32_t)ctx->nranks;
    if (P == 1) {
        if (pIn != pOut) DSB_CUDA_OK(cudaMemcpyAsync(pOut, pIn, (size_t)batch * stride * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
        return 0;
    }
    if (!ctx->comm) return fail(ctx, DSB200_ESTATE, "reduce_scatter: communicator not initialised");
    const uint64_t total = (uint64_t)batch * stride;
    if (ctx->p2pExchange && total && p2p_ready(ctx, total)) {
        const P2PArgs a = p2p_args(ctx, batch, stride);
        p2p_reduce_scatter_kernel<<<p2p_grid(ctx, total), 256, 0, ctx->stream>>>(a, pIn, pOut);
        count_launch();
        DSB_CUDA_OK(cudaGetLastError());
        return 0;
    }
    int rc = dsb200_ctx_reserve(ctx, 0, total);
    if (rc) return rc;
    float* tmp = ctx->dPartials;
    if (stride % P == 0) {
        repack_kernel<true><<<grid_of(ctx, total), 256, 0, ctx->stream>>>(pIn, tmp, batch, stride, P);
        count_launch();
        DSB_CUDA_OK(cudaGetLastError());
        DSB_NCCL_OK(g_nccl.ReduceScatter(tmp, pOut, total / P, ncclFloat32, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
    } else {
        uint32_t lo, hi; dsb200_shard_range(stride, (uint32_t)ctx->rank, P, &lo, &hi);
        DSB_NCCL_OK(g_nccl.AllReduce(pIn, tmp, total, ncclFloat32, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
        slice_kernel<<<grid_of(ctx, (uint64_t)batch * (hi - lo)), 256, 0, ctx->stream>>>(tmp, pOut, batch, stride, lo, hi - lo);
        count_launch();
        DSB_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

int dsb200_all_gather(dsb200_ctx* ctx, uint32_t batch, uint32_t stride, const float* pLocal, float* pFull)
{
    DSB_PROFILE(ctx, "all_gather");
    using namespace dsb;
    if (!ctx || !pLocal || !pFull) return fail(ctx, DSB200_EINVAL, "all_gather: null argument");
    const uint32_t P = (uint32_t)ctx->nranks;
    if (P == 1) {
        if (pLocal != pFull) DSB_CUDA_OK(cudaMemcpyAsync(pFull, pLocal, (size_t)batch * stride * sizeof(float), cudaMemcpyDeviceToDevice, ctx->stream));
        return 0;
    }
    if (!ctx->comm) return fail(ctx, DSB200_ESTATE, "all_gather: communicator not initialised");
    const uint64_t total = (uint64_t)batch * stride;
    if (ctx->p2pExchange && total && p2p_ready(ctx, total)) {
        const P2PArgs a = p2p_args(ctx, batch, stride);
        p2p_all_gather_kernel<<<p2p_grid(ctx, total), 256, 0, ctx->stream>>>(a, pLocal, pFull);
        count_launch();
        DSB_CUDA_OK(cudaGetLastError());
        return 0;
    }
    int rc = dsb200_ctx_reserve(ctx, 0, total);
    if (rc) return rc;
    float* tmp = ctx->dPartials;
    if (stride % P == 0) {
        DSB_NCCL_OK(g_nccl.AllGather(pLocal, tmp, total / P, ncclFloat32, (ncclComm_t)ctx->comm, ctx->stream));
        repack_kernel<false><<<grid_of(ctx, total), 256, 0, ctx->stream>>>(tmp, pFull, batch, stride, P);
        count_launch();
        DSB_CUDA_OK(cudaGetLastError());
    } else {
        // uneven unit split: one broadcast per owner inside a group, then place the slices
        uint64_t off = 0;
        DSB_NCCL_OK(g_nccl.GroupStart());
        for (uint32_t r = 0; r < P; r++) {
            uint32_t lo, hi; dsb200_shard_range(stride, r, P, &lo, &hi);
            const uint64_t cnt = (uint64_t)batch * (hi - lo);
            DSB_NCCL_OK(g_nccl.Broadcast(pLocal, tmp + off, cnt, ncclFloat32, (int)r, (ncclComm_t)ctx->comm, ctx->stream));
            off += cnt;
        }
        DSB_NCCL_OK(g_nccl.GroupEnd());
        off = 0;
        for (uint32_t r = 0; r < P; r++) {
            uint32_t lo, hi; dsb200_shard_range(stride, r, P, &lo, &hi);
            const uint64_t cnt = (uint64_t)batch * (hi - lo);
            if (cnt) {
                place_kernel<<<grid_of(ctx, cnt), 256, 0, ctx->stream>>>(tmp + off, pFull, batch, stride, lo, hi - lo);
                count_launch();
            }
            off += cnt;
        }
        DSB_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

int dsb200_all_reduce(dsb200_ctx* ctx, float* pBuffer, uint64_t size)
{
    DSB_PROFILE(ctx, "all_reduce");
    using namespace dsb;
    if (!ctx || !pBuffer) return fail(ctx, DSB200_EINVAL, "all_reduce: null argument");
    if (ctx->nranks == 1 || !size) return 0;
    if (!ctx->comm) return fail(ctx, DSB200_ESTATE, "all_reduce: communicator not initialised");
    DSB_NCCL_OK(g_nccl.AllReduce(pBuffer, pBuffer, size, ncclFloat32, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
    return 0;
}

int dsb200_all_reduce_u64(dsb200_ctx* ctx, unsigned long long* pBuffer, uint64_t size)
{
    DSB_PROFILE(ctx, "all_reduce_u64");
    using namespace dsb;
    if (!ctx || !pBuffer) return fail(ctx, DSB200_EINVAL, "all_reduce_u64: null argument");
    if (ctx->nranks == 1 || !size) return 0;
    if (!ctx->comm) return fail(ctx, DSB200_ESTATE, "all_reduce_u64: communicator not initialised");
    // fixed-point loss words are two's-complement int64: integer sum is exact and order independent
    DSB_NCCL_OK(g_nccl.AllReduce(pBuffer, pBuffer, size, ncclInt64, ncclSum, (ncclComm_t)ctx->comm, ctx->stream));
    return 0;
}

}  // extern "C"
