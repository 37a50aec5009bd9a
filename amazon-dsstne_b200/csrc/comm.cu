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
// Exchange steps as ONE kernel over peer memory (dsb200_p2p_*; the engine uses them when option "p2p_exchange" is on and the
// set-up succeeded on every rank, else the NCCL calls above).
//
// Why: on BASELINE config 3 the exchanged blocks are 512 KB and a training step has 8-11 of them; an NCCL collective plus the
// repack kernel next to it costs 20-28 us (driver-run SCALE_r01, round-2 profiles), as much as the compute between two
// exchanges.  Over NVSwitch every peer is one hop away, so a collective of this size is a few microseconds of stores plus one
// flag round trip.  Round 1's first attempt (stage locally, grid-wide arrival counter, every block polling, PULL over NVLink with
// scalar loads and a 64-bit division per element) measured 30 us and is gone.  This version PUSHES:
//   * every rank owns an ARENA of `slots` regions and an array of arrival counters, exported with cudaIpcGetMemHandle and mapped
//     by every peer; a collective names the slot it delivers into (the engine gives every layer boundary its own slots, so a
//     gathered operand stays valid until the same boundary is exchanged again in the next step);
//   * all-gather: each thread reads 16 bytes of the local slice once and stores them into the slot of EVERY rank (remote stores are
//     posted: nobody waits for a round trip); reduce-scatter: each rank stores the columns a peer owns into that peer's slot,
//     rank-major, and the owner sums the P contributions in rank order (deterministic) with bias and activation applied on the
//     way out (kAddBias + kCalculate*Activation, E/NNLayer.cpp:1257-1340, fused);
//   * completion: after its stores a block does __threadfence_system and one red.release.sys.add per peer on the peer's counter
//     [slot][source]; a consumer waits until every source has delivered `blocks` arrivals of this call (ld.acquire.sys).  In the
//     all-gather only block 0 waits (the kernel's end is the delivery point); in the reduce-scatter every block waits before it
//     sums, so the grid is at most one block per SM (co-resident: no block can wait for one that cannot run).
// Re-use of a slot: a rank can push into slot s of a peer only after it has itself finished the previous exchange on s, which
// needs every peer's contribution to THAT exchange, which a peer launches only after its consumers of the exchange before it
// have been queued ahead in stream order -- and between two uses of a slot (one training step apart) every rank has taken
// part in at least one other exchange.
struct P2PState {
    int       P = 0, rank = 0;
    uint32_t  slots = 0;
    size_t    slotFloats = 0;
    float*    arena = nullptr;
    unsigned long long* flags = nullptr;            // [slots][P]: arrivals from source rank r on slot s
    float**   dPeerArena = nullptr;                 // device tables [P] (own pointers at [rank])
    unsigned long long** dPeerFlags = nullptr;
    std::vector<void*> opened;
    std::vector<unsigned long long> expected;       // per slot: arrivals per source a finished call has seen
    bool      failed = false;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

struct P2PArgs {
    float* const* peerArena; unsigned long long* const* peerFlags; unsigned long long* myFlags;
    uint32_t rank, P, slot; size_t slotOff; unsigned long long target;
    uint32_t batch, stride;
};

// this block's stores are out -> one arrival on every rank's counter [slot][me]
__device__ __forceinline__ void p2p_signal(const P2PArgs& a)
{
    // the block barrier orders every thread's stores before the signalling threads; their system-scope release is cumulative over
    // them (one fence per signalling thread instead of one per storing thread)
    __syncthreads();
    if (threadIdx.x < a.P) {
        unsigned long long* f = a.peerFlags[threadIdx.x] + (size_t)a.slot * a.P + a.rank;
        __threadfence_system();
        asm volatile("red.release.sys.global.add.u64 [%0], %1;" :: "l"(f), "l"(1ull) : "memory");
    }
}
// every source has delivered all its blocks of this call
__device__ __forceinline__ void p2p_wait(const P2PArgs& a)
{
    if (threadIdx.x < a.P) {
        const unsigned long long* f = a.myFlags + (size_t)a.slot * a.P + threadIdx.x;
        while (ld_acquire_sys(f) < a.target) { }
    }
    __syncthreads();
}

__device__ __forceinline__ void p2p_range(uint32_t stride, uint32_t r, uint32_t P, uint32_t& lo, uint32_t& hi)
{
    lo = (uint32_t)((uint64_t)stride * r / P); hi = (uint32_t)((uint64_t)stride * (r + 1) / P);      // E/NNLayer.cpp:108-112
}

// local [batch][span_me] -> columns [lo_me, hi_me) of the full [batch][stride] in the slot of every rank
template <int VEC>
__global__ void __launch_bounds__(256)
p2p_all_gather_kernel(const P2PArgs a, const float* __restrict__ pLocal)
{
    uint32_t lo, hi;
    p2p_range(a.stride, a.rank, a.P, lo, hi);
    const uint32_t span = hi - lo, spanV = span / VEC;
    const uint32_t total = a.batch * spanV;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t b = i / spanV, c = (i - b * spanV) * VEC;
        const size_t dst = a.slotOff + (size_t)b * a.stride + lo + c;
        if (VEC == 4) {
            const float4 v = *reinterpret_cast<const float4*>(pLocal + (size_t)b * span + c);
            for (uint32_t q = 0; q < a.P; q++) *reinterpret_cast<float4*>(a.peerArena[(a.rank + q) % a.P] + dst) = v;
        } else {
            const float v = pLocal[(size_t)b * span + c];
            for (uint32_t q = 0; q < a.P; q++) a.peerArena[(a.rank + q) % a.P][dst] = v;
        }
    }
    p2p_signal(a);
    if (blockIdx.x == 0) p2p_wait(a);
}

__device__ __forceinline__ float p2p_act(int act, float z, float slope, float alpha, float lambda)
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

// [batch][stride] partial sums of every rank -> this rank's columns [lo, hi) summed over the ranks (+ bias, activation)
struct P2PEpi { const float* bias; int act; float slope, alpha, lambda; };
template <int VEC>
__global__ void __launch_bounds__(256)
p2p_reduce_scatter_kernel(const P2PArgs a, const float* __restrict__ pIn, float* __restrict__ pOut, const P2PEpi e)
{
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    // push: the columns of peer q go to slot region [me][batch][span_q] of q
    for (uint32_t k = 1; k < a.P; k++) {
        const uint32_t q = (a.rank + k) % a.P;
        uint32_t qlo, qhi;
        p2p_range(a.stride, q, a.P, qlo, qhi);
        const uint32_t span = qhi - qlo, spanV = span / VEC, total = a.batch * spanV;
        float* dst = a.peerArena[q] + a.slotOff + (size_t)a.rank * a.batch * span;
        for (uint32_t i = tid; i < total; i += nth) {
            const uint32_t b = i / spanV, c = (i - b * spanV) * VEC;
            if (VEC == 4) *reinterpret_cast<float4*>(dst + (size_t)b * span + c) = *reinterpret_cast<const float4*>(pIn + (size_t)b * a.stride + qlo + c);
            else          dst[(size_t)b * span + c] = pIn[(size_t)b * a.stride + qlo + c];
        }
    }
    p2p_signal(a);
    p2p_wait(a);
    uint32_t lo, hi;
    p2p_range(a.stride, a.rank, a.P, lo, hi);
    const uint32_t span = hi - lo, spanV = span / VEC, total = a.batch * spanV;
    const float* mine = a.peerArena[a.rank] + a.slotOff;
    for (uint32_t i = tid; i < total; i += nth) {
        const uint32_t b = i / spanV, c = (i - b * spanV) * VEC;
        if (VEC == 4) {
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (uint32_t r = 0; r < a.P; r++) {                                   // rank order: the same bits on every run and every rank count
                const float4 v = (r == a.rank) ? *reinterpret_cast<const float4*>(pIn + (size_t)b * a.stride + lo + c)
                                               : __ldcg(reinterpret_cast<const float4*>(mine + ((size_t)r * a.batch + b) * span + c));
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
            if (e.bias) { const float4 bb = *reinterpret_cast<const float4*>(e.bias + c); acc.x += bb.x; acc.y += bb.y; acc.z += bb.z; acc.w += bb.w; }
            if (e.act != DSB200_ACT_LINEAR) {
                acc.x = p2p_act(e.act, acc.x, e.slope, e.alpha, e.lambda); acc.y = p2p_act(e.act, acc.y, e.slope, e.alpha, e.lambda);
                acc.z = p2p_act(e.act, acc.z, e.slope, e.alpha, e.lambda); acc.w = p2p_act(e.act, acc.w, e.slope, e.alpha, e.lambda);
            }
            *reinterpret_cast<float4*>(pOut + (size_t)b * span + c) = acc;
        } else {
            float acc = 0.f;
            for (uint32_t r = 0; r < a.P; r++)
                acc += (r == a.rank) ? pIn[(size_t)b * a.stride + lo + c] : __ldcg(mine + ((size_t)r * a.batch + b) * span + c);
            if (e.bias) acc += e.bias[c];
            pOut[(size_t)b * span + c] = p2p_act(e.act, acc, e.slope, e.alpha, e.lambda);
        }
    }
}

static void p2p_release(dsb200_ctx* ctx)
{
    P2PState* st = static_cast<P2PState*>(ctx->p2p);
    if (!st) return;
    for (void* p : st->opened) cudaIpcCloseMemHandle(p);
    cudaFree(st->arena); cudaFree(st->flags); cudaFree(st->dPeerArena); cudaFree(st->dPeerFlags);
    delete st;
    ctx->p2p = nullptr;
}

// Collective (every rank calls it with the same arguments): allocate the arena, exchange the IPC handles over NCCL, map the
// peers, and VOTE -- a failure on any rank makes every rank report DSB200_EUNSUPPORTED (the caller stays on NCCL).
static int p2p_setup(dsb200_ctx* ctx, uint32_t slots, size_t slotFloats)
{
    p2p_release(ctx);
    P2PState* st = new P2PState;
    ctx->p2p = st;
    st->P = ctx->nranks; st->rank = ctx->rank; st->slots = slots;
    st->slotFloats = (slotFloats + 63) & ~(size_t)63;
    st->expected.assign(slots, 0ull);
    const int P = st->P;
    struct Handles { cudaIpcMemHandle_t x, f; };
    bool ok = P >= 2 && P <= 32 && slots > 0;
    ok = ok && cudaMalloc(&st->arena, (size_t)slots * st->slotFloats * sizeof(float)) == cudaSuccess;
    ok = ok && cudaMalloc(&st->flags, (size_t)slots * P * sizeof(unsigned long long)) == cudaSuccess;
    ok = ok && cudaMemsetAsync(st->flags, 0, (size_t)slots * P * sizeof(unsigned long long), ctx->stream) == cudaSuccess;
    Handles mine;
    memset(&mine, 0, sizeof(mine));
    ok = ok && cudaIpcGetMemHandle(&mine.x, st->arena) == cudaSuccess && cudaIpcGetMemHandle(&mine.f, st->flags) == cudaSuccess;
    // the handle exchange is itself a collective, so it runs even on a rank whose allocation failed (that rank votes "no" below)
    std::vector<Handles> all((size_t)std::max(P, 1));
    unsigned char* dH = nullptr;
    bool exchanged = cudaMalloc(&dH, (size_t)std::max(P, 1) * sizeof(Handles)) == cudaSuccess;
    if (exchanged) {
        exchanged = cudaMemcpyAsync(dH + (size_t)st->rank * sizeof(Handles), &mine, sizeof(Handles), cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess &&
                    g_nccl.AllGather(dH + (size_t)st->rank * sizeof(Handles), dH, sizeof(Handles), ncclUint8, (ncclComm_t)ctx->comm, ctx->stream) == ncclSuccess &&
                    cudaMemcpyAsync(all.data(), dH, (size_t)P * sizeof(Handles), cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess &&
                    cudaStreamSynchronize(ctx->stream) == cudaSuccess;
    }
    ok = ok && exchanged;
    std::vector<float*> px((size_t)std::max(P, 1), nullptr);
    std::vector<unsigned long long*> pf((size_t)std::max(P, 1), nullptr);
    for (int r = 0; ok && r < P; r++) {
        if (r == st->rank) { px[r] = st->arena; pf[r] = st->flags; continue; }
        void* x = nullptr; void* f = nullptr;
        ok = cudaIpcOpenMemHandle(&x, all[r].x, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess;
        if (ok) { st->opened.push_back(x); ok = cudaIpcOpenMemHandle(&f, all[r].f, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess; }
        if (ok) st->opened.push_back(f);
        px[r] = static_cast<float*>(x); pf[r] = static_cast<unsigned long long*>(f);
    }
    ok = ok && cudaMalloc(&st->dPeerArena, (size_t)P * sizeof(float*)) == cudaSuccess && cudaMalloc(&st->dPeerFlags, (size_t)P * sizeof(unsigned long long*)) == cudaSuccess;
    ok = ok && cudaMemcpyAsync(st->dPeerArena, px.data(), (size_t)P * sizeof(float*), cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess &&
         cudaMemcpyAsync(st->dPeerFlags, pf.data(), (size_t)P * sizeof(unsigned long long*), cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess;
    int coop = 0;
    ok = ok && cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device) == cudaSuccess && coop;
    // all ranks must agree: the vote is a sum of failures; it also orders every rank's flag memset before anybody's first push
    float* dVote = reinterpret_cast<float*>(dH);
    float vote = ok ? 0.0f : 1.0f;
    bool voted = dH && cudaMemcpyAsync(dVote, &vote, sizeof(float), cudaMemcpyHostToDevice, ctx->stream) == cudaSuccess &&
                 g_nccl.AllReduce(dVote, dVote, 1, ncclFloat32, ncclSum, (ncclComm_t)ctx->comm, ctx->stream) == ncclSuccess &&
                 cudaMemcpyAsync(&vote, dVote, sizeof(float), cudaMemcpyDeviceToHost, ctx->stream) == cudaSuccess &&
                 cudaStreamSynchronize(ctx->stream) == cudaSuccess;
    cudaFree(dH);
    cudaGetLastError();                                                            // a failed probe must not poison later calls
    st->failed = !(voted && vote == 0.0f);
    if (st->failed) { p2p_release(ctx); return DSB200_EUNSUPPORTED; }
    return 0;
}

static int p2p_args(dsb200_ctx* ctx, uint32_t slot, uint32_t batch, uint32_t stride, uint32_t blocks, P2PArgs* out)
{
    P2PState* st = static_cast<P2PState*>(ctx->p2p);
    if (!st || st->failed) return fail(ctx, DSB200_ESTATE, "p2p: dsb200_p2p_setup has not succeeded");
    if (slot >= st->slots) return fail(ctx, DSB200_EINVAL, "p2p: slot out of range");
    if ((size_t)batch * (stride + (uint32_t)st->P) > st->slotFloats) return fail(ctx, DSB200_EINVAL, "p2p: message larger than a slot");
    st->expected[slot] += blocks;
    P2PArgs a;
    a.peerArena = st->dPeerArena; a.peerFlags = st->dPeerFlags; a.myFlags = st->flags;
    a.rank = (uint32_t)st->rank; a.P = (uint32_t)st->P; a.slot = slot; a.slotOff = (size_t)slot * st->slotFloats; a.target = st->expected[slot];
    a.batch = batch; a.stride = stride;
    *out = a;
    return 0;
}

// same block count on every rank (it is part of the arrival arithmetic): from the largest slice
static uint32_t p2p_blocks(dsb200_ctx* ctx, uint64_t elems)
{
    uint64_t g = (elems / 4 + 511) / 512;
    if (g > (uint64_t)ctx->numSMs) g = (uint64_t)ctx->numSMs;
    return (uint32_t)std::max<uint64_t>(g, 1);
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

int dsb200_p2p_setup(dsb200_ctx* ctx, uint32_t slots, size_t slotFloats)
{
    using namespace dsb;
    if (!ctx) return DSB200_EINVAL;
    if (ctx->nranks < 2 || !ctx->comm) return fail(ctx, DSB200_EUNSUPPORTED, "p2p_setup: needs an initialised communicator with at least two ranks");
    const int rc = p2p_setup(ctx, slots, slotFloats);
    if (rc) return fail(ctx, rc, "p2p_setup: peer mapping unavailable on at least one rank (the exchange steps stay on NCCL)");
    return 0;
}

float* dsb200_p2p_slot(dsb200_ctx* ctx, uint32_t slot)
{
    using namespace dsb;
    P2PState* st = ctx ? static_cast<P2PState*>(ctx->p2p) : nullptr;
    if (!st || st->failed || slot >= st->slots) return nullptr;
    return st->arena + (size_t)slot * st->slotFloats;
}

int dsb200_p2p_all_gather(dsb200_ctx* ctx, uint32_t slot, uint32_t batch, uint32_t stride, const float* pLocal)
{
    DSB_PROFILE_T(ctx, "p2p_all_gather", stride);
    using namespace dsb;
    if (!ctx || !pLocal) return fail(ctx, DSB200_EINVAL, "p2p_all_gather: null argument");
    if (!batch || !stride) return 0;
    const uint32_t P = (uint32_t)ctx->nranks;
    const uint32_t spanMax = (stride + P - 1) / P;
    const uint32_t blocks = p2p_blocks(ctx, (uint64_t)batch * spanMax);
    P2PArgs a;
    int rc = p2p_args(ctx, slot, batch, stride, blocks, &a);
    if (rc) return rc;
    uint32_t lo, hi; dsb200_shard_range(stride, (uint32_t)ctx->rank, P, &lo, &hi);
    const P2PState* st = static_cast<P2PState*>(ctx->p2p);
    // every rank must take the same path only for the block count; the vector width is a local matter
    const bool vec = ((hi - lo) % 4 == 0) && (lo % 4 == 0) && (stride % 4 == 0) && (((uintptr_t)pLocal) % 16 == 0) && (st->slotFloats % 4 == 0);
    if (vec) p2p_all_gather_kernel<4><<<blocks, 256, 0, ctx->stream>>>(a, pLocal);
    else     p2p_all_gather_kernel<1><<<blocks, 256, 0, ctx->stream>>>(a, pLocal);
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
    return 0;
}

int dsb200_p2p_reduce_scatter(dsb200_ctx* ctx, uint32_t slot, uint32_t batch, uint32_t stride, const float* pIn, float* pOut, const float* pBias,
                              int activation, float slope, float alpha, float lambda)
{
    DSB_PROFILE_T(ctx, "p2p_reduce_scatter", stride);
    using namespace dsb;
    if (!ctx || !pIn || !pOut) return fail(ctx, DSB200_EINVAL, "p2p_reduce_scatter: null argument");
    if (activation == DSB200_ACT_SOFTMAX) return fail(ctx, DSB200_EUNSUPPORTED, "p2p_reduce_scatter: softmax is a row operation; pass Linear and call dsb200_activation");
    if (!batch || !stride) return 0;
    const uint32_t P = (uint32_t)ctx->nranks;
    const uint32_t blocks = p2p_blocks(ctx, (uint64_t)batch * stride);
    P2PArgs a;
    int rc = p2p_args(ctx, slot, batch, stride, blocks, &a);
    if (rc) return rc;
    // 16-byte path: every rank's column range must start and end on a multiple of four for every peer (the pushes address the
    // peers' ranges), i.e. stride divisible by 4 P
    const bool vec = (stride % (4 * P) == 0) && (((uintptr_t)pIn) % 16 == 0) && (((uintptr_t)pOut) % 16 == 0) && (!pBias || ((uintptr_t)pBias) % 16 == 0);
    P2PEpi e; e.bias = pBias; e.act = activation; e.slope = slope; e.alpha = alpha; e.lambda = lambda;
    // every block waits for the peers before it sums, so all blocks must be able to run at once: at most one 256-thread block per
    // SM without shared memory always can (a concurrent kernel of this library can delay them, but it never waits for this one)
    if (vec) p2p_reduce_scatter_kernel<4><<<blocks, 256, 0, ctx->stream>>>(a, pIn, pOut, e);
    else     p2p_reduce_scatter_kernel<1><<<blocks, 256, 0, ctx->stream>>>(a, pIn, pOut, e);
    count_launch();
    DSB_CUDA_OK(cudaGetLastError());
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
