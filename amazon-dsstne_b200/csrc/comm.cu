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
#include <cstring>

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
