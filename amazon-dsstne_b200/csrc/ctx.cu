// ctx.cu -- context behind the opaque dsb200_ctx handle: device selection, stream, the explicit
// parameter block that replaces `__constant__ GpuData cData` (E/GpuTypes.h:265-311), and the
// scratch the kernels borrow (they never allocate).  Replaces, for this path only, what
// GpuContext::Startup/SetNeuralNetwork/CopyConstants do (E/GpuTypes.cpp:62-498).
#include "common.cuh"
#include "launch.h"

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>

namespace dsb {

static std::atomic<uint64_t> g_launches{0};
int g_pdl = 1;                                      // option "pdl": programmatic dependent launch of the main-stream kernels (launch.h)
void count_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int fail(dsb200_ctx* ctx, int code, const char* what)
{
    if (ctx) {
        const char* cs = (code > 0 && code < 10000) ? cudaGetErrorString((cudaError_t)code) : "dsb200 error";
        snprintf(ctx->lastError, sizeof(ctx->lastError), "%s: %s (%d)", what, cs, code);
    }
    return code ? code : DSB200_EINVAL;
}

}  // namespace dsb

extern "C" {

int dsb200_version(void) { return DSB200_VERSION; }

uint64_t dsb200_launch_count(void) { return dsb::g_launches.load(); }

void dsb200_params_default(dsb200_params* p)
{
    // E/NNNetwork.cpp:27-58, E/GpuTypes.cpp:475-498
    memset(p, 0, sizeof(*p));
    p->denoising_p = 0.0f;  p->denoising_q = 1.0f;
    p->deltaBoost_one = 1.0f; p->deltaBoost_zero = 1.0f;
    p->SMCE_oneTarget = 0.9f; p->SMCE_zeroTarget = 0.1f; p->SMCE_oneScale = 1.0f; p->SMCE_zeroScale = 1.0f;
}

int dsb200_ctx_create(dsb200_ctx** out, int device)
{
    if (!out) return DSB200_EINVAL;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) {
        fprintf(stderr, "dsstne_b200: no CUDA device %d (there is no CPU fallback)\n", device);
        return DSB200_ENOGPU;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return DSB200_ENOGPU;
    if (prop.major != 10) {
        fprintf(stderr, "dsstne_b200: device %d is sm_%d%d; this library is built for sm_100a (B200) only\n",
                device, prop.major, prop.minor);
        return DSB200_ENOGPU;
    }
    dsb200_ctx* ctx = new dsb200_ctx();
    memset(ctx->lastError, 0, sizeof(ctx->lastError));
    ctx->device = device;
    ctx->numSMs = prop.multiProcessorCount;
    dsb200_params_default(&ctx->params);
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->dAccumulator, 4 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemset(ctx->dAccumulator, 0, 4 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaHostAlloc(&ctx->hAccumulator, 4 * sizeof(unsigned long long), cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaHostAlloc(&ctx->dStatus, sizeof(uint32_t), cudaHostAllocMapped);
    if (e != cudaSuccess) { delete ctx; return (int)e; }
    *ctx->dStatus = 0;
    const char* env = getenv("DSB200_NO_TMA");
    ctx->noTma = (env && env[0] == '1') ? 1 : 0;
    *out = ctx;
    // default scratch: 4096 row counters, 32M partial floats (128 MB of the 180 GB)
    int rc = dsb200_ctx_reserve(ctx, 4096, (size_t)32 << 20);
    if (rc) { dsb200_ctx_destroy(ctx); *out = nullptr; return rc; }
    return 0;
}

int dsb200_ctx_destroy(dsb200_ctx* ctx)
{
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    dsb200_comm_destroy(ctx);
    dsb::gemm_release(ctx);
    cudaFree(ctx->dAccumulator);
    cudaFreeHost(ctx->hAccumulator);
    cudaFreeHost(ctx->dStatus);
    cudaFree(ctx->dRowCounters);
    cudaFree(ctx->dPartials);
    cudaFree(ctx->dGemmWs);
    cudaFree(ctx->dDenseWs);
    cudaFree(ctx->dHeavy);
    cudaFree(ctx->dHeavy3);
    cudaFree(ctx->dGsWs);
    cudaFree(ctx->prepBits.buf); cudaFree(ctx->prepW.buf); cudaFree(ctx->prepX.buf);
    delete ctx;
    return 0;
}

int dsb200_ctx_set_stream(dsb200_ctx* ctx, void* stream)
{
    if (!ctx) return DSB200_EINVAL;
    ctx->stream = (cudaStream_t)stream;
    return 0;
}

int dsb200_ctx_set_params(dsb200_ctx* ctx, const dsb200_params* p)
{
    if (!ctx || !p) return DSB200_EINVAL;
    ctx->params = *p;
    return 0;
}

int dsb200_ctx_reserve(dsb200_ctx* ctx, uint32_t maxBatch, size_t partialFloats)
{
    if (!ctx) return DSB200_EINVAL;
    if (maxBatch > ctx->rowCounterCap) {
        DSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->dRowCounters);
        ctx->dRowCounters = nullptr; ctx->rowCounterCap = 0;
        DSB_CUDA_OK(cudaMalloc(&ctx->dRowCounters, (size_t)maxBatch * sizeof(uint32_t)));
        DSB_CUDA_OK(cudaMemset(ctx->dRowCounters, 0, (size_t)maxBatch * sizeof(uint32_t)));
        ctx->rowCounterCap = maxBatch;
    }
    if (partialFloats > ctx->partialsCap) {
        DSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
        cudaFree(ctx->dPartials);
        ctx->dPartials = nullptr; ctx->partialsCap = 0;
        DSB_CUDA_OK(cudaMalloc(&ctx->dPartials, partialFloats * sizeof(float)));
        ctx->partialsCap = partialFloats;
    }
    return 0;
}

int dsb200_ctx_set_option(dsb200_ctx* ctx, const char* name, int value)
{
    if (!ctx || !name) return DSB200_EINVAL;
    if (!strcmp(name, "no_tma")) { ctx->noTma = value; return 0; }
    if (!strcmp(name, "transpose_sort")) { ctx->transposeSort = value; return 0; }
    if (!strcmp(name, "gemm_mode")) { ctx->gemmMode = value; return 0; }
    if (!strcmp(name, "profile")) { ctx->profile = value; return 0; }
    if (!strcmp(name, "fast_math")) { ctx->fastMath = value; return 0; }
    if (!strcmp(name, "wgrad_tile_kernel")) { ctx->wgradTileKernel = value; return 0; }
    if (!strcmp(name, "z_staged_kernel")) { ctx->zStagedKernel = value; return 0; }
    if (!strcmp(name, "output_tile_kernel")) { ctx->outputTileKernel = value; return 0; }
    if (!strcmp(name, "no_small_dense")) { ctx->noSmallDense = value; return 0; }
    if (!strcmp(name, "gemm_splits")) { ctx->gemmSplits = value; return 0; }
    if (!strcmp(name, "gemm_loader")) { ctx->gemmLoader = value; return 0; }
    if (!strcmp(name, "p2p_exchange")) { ctx->p2pExchange = value; return 0; }
    if (!strcmp(name, "wgrad_light_blocks")) { ctx->wgradLightBlocks = value; return 0; }
    if (!strcmp(name, "gemm_tc_min_work")) { ctx->gemmTcMinWork = value; return 0; }
    if (!strcmp(name, "gemm_debug")) { ctx->gemmDebug = value; return 0; }
    if (!strcmp(name, "pdl")) { dsb::g_pdl = value ? 1 : 0; return 0; }
    if (!strcmp(name, "gemm_stream")) { ctx->gemmStream = value; return 0; }
    if (!strcmp(name, "wgrad_two_kernel")) { ctx->wgradTwoKernel = value; return 0; }
    if (!strcmp(name, "wgrad_max_entries")) { ctx->wgradMaxEntries = (uint32_t)value; return 0; }
    return dsb::fail(ctx, DSB200_EINVAL, "unknown option");
}

int dsb200_ctx_sync(dsb200_ctx* ctx)
{
    if (!ctx) return DSB200_EINVAL;
    DSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    if (*ctx->dStatus) {
        uint32_t st = *ctx->dStatus;
        *ctx->dStatus = 0;
        if (st == DSB200_STATUS_Z_WORKSPACE)
            return dsb::fail(ctx, DSB200_ESTATE, "sparse_z: split-row workspace too small, call dsb200_ctx_reserve");
        if (st & DSB200_STATUS_G_CAPACITY)
            return dsb::fail(ctx, DSB200_ESTATE, "sparse_wgrad: the batch holds more entries than option wgrad_max_entries allows");
        if (st == DSB200_STATUS_T_OVERFLOW)
            return dsb::fail(ctx, DSB200_ESTATE, "sparse_transpose: a column overran its capacity slot");
        return dsb::fail(ctx, DSB200_ESTATE, "device status");
    }
    return 0;
}

const char* dsb200_last_error(dsb200_ctx* ctx) { return ctx ? ctx->lastError : "null context"; }

void dsb200_shard_range(uint32_t N, uint32_t rank, uint32_t nranks, uint32_t* pMinX, uint32_t* pMaxX)
{
    // E/NNLayer.cpp:108-112
    *pMinX = (uint32_t)(((uint64_t)N * rank) / nranks);
    *pMaxX = (uint32_t)(((uint64_t)N * (rank + 1)) / nranks);
}

int dsb200_weight_outgoing_larger(uint32_t inputStride, uint32_t outputStride)
{
    // E/NNWeight.cpp:435-457
    return (uint64_t)outputStride * 3 > (uint64_t)inputStride * 2;
}

}  // extern "C"
