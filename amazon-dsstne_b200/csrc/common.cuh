// common.cuh -- device-side helpers shared by the dsstne_b200 kernels (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dsstne_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "dsstne_b200 kernels are written for sm_100a (B200) only"
#endif

namespace dsb {

constexpr int      kWarp        = 32;
constexpr float    kErrorScaleF = 1073741824.0f;           // ESCALE = 2^30 (E/GpuTypes.h:65-69)
constexpr double   kOneOverErrorScale = 1.0 / 1073741824.0;
constexpr float    kMinError    = 1.0e-12f;                // E/NNTypes.h:46
constexpr float    kMinActivation = 0.000001f;             // E/NNTypes.h:47
constexpr float    kMaxActivation = 0.999999f;             // E/NNTypes.h:48
constexpr float    kMaxValue    = 999999999999999.0f;      // E/NNTypes.h:49

// ---------------------------------------------------------------- context
// Host-side context behind the opaque dsb200_ctx handle.
struct Workspace {
    void*  ptr   = nullptr;
    size_t bytes = 0;
};

}  // namespace dsb

struct dsb200_ctx {
    int            device      = 0;
    int            numSMs      = 148;
    cudaStream_t   stream      = nullptr;
    dsb200_params  params;                 // explicit replacement of `__constant__ GpuData cData`
    // scratch owned by the context (kernels never allocate):
    unsigned long long* dAccumulator = nullptr;   // fixed-point loss accumulator (2 slots)
    unsigned long long* hAccumulator = nullptr;   // pinned host mirror
    uint32_t*      dRowCounters = nullptr;        // self-resetting arrival counters (split rows)
    uint32_t       rowCounterCap = 0;
    float*         dPartials   = nullptr;         // split-row partial sums
    size_t         partialsCap = 0;               // in floats
    uint32_t*      dStatus     = nullptr;         // sticky device status word
    int            noTma       = 0;               // DSB200_NO_TMA=1 -> plain-load index staging
    int            transposeSort = 1;             // emit ascending rows inside each transposed column
    void*          comm        = nullptr;         // ncclComm_t when model parallel
    int            rank = 0, nranks = 1;
    void*          dDenseWs    = nullptr;         // dense_small.cu: arrival counters + segment partials of the batch-split weight gradient
    size_t         denseWsBytes = 0, denseWsTiles = 0;
    int            gemmMode    = 0;
    int            gemmDebug   = 0;               // bring-up switches of gemm_tc.cu (option "gemm_debug")
    int            noSmallDense = 0;              // option "no_small_dense": keep small dense layers on the library SGEMM
    int            gemmTcMinWork = 1024;          // option "gemm_tc_min_work": tiles x k-iterations below which the SIMT kernel runs
    int            gemmLoader  = -1;              // option "gemm_loader": -1 / 2 = A through tensor memory, 1 = register-staged loader, 0 = cp.async + split warps
    int            gemmSplits  = 0;               // option "gemm_splits": 0 = automatic split-K factor
    float*         dGemmWs     = nullptr;         // split-K partial tiles of the tcgen05 GEMM
    size_t         gemmWsCap   = 0;               // in floats
    uint32_t*      dHeavy      = nullptr;         // sparse gradient: [0] heavy-column count, [1..] heavy-column list
    size_t         heavyCap    = 0;
    void*          dHeavy3     = nullptr;         // unified sparse gradient: zeroed int64 accumulators, item list, slots, arrival counters
    size_t         heavy3Bytes = 0;
    uint32_t       heavySlots = 0, heavyN = 0, heavyM = 0;   // layout the zeroed accumulators of the unified scheme were laid out for
    int            wgradTwoKernel = 0;            // option "wgrad_two_kernel": round 1's light + heavy kernel pair instead of the unified kernel
    uint32_t       wgradMaxEntries = 1u << 20;    // option "wgrad_max_entries": capacity of the transposed matrix the gradient kernels are called with
    int            wgradTileKernel = 0;           // option "wgrad_tile_kernel": force the one-kernel tile scheme
    int            zStagedKernel = 0;             // option "z_staged_kernel": force the TMA-staged CTA kernel for sparse Z
    int            outputTileKernel = 0;          // option "output_tile_kernel": force the two-phase tile kernel in dsb200_output_pass
    int            fastMath    = 1;               // option "fast_math": MUFU exp/log/rcp in the output pass (default on)
    int            profile     = 0;               // option "profile": event pairs around every kernel entry
    int            wgradLightBlocks = 0;          // option "wgrad_light_blocks": blocks per SM of the light-column gradient kernel, 0 = by occupancy
    int            p2pExchange = 0;               // option "p2p_exchange": 1 = exchange steps as one kernel over peer memory (comm.cu, experimental)
    void*          p2p         = nullptr;         // dsb::P2PState, owned by comm.cu
    int            gemmStream  = 1;               // option "gemm_stream": output-layer shapes on the TMA / tensor-memory kernels of gemm_stream.cu
    void*          dGsWs       = nullptr;         // gemm_stream.cu: split-K partials and operand copies made inside a call (grow only)
    size_t         gsWsBytes   = 0;
    // gemm_stream.cu: operands prepared AHEAD of the call that uses them (dsb200_gemm_fwd_output_prepare / _dx_prepare on a side stream,
    // or the forward pass preparing X for the weight gradient).  One-shot: `valid` is cleared by the call that consumes the buffer.
    struct Prep { void* buf = nullptr; size_t bytes = 0; const void* key = nullptr; uint32_t a = 0, b = 0, c = 0; bool valid = false; };
    Prep           prepBits, prepW, prepX;
    char           lastError[256];
};

namespace dsb {

#define DSB_CUDA_OK(expr)                                                     \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) return dsb::fail(ctx, (int)_e, #expr);         \
    } while (0)

int fail(dsb200_ctx* ctx, int code, const char* what);

// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__

__device__ __forceinline__ uint32_t smem_u32(const void* p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

// Programmatic dependent launch (launch.h: launch_pdl): the kernels of the training step's main stream are launched with
// programmatic stream serialization, so the NEXT kernel's CTAs are dispatched -- launch latency, parameter loads, barrier / tensor
// memory set-up -- while this one still runs.  pdl_launch_dependents() early: lets that dispatch start once every CTA of this grid
// has got here.  pdl_wait() before the first access to global memory that an earlier kernel may have written, and before any write:
// returns when the preceding grids have completed and flushed.  Both are no-ops in a normally launched kernel.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// example lookup shared by every kernel on the path (shuffle, then Indexed):
//   E/kernels.cu:670 and :751
__device__ __forceinline__ uint32_t example_of(const dsb200_params& P, const uint32_t* __restrict__ exIndex,
                                               uint32_t position, uint32_t b)
{
    uint32_t pos = P.bShuffleIndices ? __ldg(P.pShuffleIndex + position + b) : position + b;
    if (exIndex) pos = __ldg(exIndex + pos);
    return pos;
}

// analog element -> float (uchar * 1/256, char * 1/128; see oracle/dsstne_oracle.c:orc_value)
__device__ __forceinline__ float load_value(const void* __restrict__ data, int dataType, uint64_t j)
{
    switch (dataType) {
    case DSB200_DT_FLOAT:  return __ldg((const float*)data + j);
    case DSB200_DT_DOUBLE: return (float)__ldg((const double*)data + j);
    case DSB200_DT_UINT:   return (float)__ldg((const uint32_t*)data + j);
    case DSB200_DT_INT:    return (float)__ldg((const int32_t*)data + j);
    case DSB200_DT_LLINT:  return (float)__ldg((const long long*)data + j);
    case DSB200_DT_ULLINT: return (float)__ldg((const unsigned long long*)data + j);
    case DSB200_DT_UCHAR:  return (float)__ldg((const unsigned char*)data + j) * (float)(1.0 / 256.0);
    case DSB200_DT_CHAR:   return (float)__ldg((const signed char*)data + j) * (float)(1.0 / 128.0);
    default:               return 0.0f;
    }
}

// 128-bit read-only gather that does not pollute L1 (weight/delta rows are
// re-used across CTAs through L2, never inside one CTA).
__device__ __forceinline__ float4 ldg_nc_f4(const float4* p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
// streaming (evict-first) 128-bit load/store for tensors touched once per pass
__device__ __forceinline__ float4 ldg_cs_f4(const float4* p)
{
    float4 r;
    asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_cs_f4(float4* p, const float4& v)
{
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// L2-coherent loads for data written by other CTAs of the same grid
__device__ __forceinline__ float4 ldg_cg_f4(const float4* p)
{
    float4 r;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}
__device__ __forceinline__ float ldg_cg_f(const float* p)
{
    float r;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(r) : "l"(p) : "memory");
    return r;
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP) ----
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "DSB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DSB_DONE;\n\t"
        "bra DSB_WAIT;\n\t"
        "DSB_DONE:\n\t"
        "}" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared, `bytes` multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(void* smemDst, const void* gmemSrc, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(smemDst)), "l"(gmemSrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- reductions ----
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// fixed-point conversion used by the reference for order-independent sums:
// llrintf(2^30 * x).  For |x| < 2 the result fits int32, so the cheap F2I.S32
// path gives the identical integer; larger magnitudes take the 64-bit path.
__device__ __forceinline__ long long fix30(float x)
{
    float y = kErrorScaleF * x;
    if (fabsf(x) < 1.984375f) return (long long)__float2int_rn(y);
    return llrintf(y);
}

__device__ __forceinline__ float sgnf(float x) { return (float)((x > 0.0f) - (x < 0.0f)); }

#endif  // __CUDACC__

}  // namespace dsb
