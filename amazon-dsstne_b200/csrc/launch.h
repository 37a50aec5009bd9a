// launch.h -- host-side helpers shared by the launchers.
#pragma once
#include <stdint.h>

#define DSB200_STATUS_Z_WORKSPACE 1u   // sparse-Z split-row workspace too small (dsb200_ctx_reserve)
#define DSB200_STATUS_T_OVERFLOW  2u   // transposed column overran its capacity slot
#define DSB200_STATUS_G_CAPACITY  4u   // sparse gradient: more heavy-column work items than option "wgrad_max_entries" allows

namespace dsb {
void count_launch(uint64_t n = 1);
}

struct dsb200_ctx;
namespace dsb {
void gemm_release(dsb200_ctx* ctx);
}

// ---- optional per-family device timing (option "profile"): CUDA events on the launching stream
// around every C-ABI kernel entry; bench.py reads the totals with dsb200_profile_report ----
#include <cuda_runtime.h>
namespace dsb {
struct ProfileScope {
    dsb200_ctx* ctx;
    int slot;
    ProfileScope(dsb200_ctx* c, const char* name, unsigned long long tag = 0);
    ~ProfileScope();
};
}
#define DSB_PROFILE(ctx, name) dsb::ProfileScope _dsb_prof_scope((ctx), (name))
// the same with a size tag: reported as "name@tag", so that call sites of one family with different operand sizes (the
// output layer's 27,278 biases and a hidden layer's 128) are timed apart
#define DSB_PROFILE_T(ctx, name, tag) dsb::ProfileScope _dsb_prof_scope((ctx), (name), (unsigned long long)(tag))

// ---- programmatic dependent launch (see pdl_wait in common.cuh).  Every kernel launched through this helper executes
// griddepcontrol.wait before it touches global memory, which keeps the chain transitive: a kernel that has passed its wait knows that
// every earlier kernel of the stream has completed.  Option "pdl" = 0 falls back to plain launches.
#ifdef __CUDACC__
#include <utility>
namespace dsb {
extern int g_pdl;
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = g_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}
}
#endif
