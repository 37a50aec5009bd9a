// gemm.cu -- dense hidden / output GEMMs of the fully-connected path (hot-path row a11).
//
// Replaces the cublasSgemm calls of NNLayer::ForwardPropagateFullyConnected /
// BackPropagateFullyConnected (E/NNLayer.cpp:1073, 2223, 2274), row-major operands:
//   fwd: C[B][n]  = beta*C  + A[B][k] * W[k][n]
//   dw : G[k][n]  = beta*G  + alpha * A[B][k]^T * D[B][n]
//   dx : Dp[B][k] = beta*Dp + D[B][n] * W[k][n]^T
// Modes (option "gemm_mode"):
//   DSB200_GEMM_FP32    exact fp32 FMA arithmetic on the SIMT kernel of dense_small.cu (what the reference's SGEMM computes; parity baseline)
//   DSB200_GEMM_TF32    hand-written tcgen05 kernels (gemm_tc.cu, gemm_stream.cu), one tf32 MMA per k-step (~1e-3 relative)
//   DSB200_GEMM_TF32X3  the same kernels with the 3xTF32 split (fp32-grade, bound in tests/test_gpu_gemm.py)
// No library GEMM: shapes too small for the tensor-core kernels run on the SIMT kernel in every mode.
#include "common.cuh"
#include "launch.h"

namespace dsb {

// gemm_stream.cu: output-layer shapes (one dimension = the hidden width) on the TMA + tensor-memory kernels
bool gemm_stream_available();
int gemm_stream_debug_counters(unsigned long long* out, size_t count);
int gemm_stream_dw(dsb200_ctx* ctx, uint32_t B, uint32_t k, uint32_t n, float alpha, const float* X, const float* D, uint32_t ldd, float beta,
                   float* G, uint32_t ldg);
int gemm_stream_dx(dsb200_ctx* ctx, uint32_t B, uint32_t k, uint32_t n, const float* D, uint32_t ldd, const float* W, uint32_t ldw, float beta,
                   float* Dp, uint32_t ldp, const float* hadUnit, int hadAct, float hadScale, float slope, float ealpha, float lambda);
int gemm_stream_prepare_targets(dsb200_ctx* ctx, const dsb200_sparse* s, uint32_t position, uint32_t batch, uint32_t n);
int gemm_stream_prepare_dx(dsb200_ctx* ctx, uint32_t k, uint32_t n, const float* W, uint32_t ldw);
int dense_small_dx_hadamard(dsb200_ctx* ctx, uint32_t B, uint32_t k, uint32_t n, const float* D, const float* W, int activation, float scale,
                            const float* pUnit, float* Dp, float slope, float alpha, float lambda);
int gemm_stream_out_fwd(dsb200_ctx* ctx, const dsb200_sparse* s, int ef, uint32_t position, uint32_t batch, uint32_t k, uint32_t n, const float* X,
                        const float* W, uint32_t ldw, const float* bias, float* unitOut, float* delta, uint32_t ldd, unsigned long long* acc,
                        float* pColPartials, uint32_t* pNumPartials);
int gemm_tc_launch(dsb200_ctx* ctx, const float* A, int aMN, uint32_t lda, const float* B, int bMN, uint32_t ldb, float* C, uint32_t ldc,
                   uint32_t M, uint32_t N, uint32_t K, float alpha, float beta, const float* bias, int act, float slope, float ealpha,
                   float lambda);

// The persistent tcgen05 kernel pays ~15 us of fixed latency (launch, TMEM allocation, pipeline fill, epilogue); a GEMM
// that covers only a handful of 128 x 128 tiles (the 128 x 128 hidden weights of BASELINE config 2: 8 tiles) is
// latency bound and stays on the plain library SGEMM, which is also the exact-fp32 path.  Option "gemm_tc_min_tiles".
int dense_small_fwd(dsb200_ctx* ctx, uint32_t B, uint32_t k, uint32_t n, const float* A, const float* W, const float* bias, int act, float* C,
                    float slope, float alpha, float lambda);
// exact-fp32 SIMT GEMM (dense_small.cu): form 0 A*B, 1 A*B^T, 2 A^T*B
int dense_gemm(dsb200_ctx* ctx, int form, uint32_t M, uint32_t N, uint32_t K, float alpha, const float* A, uint32_t lda, const float* B, uint32_t ldb,
               float beta, float* C, uint32_t ldc);

static inline bool use_tc(const dsb200_ctx* ctx, uint64_t M, uint64_t N, uint64_t K)
{
    if (ctx->gemmMode != DSB200_GEMM_TF32 && ctx->gemmMode != DSB200_GEMM_TF32X3) return false;
    const uint64_t tiles = ((M + 127) / 128) * ((N + 127) / 128);
    return tiles * ((K + 15) / 16) >= (uint64_t)ctx->gemmTcMinWork;
}

// weight gradient / input delta of a layer whose narrow side (k, the hidden width) fits one or two 128-column tiles while the
// other side (n) is the long one: the streamed kernels read the big delta operand exactly once
static inline bool use_stream(const dsb200_ctx* ctx, uint64_t B, uint64_t k, uint64_t n)
{
    return ctx->gemmStream && k <= 256 && n >= 512 && B >= 1 && gemm_stream_available();
}

void gemm_release(dsb200_ctx*) {}

}  // namespace dsb

extern "C" {

int dsb200_gemm_fwd(dsb200_ctx* ctx, uint32_t B, uint32_t k, uint32_t n, const float* A, const float* W, float beta, float* C)
{
    using namespace dsb;
    if (!ctx || !A || !W || !C) return fail(ctx, DSB200_EINVAL, "gemm_fwd: null argument");
    if (!B || !k || !n) return 0;
    DSB_PROFILE(ctx, use_tc(ctx, B, n, k) ? "gemm_fwd_tc" : "gemm_fwd");
    if (use_tc(ctx, B, n, k)) return gemm_tc_launch(ctx, A, 0, k, W, 1, n, C, n, B, n, k, 1.0f, beta, nullptr, DSB200_ACT_LINEAR, 0.f, 0.f, 0.f);
    return dense_gemm(ctx, 0, B, n, k, 1.0f, A, k, W, n, beta, C, n);                                  // E/NNLayer.cpp:1072-1086
}

int dsb200_gemm_dw(dsb200_ctx* ctx, uint32_t B, uint32_t k, uint32_t n, float alpha, const float* A, const float* D, float beta, float* G)
{
    using namespace dsb;
    if (!ctx || !A || !D || !G) return fail(ctx, DSB200_EINVAL, "gemm_dw: null argument");
    if (!B || !k || !n) return 0;
    if (use_tc(ctx, k, n, B) && use_stream(ctx, B, k, n)) {
        DSB_PROFILE(ctx, "gemm_dw_stream");
        return gemm_stream_dw(ctx, B, k, n, alpha, A, D, n, beta, G, n);
    }
    DSB_PROFILE(ctx, use_tc(ctx, k, n, B) ? "gemm_dw_tc" : "gemm_dw");
    if (use_tc(ctx, k, n, B)) return gemm_tc_launch(ctx, A, 1, k, D, 1, n, G, n, k, n, B, alpha, beta, nullptr, DSB200_ACT_LINEAR, 0.f, 0.f, 0.f);
    return dense_gemm(ctx, 2, k, n, B, alpha, A, k, D, n, beta, G, n);                                 // E/NNLayer.cpp:2223-2236
}

int dsb200_gemm_dx(dsb200_ctx* ctx, uint32_t B, uint32_t k, uint32_t n, const float* D, const float* W, float beta, float* Dp)
{
    using namespace dsb;
    if (!ctx || !D || !W || !Dp) return fail(ctx, DSB200_EINVAL, "gemm_dx: null argument");
    if (!B || !k || !n) return 0;
    if (use_tc(ctx, B, k, n) && use_stream(ctx, B, k, n)) {
        DSB_PROFILE(ctx, "gemm_dx_stream");
        return gemm_stream_dx(ctx, B, k, n, D, n, W, n, beta, Dp, k, nullptr, DSB200_ACT_LINEAR, 1.0f, 0.f, 0.f, 0.f);
    }
    DSB_PROFILE(ctx, use_tc(ctx, B, k, n) ? "gemm_dx_tc" : "gemm_dx");
    if (use_tc(ctx, B, k, n)) return gemm_tc_launch(ctx, D, 0, n, W, 0, n, Dp, k, B, k, n, 1.0f, beta, nullptr, DSB200_ACT_LINEAR, 0.f, 0.f, 0.f);
    return dense_gemm(ctx, 1, B, k, n, 1.0f, D, n, W, n, beta, Dp, k);                                 // E/NNLayer.cpp:2274-2287
}

/* fused input delta: Dp = (D * W^T) (.) f'(pUnit) * scale -- cublasSgemm (E/NNLayer.cpp:2274) + kCalculateHadamardProduct of the layer
 * below (E/NNLayer.cpp:2137).  Output-layer shapes: the streamed tcgen05 kernel with the product in its split-K reduction; small
 * layers: one SIMT launch (dense_small.cu); anything else: the two calls. */
int dsb200_gemm_dx_hadamard(dsb200_ctx* ctx, uint32_t B, uint32_t k, uint32_t n, const float* D, const float* W, int activation, float scale,
                            const float* pUnit, float* Dp, float slope, float alpha, float lambda)
{
    using namespace dsb;
    if (!ctx || !D || !W || !pUnit || !Dp) return fail(ctx, DSB200_EINVAL, "gemm_dx_hadamard: null argument");
    if (!B || !k || !n) return 0;
    if (use_tc(ctx, B, k, n) && use_stream(ctx, B, k, n)) {
        DSB_PROFILE(ctx, "gemm_dx_stream");
        return gemm_stream_dx(ctx, B, k, n, D, n, W, n, 0.0f, Dp, k, pUnit, activation, scale, slope, alpha, lambda);
    }
    if ((uint64_t)B * k * n <= (1ull << 27)) {
        DSB_PROFILE(ctx, "gemm_dx_hadamard");
        return dense_small_dx_hadamard(ctx, B, k, n, D, W, activation, scale, pUnit, Dp, slope, alpha, lambda);
    }
    int rc = dsb200_gemm_dx(ctx, B, k, n, D, W, 0.0f, Dp);
    if (!rc) rc = dsb200_hadamard(ctx, activation, (uint64_t)B * k, scale, pUnit, Dp, slope, alpha, lambda);
    return rc;
}

/* fused forward of a dense layer: C = act(A * W + bias), i.e. kClearUnit + cublasSgemm(beta = 1) + kCalculate*Activation
 * (E/NNLayer.cpp:1009, 1073, 1157) in one kernel in the tensor-core modes; three calls in DSB200_GEMM_FP32. */
int dsb200_gemm_fwd_bias_act(dsb200_ctx* ctx, uint32_t B, uint32_t k, uint32_t n, const float* A, const float* W, const float* pBias,
                             int activation, float* C, float slope, float alpha, float lambda)
{
    using namespace dsb;
    if (!ctx || !A || !W || !C || !pBias) return fail(ctx, DSB200_EINVAL, "gemm_fwd_bias_act: null argument");
    if (!B || !k || !n) return 0;
    if (use_tc(ctx, B, n, k) && activation != DSB200_ACT_SOFTMAX) {
        DSB_PROFILE(ctx, "gemm_fwd_bias_act_tc");
        return gemm_tc_launch(ctx, A, 0, k, W, 1, n, C, n, B, n, k, 1.0f, 0.0f, pBias, activation, slope, alpha, lambda);
    }
    if ((uint64_t)B * k * n <= (1ull << 27) && activation != DSB200_ACT_SOFTMAX && !ctx->noSmallDense) {
        DSB_PROFILE(ctx, "gemm_fwd_bias_act_small");           // one SIMT launch instead of three (dense_small.cu)
        return dense_small_fwd(ctx, B, k, n, A, W, pBias, activation, C, slope, alpha, lambda);
    }
    int rc = dsb200_clear_unit(ctx, C, pBias, n, B);
    if (!rc) rc = dsb200_gemm_fwd(ctx, B, k, n, A, W, 1.0f, C);
    if (!rc && activation != DSB200_ACT_LINEAR) rc = dsb200_activation(ctx, activation, C, B, n, slope, alpha, lambda);
    return rc;
}

// Optional hints: build, AHEAD of the calls that use them and typically on another stream (dsb200_ctx_set_stream), the operands
// that depend only on the data batch (the transposed target bitmap of dsb200_gemm_fwd_output_pass) or only on the weights (the
// hi / lo copies of W of dsb200_gemm_dx).  One-shot: the next matching call consumes them; without a match the call builds its own.
int dsb200_gemm_fwd_output_prepare(dsb200_ctx* ctx, const dsb200_sparse* s, uint32_t position, uint32_t batch, uint32_t n)
{
    using namespace dsb;
    if (!ctx || !s || !s->sparseStart || !s->sparseEnd || !s->sparseIndex) return fail(ctx, DSB200_EINVAL, "gemm_fwd_output_prepare: null argument");
    if ((ctx->gemmMode != DSB200_GEMM_TF32 && ctx->gemmMode != DSB200_GEMM_TF32X3) || !ctx->gemmStream || !gemm_stream_available() || !batch || !n) return 0;
    DSB_PROFILE(ctx, "gemm_fwd_output_prepare");
    return gemm_stream_prepare_targets(ctx, s, position, batch, n);
}

int dsb200_gemm_dx_prepare(dsb200_ctx* ctx, uint32_t B, uint32_t k, uint32_t n, const float* W)
{
    using namespace dsb;
    if (!ctx || !W) return fail(ctx, DSB200_EINVAL, "gemm_dx_prepare: null argument");
    if (!B || !k || !n || !(use_tc(ctx, B, k, n) && use_stream(ctx, B, k, n))) return 0;
    DSB_PROFILE(ctx, "gemm_dx_prepare");
    return gemm_stream_prepare_dx(ctx, k, n, W, n);
}

// Forward pass of a sigmoid output layer over Boolean sparse targets with loss + delta in the GEMM epilogue (gemm_stream.cu,
// out_fwd_kernel): Z and the activations never exist.  DSB200_EUNSUPPORTED = take the two-call path.
int dsb200_gemm_fwd_output_pass(dsb200_ctx* ctx, const dsb200_sparse* s, int errorFunction, int activation, uint32_t position, uint32_t batch,
                                uint32_t k, uint32_t n, const float* A, const float* W, const float* pBias, float* pUnitOut, float* pDelta,
                                unsigned long long* pDevAccumulator, float* pColumnSumPartials, uint32_t* pNumPartials)
{
    using namespace dsb;
    if (!ctx || !s || !A || !W || !pDelta) return fail(ctx, DSB200_EINVAL, "gemm_fwd_output_pass: null argument");
    if (!s->sparseStart || !s->sparseEnd || !s->sparseIndex) return fail(ctx, DSB200_EINVAL, "gemm_fwd_output_pass: incomplete target data set");
    if (pNumPartials) *pNumPartials = 0;
    if (activation != DSB200_ACT_SIGMOID || s->sparseData ||
        (errorFunction != DSB200_ERR_L2 && errorFunction != DSB200_ERR_CROSS_ENTROPY && errorFunction != DSB200_ERR_SMCE))
        return fail(ctx, DSB200_EUNSUPPORTED, "gemm_fwd_output_pass: sigmoid with L2 / CrossEntropy / ScaledMarginalCrossEntropy over Boolean targets only");
    if ((ctx->gemmMode != DSB200_GEMM_TF32 && ctx->gemmMode != DSB200_GEMM_TF32X3) || !ctx->gemmStream || !gemm_stream_available())
        return fail(ctx, DSB200_EUNSUPPORTED, "gemm_fwd_output_pass: needs a tensor-core gemm_mode");
    if (!batch || !k || !n) return 0;
    DSB_PROFILE(ctx, "gemm_fwd_output_pass");
    const int rc = gemm_stream_out_fwd(ctx, s, errorFunction, position, batch, k, n, A, W, n, pBias, pUnitOut, pDelta, n, pDevAccumulator,
                                       pColumnSumPartials, pNumPartials);
    if (rc == DSB200_EUNSUPPORTED) return fail(ctx, DSB200_EUNSUPPORTED, "gemm_fwd_output_pass: hidden width above 128 or rows not 16-byte aligned");
    return rc;
}

}  // extern "C"

extern "C" int dsb200_debug_counters(dsb200_ctx* ctx, unsigned long long* out, size_t count)
{
    if (!ctx || !out) return DSB200_EINVAL;
    DSB_CUDA_OK(cudaDeviceSynchronize());
    return dsb::gemm_stream_debug_counters(out, count);
}
