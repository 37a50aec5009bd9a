/*
 * dsstne_b200.h -- C ABI of libdsstne_b200.so: a from-scratch, B200-native (sm_100a)
 * implementation of DSSTNE's sparse fully-connected training / prediction hot path.
 *
 * This header is the DROP-IN BOUNDARY.  It replaces the reference's kernel API,
 * the ~150 C++ free functions of  E/kernels.h:17-243  (E = src/amazon/dsstne/engine
 * of amazon-archives/amazon-dsstne) that NNDataSet<T> (E/NNTypes.h:478-1251), NNLayer
 * (E/NNLayer.cpp:994-2826), NNWeight::UpdateWeights (E/NNWeight.cpp:718-851),
 * NNNetwork::CalculateTopK (E/NNNetwork.cpp:1792-1822) and the recommendation generator
 * (U/NNRecsGenerator.cpp:150) call.  One `extern "C"` function per kernel FAMILY; the
 * reference's hand-expanded variants (Boolean/Analog<T> x Indexed x Denoised x Weighted)
 * are selected by which pointers of `dsb200_sparse` are non-NULL, exactly as the reference
 * selects them from the dataset attributes (E/NNTypes.h:527-649).
 *
 * Conventions
 *  - plain C: raw DEVICE pointers + sizes; no C++/torch types.  NULL optional pointers
 *    mean "absent" as in the reference (pDataWeight == NULL => weight 1, E/kernels.cu:673).
 *  - every call returns 0 on success, else a cudaError_t / DSB200_E* code (the reference
 *    prints and exit(-1)s: LAUNCHERROR, E/GpuTypes.h:215-224); dsb200_last_error() explains.
 *  - kernels borrow pointers for the duration of the call and never allocate; scratch lives
 *    in the context.  All launches go to the context's stream (dsb200_ctx_set_stream);
 *    calls are asynchronous unless documented otherwise.
 *  - the reference's hidden inputs (`__constant__ GpuData cData`, E/GpuTypes.h:265-311,
 *    pushed by GpuContext::SetNeuralNetwork, E/GpuTypes.cpp:475-498) are the explicit
 *    `dsb200_params` block of the context.
 *  - layouts are the reference's: activations/deltas row-major [batch][stride]; weights
 *    row-major W[inputUnits][outputUnits] (E/NNWeight.cpp:284-288); sparse data is CSR with
 *    separate start/end arrays (E/NNTypes.h:213-225).
 *  - there is NO CPU fallback: every entry point needs a CUDA device of compute capability 10.x.
 *
 * The C++ shim include/dsstne_b200_kernels.hpp re-declares the E/kernels.h names on top of
 * this ABI so reference call sites compile unchanged; INTEGRATION.md shows the binding.
 */
#ifndef DSSTNE_B200_H
#define DSSTNE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSB200_VERSION 100

/* error codes beyond cudaError_t */
#define DSB200_EINVAL      10001   /* bad argument                                  */
#define DSB200_EUNSUPPORTED 10002  /* combination outside the hot path (see DESIGN) */
#define DSB200_ENOGPU      10003   /* no sm_100 device: there is no CPU fallback    */
#define DSB200_ENCCL       10004   /* NCCL failure                                  */
#define DSB200_ESTATE      10005   /* call order / missing dataset etc.             */

/* NNDataSetEnums::DataType, E/NNEnum.h:33-45 */
enum { DSB200_DT_UINT = 0, DSB200_DT_INT = 1, DSB200_DT_LLINT = 2, DSB200_DT_ULLINT = 3,
       DSB200_DT_FLOAT = 4, DSB200_DT_DOUBLE = 5, DSB200_DT_UCHAR = 8, DSB200_DT_CHAR = 9 };
/* Activation, E/NNTypes.h:90-104 */
enum { DSB200_ACT_SIGMOID = 0, DSB200_ACT_TANH = 1, DSB200_ACT_RELU = 2, DSB200_ACT_LINEAR = 3,
       DSB200_ACT_SOFTMAX = 7, DSB200_ACT_ELU = 10, DSB200_ACT_LRELU = 11, DSB200_ACT_SELU = 12 };
/* ErrorFunction, E/NNTypes.h:77-86 */
enum { DSB200_ERR_L1 = 0, DSB200_ERR_L2 = 1, DSB200_ERR_CROSS_ENTROPY = 2, DSB200_ERR_SMCE = 3,
       DSB200_ERR_DATA_SMCE = 4, DSB200_ERR_HINGE = 5, DSB200_ERR_L2HINGE = 6 };
/* TrainingMode, E/NNTypes.h:65-74 */
enum { DSB200_SGD = 0, DSB200_MOMENTUM = 1, DSB200_ADAGRAD = 2, DSB200_NESTEROV = 3,
       DSB200_RMSPROP = 4, DSB200_ADADELTA = 5, DSB200_ADAM = 6 };
/* dense-GEMM arithmetic (a11): fp32 SIMT-exact, 1xTF32 tensor core, 3xTF32 split (fp32-grade) */
enum { DSB200_GEMM_FP32 = 0, DSB200_GEMM_TF32 = 1, DSB200_GEMM_TF32X3 = 2 };

typedef struct dsb200_ctx dsb200_ctx;

/* replaces GpuData (E/GpuTypes.h:265-311), restricted to what this path reads */
typedef struct dsb200_params {
    int32_t         bShuffleIndices;      /* _bShuffleIndices                          */
    const uint32_t* pShuffleIndex;        /* _pShuffleIndex (device)                   */
    float           denoising_p;          /* _denoising_p                              */
    float           denoising_q;          /* _denoising_q = 1/(1-p), E/GpuTypes.cpp:488 */
    float           deltaBoost_one, deltaBoost_zero;
    float           SMCE_oneTarget, SMCE_zeroTarget, SMCE_oneScale, SMCE_zeroScale;
} dsb200_params;

/* a sparse dataset as NNDataSet<T> holds it on the device (E/NNTypes.h:213-236) */
typedef struct dsb200_sparse {
    const uint64_t* sparseStart;      /* _pbSparseStart  [uniqueExamples]              */
    const uint64_t* sparseEnd;        /* _pbSparseEnd    [uniqueExamples]              */
    const uint32_t* sparseIndex;      /* _pbSparseIndex  [nnz]                         */
    const void*     sparseData;       /* _pbSparseData   [nnz] or NULL (Boolean)       */
    int32_t         dataType;         /* DSB200_DT_* of sparseData                     */
    const float*    dataWeight;       /* _pbDataWeight   [uniqueExamples] or NULL      */
    const uint32_t* index;            /* _pbIndex        [examples] or NULL (Indexed)  */
    const float*    denoisingRandom;  /* _pbDenoisingRandom [nnz] or NULL              */
} dsb200_sparse;

/* ------------------------------------------------------------------ context
 * replaces getGpu()/GpuContext::Startup/Shutdown/SetNeuralNetwork/CopyConstants
 * (E/GpuTypes.h:375-384, E/GpuTypes.cpp:62-498) for this path.                   */
int  dsb200_version(void);
int  dsb200_ctx_create(dsb200_ctx** out, int device);
int  dsb200_ctx_destroy(dsb200_ctx* ctx);
int  dsb200_ctx_set_stream(dsb200_ctx* ctx, void* cudaStream);
int  dsb200_ctx_set_params(dsb200_ctx* ctx, const dsb200_params* p);
void dsb200_params_default(dsb200_params* p);                 /* E/NNNetwork.cpp:27-58 */
int  dsb200_ctx_reserve(dsb200_ctx* ctx, uint32_t maxBatch, size_t partialFloats);
/* Options (all have working defaults; the alternative kernels stay selectable because the parity tests run every one of them):
 *   "gemm_mode"          DSB200_GEMM_FP32 (default: exact fp32 FMA arithmetic on the SIMT kernel of csrc/dense_small.cu, the 1e-5 parity mode) | DSB200_GEMM_TF32 | DSB200_GEMM_TF32X3
 *   "gemm_loader"        operand path of the tcgen05 GEMM: -1 per shape (default) | 0 cp.async + split warps | 1 registers ->
 *                        shared memory | 2 A operand through tensor memory | 3 the same with a coalesced A loader (experimental)
 *   "gemm_stream"        1 (default) = weight gradient / input delta of layers with a narrow side (<= 256 units) run on the TMA +
 *                        tensor-memory kernels of csrc/gemm_stream.cu; 0 = the general kernel of csrc/gemm_tc.cu for every shape
 *   "gemm_tc_min_work"   tiles x k-iterations below which a GEMM stays off the tensor-core kernels (default 1024)
 *   "gemm_splits"        split-K factor, 0 = automatic;   "gemm_debug"  bring-up switches of csrc/gemm_tc.cu (wrong results)
 *   "transpose_sort"     1 = sort every column of the transposed matrix (canonical order for bit-exact comparison)
 *   "fast_math"          1 (default) = MUFU exp / log / reciprocal in the output pass, as the reference (-use_fast_math); 0 = libm grade
 *   "no_tma" / "z_staged_kernel" / "wgrad_tile_kernel" / "output_tile_kernel" / "no_small_dense"   earlier kernels of a family
 *   "pdl"                1 (default) = the kernels of the training step's main stream are launched with programmatic stream
 *                        serialization (the next kernel's dispatch and set-up overlap the tail of the running one); 0 = plain launches
 *   "profile"            1 = bracket every entry point with CUDA events (dsb200_profile_report)                              */
int  dsb200_ctx_set_option(dsb200_ctx* ctx, const char* name, int value);
int  dsb200_ctx_sync(dsb200_ctx* ctx);
const char* dsb200_last_error(dsb200_ctx* ctx);
uint64_t dsb200_launch_count(void);      /* kernels of this library launched so far */
/* with option "profile" = 1 every kernel entry is bracketed by CUDA events on the stream; this call
 * synchronises and writes one line per family, "name calls total_ms\n", then clears the records */
int  dsb200_profile_report(dsb200_ctx* ctx, char* buf, size_t cap);
/* bring-up: the clock64 counters the kernels of csrc/gemm_stream.cu leave behind under "gemm_debug" & 65536 ([CTA][16] cycle sums,
 * layout in tools/fwd_probe.py); synchronises the device */
int  dsb200_debug_counters(dsb200_ctx* ctx, unsigned long long* out, size_t count);

/* ------------------------------------------------------------------ a14
 * kClearUnit / kAddBias, E/kernels.h:30,26 (E/kernels.cu:60-80, 564-584)          */
int dsb200_clear_unit(dsb200_ctx*, float* pUnit, const float* pBias, uint32_t stride, uint32_t batch);
int dsb200_add_bias(dsb200_ctx*, float* pUnit, const float* pBias, uint32_t stride, uint32_t batch);
/* kAddBuffers, E/kernels.h:45 (E/kernels.cu:39-57): pDst[i] += pSrc[i] */
int dsb200_add_buffers(dsb200_ctx*, float* pDst, const float* pSrc, uint64_t size);

/* ------------------------------------------------------------------ a1-a3
 * kCalculate[Indexed]Sparse[Analog][Denoised]Z, E/kernels.h:67-74 (E/kernels.cu:662-1977).
 *   pUnit[b,:] = beta*pUnit[b,:] + w_b * sum_j v_j * pWeight[idx_j,:]
 * `denoised` != 0 selects the *Denoised* family (needs s->denoisingRandom).          */
int dsb200_sparse_z(dsb200_ctx*, const dsb200_sparse* s, uint32_t position, uint32_t batch, uint32_t stride,
                    const float* pWeight, float* pUnit, float beta, int denoised);
/* fused forward of a sparse-input layer: kClearUnit + sparse Z (beta=1) + activation
 * (E/NNLayer.cpp:1009,1052-1054,1157) in one pass; activation = DSB200_ACT_* (not SoftMax) */
int dsb200_sparse_z_bias_act(dsb200_ctx*, const dsb200_sparse* s, uint32_t position, uint32_t batch, uint32_t stride,
                             const float* pWeight, const float* pBias, int activation, float* pUnit, int denoised);

/* ------------------------------------------------------------------ a5
 * kCalculate[Indexed]SparseTransposed[Analog][Denoised]Matrix, E/kernels.h:77-92
 * (E/kernels.cu:1980-2534) preceded by End<-Start (E/NNTypes.h:576): pass the capacity
 * table in pTransposedStart (N entries) and the copy is fused; pass NULL if the caller
 * already initialised pTransposedEnd.  pTransposedData may be NULL (Boolean unweighted).
 * Output is canonical: ascending batch row inside each column.                        */
int dsb200_sparse_transpose(dsb200_ctx*, const dsb200_sparse* s, uint32_t position, uint32_t batch, int denoised,
                            uint32_t N, const uint32_t* pTransposedStart, uint32_t* pTransposedEnd,
                            uint32_t* pTransposedIndex, float* pTransposedData);

/* a4 on the device: the capacity table of NNDataSet<T>::GenerateSparseTransposedMatrix (E/NNTypes.cpp:1631-1735) for a
 * non-indexed dataset that is replaced every step: per-column entry counts over the first `rows` examples (exact, hence
 * an upper bound for any batch window), pTransposedStart = exclusive prefix of the counts rounded up to multiples of 32.
 * pCountScratch: N uint32 of scratch; pDevTotal (device, may be NULL) receives the total capacity.                       */
int dsb200_transposed_capacity(dsb200_ctx*, const dsb200_sparse* s, uint32_t rows, uint32_t N, uint32_t* pCountScratch,
                               uint32_t* pTransposedStart, uint32_t* pDevTotal);

/* ------------------------------------------------------------------ a6
 * kCalculateSparseTransposed[Analog]WeightGradient, E/kernels.h:81,93 (E/kernels.cu:2537-2692)
 *   dW[c,:] = beta*dW[c,:] + alpha*q * sum_{e in column c} (tdata_e *) pDelta[row_e,:]
 * m = columns (rows of dW), n = stride; sum in 2^30 fixed point => order independent.   */
int dsb200_sparse_wgrad(dsb200_ctx*, float alpha, float beta, uint32_t m, uint32_t n,
                        const uint32_t* pTransposedStart, const uint32_t* pTransposedEnd,
                        const uint32_t* pTransposedIndex, const float* pTransposedData,
                        const float* pDelta, float* pWeightGradient);
/* fused a6 + a12 for beta == 0, unshared weights: the gradient row is applied straight to
 * the weight row (and optimizer state); dW is never written.  Same arithmetic as
 * dsb200_sparse_wgrad followed by dsb200_update_weights.                                 */
int dsb200_sparse_wgrad_update(dsb200_ctx*, int mode, float galpha, uint32_t m, uint32_t n,
                               const uint32_t* pTransposedStart, const uint32_t* pTransposedEnd,
                               const uint32_t* pTransposedIndex, const float* pTransposedData, const float* pDelta,
                               float alpha, float lambda, float lambda1, float mu, float mu1, float t,
                               float* pWeightVelocity, float* pWeightGradientVelocity, float* pWeight);

/* ------------------------------------------------------------------ a9
 * kCalculate{Sigmoid,Tanh,RELU,LRELU,ELU,SELU,SoftMax}Activation, E/kernels.h:208-214   */
int dsb200_activation(dsb200_ctx*, int activation, float* pData, uint32_t batch, uint32_t stride,
                      float slope, float alpha, float lambda);

/* ------------------------------------------------------------------ a7
 * kCalculate[Indexed]Sparse{L2,CrossEntropy,ScaledMarginalCrossEntropy,Multinomial...}Error,
 * E/kernels.h:113-124 (E/kLoss.cu:595-691,1749-1980,2213-2352,2566-2599).  Like the reference
 * this variant returns the loss BY VALUE after synchronising the stream.                  */
int dsb200_sparse_loss(dsb200_ctx*, const dsb200_sparse* s, int errorFunction, int activation,
                       uint32_t position, uint32_t batch, uint32_t stride, const float* pUnit,
                       int bSparseIgnoreZero, float* pLossOut);
/* asynchronous variant: adds the fixed-point (2^30) loss into *pDevAccumulator (device u64) */
int dsb200_sparse_loss_async(dsb200_ctx*, const dsb200_sparse* s, int errorFunction, int activation,
                             uint32_t position, uint32_t batch, uint32_t stride, const float* pUnit,
                             int bSparseIgnoreZero, unsigned long long* pDevAccumulator);

/* ------------------------------------------------------------------ a8
 * kCalculate[Indexed]Sparse{,CrossEntropy,ScaledMarginalCrossEntropy}OutputDelta,
 * E/kernels.h:174-187 (E/kDelta.cu:2193-2618, 6533-6608, 7182-7305)                       */
int dsb200_sparse_output_delta(dsb200_ctx*, const dsb200_sparse* s, int errorFunction, int activation,
                               uint32_t position, uint32_t batch, uint32_t stride, const float* pUnit,
                               float* pDelta, int bSparseIgnoreZero, float slope, float alpha, float lambda);
/* fused output pass: activation (a9) + loss (a7) + delta (a8) in ONE read of Z and ONE write
 * of delta (sigmoid / linear activations).  pUnitOut may be NULL (training does not need the
 * activations of a sparse-target output layer again); pDevAccumulator may be NULL.         */
int dsb200_output_pass(dsb200_ctx*, const dsb200_sparse* s, int errorFunction, int activation,
                       uint32_t position, uint32_t batch, uint32_t stride, const float* pZ,
                       float* pUnitOut, float* pDelta, unsigned long long* pDevAccumulator);
/* cublasSgemm of the output layer (E/NNLayer.cpp:1073) + kCalculateSigmoidActivation (E/kActivation.cu:46-64) + the Raw / NonZero
 * loss kernels (E/kLoss.cu:595-691, 1749-1865, 2213-2352) + the Raw / NonZero delta kernels (E/kDelta.cu:2193-2232, 6533-6572,
 * 7182-7227) as ONE tcgen05 kernel (csrc/gemm_stream.cu): the epilogue turns the accumulator into loss and delta, so neither Z nor
 * the activations are written or re-read (pUnitOut, optional, receives the activations).  A [batch][k], W [k][n], bias [n].
 * Sigmoid with L2 / CrossEntropy / ScaledMarginalCrossEntropy over Boolean targets, k <= 128, a tensor-core gemm_mode: any other
 * combination returns DSB200_EUNSUPPORTED and the caller makes the two calls (dsb200_gemm_fwd_bias_act, dsb200_output_pass).
 * pColumnSumPartials (optional, device, capacity 4 * ceil(batch / 128) * n floats) receives *pNumPartials rows of [n] partial
 * column sums of delta -- the bias gradient of E/NNWeight.cpp:760-794 -- for dsb200_update_biases_partials.                       */
int dsb200_gemm_fwd_output_pass(dsb200_ctx*, const dsb200_sparse* s, int errorFunction, int activation, uint32_t position, uint32_t batch,
                                uint32_t k, uint32_t n, const float* A, const float* W, const float* pBias, float* pUnitOut, float* pDelta,
                                unsigned long long* pDevAccumulator, float* pColumnSumPartials, uint32_t* pNumPartials);

/* optional hints for the two calls above / below: build ahead of time -- typically on another stream, next to the forward pass --
 * what depends only on the data batch (target bitmap of dsb200_gemm_fwd_output_pass) or only on the weights (hi / lo copies of W
 * of dsb200_gemm_dx / dsb200_gemm_dx_hadamard).  One-shot; a call without a matching hint prepares its own operands.            */
int dsb200_gemm_fwd_output_prepare(dsb200_ctx*, const dsb200_sparse* s, uint32_t position, uint32_t batch, uint32_t n);
int dsb200_gemm_dx_prepare(dsb200_ctx*, uint32_t B, uint32_t k, uint32_t n, const float* W);

/* ------------------------------------------------------------------ a10
 * kCalculateSparsenessPenalty / kCalculateHadamardProduct, E/kernels.h:202,205            */
int dsb200_sparseness_penalty(dsb200_ctx*, uint32_t batch, uint32_t stride, const float* pUnit, float* pDelta,
                              float p, float beta);
int dsb200_hadamard(dsb200_ctx*, int activation, uint64_t size, float scale, const float* pUnit, float* pDelta,
                    float slope, float alpha, float lambda);

/* ------------------------------------------------------------------ a11
 * the cublasSgemm calls of NNLayer (E/NNLayer.cpp:1073, 2223, 2274), row-major:
 *   fwd: C[B][n]  = beta*C  + A[B][k] * W[k][n]        (+ optional bias/activation epilogue)
 *   dw : G[k][n]  = beta*G  + alpha * A[B][k]^T * D[B][n]
 *   dx : Dp[B][k] = beta*Dp + D[B][n] * W[k][n]^T                                        */
int dsb200_gemm_fwd(dsb200_ctx*, uint32_t B, uint32_t k, uint32_t n, const float* A, const float* W, float beta, float* C);
int dsb200_gemm_dw(dsb200_ctx*, uint32_t B, uint32_t k, uint32_t n, float alpha, const float* A, const float* D, float beta, float* G);
int dsb200_gemm_dx(dsb200_ctx*, uint32_t B, uint32_t k, uint32_t n, const float* D, const float* W, float beta, float* Dp);
/* fused input delta of a SMALL dense layer: Dp = (D * W^T) (.) f'(pUnit) * scale -- cublasSgemm (E/NNLayer.cpp:2274) +
 * kCalculateHadamardProduct of the layer below (E/NNLayer.cpp:2137) in one SIMT launch (csrc/dense_small.cu)            */
int dsb200_gemm_dx_hadamard(dsb200_ctx*, uint32_t B, uint32_t k, uint32_t n, const float* D, const float* W, int activation, float scale,
                            const float* pUnit, float* Dp, float slope, float alpha, float lambda);
/* fused dense forward C = act(A*W + bias): kClearUnit + cublasSgemm(beta=1) + activation (E/NNLayer.cpp:1009,1073,1157) */
int dsb200_gemm_fwd_bias_act(dsb200_ctx*, uint32_t B, uint32_t k, uint32_t n, const float* A, const float* W, const float* pBias,
                             int activation, float* C, float slope, float alpha, float lambda);

/* ------------------------------------------------------------------ a12
 * k{SGD,Momentum,AdaGrad,Nesterov,RMSProp,AdaDelta,Adam}Update{Weights,Biases} and
 * kCalculateRegularizationError, E/kernels.h:127,217-232 (E/kernels.cu:2719-3199).
 * `t` is the caller's _batches; the Adam kernels add 1 themselves (E/kernels.cu:2991).     */
int dsb200_update_weights(dsb200_ctx*, int mode, float alpha, float lambda, float lambda1, float mu, float mu1, float t,
                          uint64_t size, float* pWeightVelocity, const float* pWeightGradient,
                          float* pWeightGradientVelocity, float* pWeight);
int dsb200_update_biases(dsb200_ctx*, int mode, float alpha, float mu, float mu1, float t, uint32_t batch, uint32_t width,
                         const float* pDelta, float* pBiasVelocity, float* pBiasGradientVelocity, float* pBias);
/* k*UpdateBiases with the column sums of delta already reduced to nPartials rows of [width] (dsb200_gemm_fwd_output_pass):
 * gbar[c] = sum_p pPartials[p][c] / batch, summed in a fixed order                                                         */
int dsb200_update_biases_partials(dsb200_ctx*, int mode, float alpha, float mu, float mu1, float t, uint32_t batch, uint32_t width,
                                  const float* pPartials, uint32_t nPartials, float* pBiasVelocity, float* pBiasGradientVelocity, float* pBias);
/* small dense layer: cublasSgemm weight gradient (E/NNLayer.cpp:2223, beta = 0) + k*UpdateWeights + k*UpdateBiases
 * (E/NNWeight.cpp:729-794) in ONE launch -- g = galpha * X[B][k]^T * D[B][n] is applied to W[k][n] and never written, the bias
 * rule runs on the column means of D (pBias NULL = weights only).  Exact fp32; B <= 4,096 (csrc/dense_small.cu).          */
int dsb200_dense_update(dsb200_ctx*, int mode, uint32_t B, uint32_t k, uint32_t n, float galpha, const float* X, const float* D,
                        float alpha, float lambda, float lambda1, float mu, float mu1, float t, float* pWeightVelocity,
                        float* pWeightGradientVelocity, float* pWeight, float* pBiasVelocity, float* pBiasGradientVelocity, float* pBias);
int dsb200_regularization_error(dsb200_ctx*, float lambda, float lambda1, const float* pWeight, uint64_t size,
                                float* pErrorOut);            /* synchronous, by value     */
/* asynchronous variant: adds the fixed-point (2^30) value into *pDevAccumulator (device u64) */
int dsb200_regularization_error_async(dsb200_ctx*, float lambda, float lambda1, const float* pWeight, uint64_t size,
                                      unsigned long long* pDevAccumulator);

/* ------------------------------------------------------------------ a13
 * kCalculateTopK 3-arg (E/kernels.h:41) and 4-arg (E/kernels.h:42-43), E/kernels.cu:3201-4385.
 * Output rows are sorted by descending key, ties by ascending column; unused slots hold the
 * reference sentinel (-MAX_VALUE, 0).  The optional filter (device CSR of column ids per batch
 * row) applies U/Filters.cpp:49-67 in-kernel: score *= 0.0f at the listed columns.         */
int dsb200_topk(dsb200_ctx*, const float* pScores, uint32_t batch, uint32_t width, uint32_t k,
                const uint64_t* pFilterStart, const uint64_t* pFilterEnd, const uint32_t* pFilterIndex,
                float* pOutKey, uint32_t* pOutValue);
int dsb200_topk_kv(dsb200_ctx*, const float* pKey, const uint32_t* pValue, uint32_t batch, uint32_t width, uint32_t k,
                   float* pOutKey, uint32_t* pOutValue);
/* pValue[i] += offset for i < n: local column ids -> global ids (the reference rebuilds them on the host as
 * gpuId * localStride + local, U/NNRecsGenerator.cpp:214) before the per-rank lists are merged with dsb200_topk_kv */
int dsb200_topk_offset(dsb200_ctx*, uint32_t* pValue, uint64_t n, uint32_t offset);

/* ------------------------------------------------------------------ denoising randoms ("next" row 4)
 * replaces the whole-dataset curandGenerateUniform fill of NNDataSet::GenerateDenoisingData
 * (E/NNTypes.cpp:1617-1629): counter-based generator, pOut[i] = U(0,1] as a pure function of
 * (seed, stream, i) -- reproducible for any launch geometry.                                    */
int dsb200_fill_uniform(dsb200_ctx*, float* pOut, uint64_t n, uint64_t seed, uint64_t stream);
/* NNLayer::CalculateDropout (E/NNLayer.cpp:1685-1708; kCalculateDropout / kCalculateScaledBiasedDropout, E/kernels.h:48-49,
 * E/kernels.cu:4497-4537), random draw fused in: pUnit[b][c] = (r < p) ? target : a * pUnit[b][c] + b' with r the uniform
 * dsb200_fill_uniform would give element b * fullStride + colOffset + c (fullStride / colOffset: the un-sharded layer width and
 * this rank's first unit, so the mask is independent of the sharding).  Sigmoid drops to 0.5 unscaled, ELU / SELU use the
 * self-normalising affine form, every other activation drops to 0 and rescales by 1 / (1 - p).                              */
int dsb200_dropout(dsb200_ctx*, int activation, float* pUnit, uint32_t batch, uint32_t stride, uint32_t fullStride, uint32_t colOffset,
                   float p, float eluAlpha, float seluLambda, uint64_t seed, uint64_t stream);

/* ------------------------------------------------------------------ a15
 * NNLayer::Reduce / Gather and NNNetwork::P2P_Allreduce (E/NNLayer.cpp:2702-2826,
 * E/NNNetwork.cpp:4127-4197) as NCCL collectives over NVLink.  `uniqueId` is the 128-byte
 * ncclUniqueId produced by dsb200_comm_unique_id on rank 0 and broadcast by the launcher.  */
int dsb200_comm_unique_id(void* uniqueId128);
int dsb200_comm_init(dsb200_ctx*, const void* uniqueId128, int rank, int nranks);
int dsb200_comm_destroy(dsb200_ctx*);
/* [batch][stride] summed over ranks; rank r keeps columns [stride*r/P, stride*(r+1)/P) in pOut [batch][span] */
int dsb200_reduce_scatter(dsb200_ctx*, uint32_t batch, uint32_t stride, const float* pIn, float* pOut);
/* local slice [batch][span_r] of every rank -> full [batch][stride] on every rank */
int dsb200_all_gather(dsb200_ctx*, uint32_t batch, uint32_t stride, const float* pLocal, float* pFull);
int dsb200_all_reduce(dsb200_ctx*, float* pBuffer, uint64_t size);
/* The same exchange steps as ONE kernel each over peer memory (csrc/comm.cu): every rank owns an arena of `slots` regions of
 * `slotFloats` floats, mapped by every peer (cudaIpc); a collective names the slot it delivers into, so the engine can keep the
 * gathered operand of every layer boundary until that boundary is exchanged again.
 *   dsb200_p2p_setup           collective; DSB200_EUNSUPPORTED when any rank could not map its peers -> stay on the NCCL calls above
 *   dsb200_p2p_slot            local address of a slot: the full [batch][stride] result of an all-gather into it
 *   dsb200_p2p_all_gather      local slice [batch][span_r] of every rank -> slot (full [batch][stride]) on every rank
 *   dsb200_p2p_reduce_scatter  [batch][stride] summed over the ranks in rank order; this rank keeps its columns as
 *                              pOut[batch][span] = activation(sum + pBias[span]) -- pBias may be NULL, activation DSB200_ACT_LINEAR
 *                              (kAddBias + kCalculate*Activation of E/NNLayer.cpp:1257-1340 fused into the exchange)           */
int    dsb200_p2p_setup(dsb200_ctx*, uint32_t slots, size_t slotFloats);
float* dsb200_p2p_slot(dsb200_ctx*, uint32_t slot);
int    dsb200_p2p_all_gather(dsb200_ctx*, uint32_t slot, uint32_t batch, uint32_t stride, const float* pLocal);
int    dsb200_p2p_reduce_scatter(dsb200_ctx*, uint32_t slot, uint32_t batch, uint32_t stride, const float* pIn, float* pOut, const float* pBias,
                                 int activation, float slope, float alpha, float lambda);
int dsb200_all_reduce_u64(dsb200_ctx*, unsigned long long* pBuffer, uint64_t size);
/* model-parallel partition rules: E/NNLayer.cpp:108-112, E/NNWeight.cpp:435-457 (host only, no GPU) */
void dsb200_shard_range(uint32_t N, uint32_t rank, uint32_t nranks, uint32_t* pMinX, uint32_t* pMaxX);
int  dsb200_weight_outgoing_larger(uint32_t inputStride, uint32_t outputStride);

#ifdef __cplusplus
}
#endif
#endif /* DSSTNE_B200_H */
