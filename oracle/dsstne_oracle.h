/*
 * dsstne_oracle.h -- CPU restatement of DSSTNE's sparse fully-connected hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may build, load or call this library, and only as
 * the checker / reported CPU baseline.  The product (libdsstne_b200.so) never
 * links or calls it and has no CPU fallback.
 *
 * The reference (amazon-archives/amazon-dsstne) has NO CPU implementation of
 * this path; every function below restates the semantics of a reference CUDA
 * kernel in plain C99 (+OpenMP), citing the kernel it follows as
 * `E/<file>:<lines>` where E = src/amazon/dsstne/engine.
 *
 * Parity pinning: the restatement is checked (a) against the reference's own
 * CUDA kernels, compiled unmodified from /root/reference into
 * oracle/_ref/libdsstne_refkernels.so and run on the GPU box
 * (tests/test_ref_kernels_gpu.py), (b) against the reference's CPU top-K
 * comparator U/Utils.cpp:213-243 compiled into oracle/_ref/libdsstne_refutils.so
 * (tests/test_oracle_topk.py), and (c) by the finite-difference gradient check
 * that NNNetwork::Validate (E/NNNetwork.cpp:2459-2633) applies, re-hosted on
 * this oracle (tests/test_oracle_gradcheck.py) using the reference's
 * tst/test_data/validate_*.json shapes.
 */
#ifndef DSSTNE_ORACLE_H
#define DSSTNE_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* NNDataSetEnums::DataType, E/NNEnum.h:33-45 (values kept identical) */
enum {
    ORC_DT_UINT = 0, ORC_DT_INT = 1, ORC_DT_LLINT = 2, ORC_DT_ULLINT = 3,
    ORC_DT_FLOAT = 4, ORC_DT_DOUBLE = 5, ORC_DT_UCHAR = 8, ORC_DT_CHAR = 9
};

/* Activation, E/NNTypes.h (enum Activation) -- order kept identical */
enum {
    ORC_ACT_SIGMOID = 0, ORC_ACT_TANH = 1, ORC_ACT_RELU = 2, ORC_ACT_LINEAR = 3,
    ORC_ACT_PRELU = 4, ORC_ACT_SOFTPLUS = 5, ORC_ACT_SOFTSIGN = 6, ORC_ACT_SOFTMAX = 7,
    ORC_ACT_RELUMAX = 8, ORC_ACT_LINEARMAX = 9, ORC_ACT_ELU = 10, ORC_ACT_LRELU = 11,
    ORC_ACT_SELU = 12
};

/* ErrorFunction, E/NNTypes.h (enum ErrorFunction) -- order kept identical */
enum {
    ORC_ERR_L1 = 0, ORC_ERR_L2 = 1, ORC_ERR_CROSS_ENTROPY = 2, ORC_ERR_SMCE = 3,
    ORC_ERR_DATA_SMCE = 4, ORC_ERR_HINGE = 5, ORC_ERR_L2HINGE = 6
};

/* TrainingMode, E/NNTypes.h:65-74 */
enum {
    ORC_SGD = 0, ORC_MOMENTUM = 1, ORC_ADAGRAD = 2, ORC_NESTEROV = 3,
    ORC_RMSPROP = 4, ORC_ADADELTA = 5, ORC_ADAM = 6
};

/* Hidden inputs the reference kernels read from `__constant__ GpuData cData`
 * (E/GpuTypes.h:265-311), restricted to the fields this path uses. */
typedef struct orc_params {
    int             bShuffleIndices;
    const uint32_t* pShuffleIndex;
    float           denoising_p;          /* drop nnz when random < p            */
    float           denoising_q;          /* 1/(1-p), E/GpuTypes.cpp:488         */
    float           deltaBoost_one, deltaBoost_zero;
    float           SMCE_oneTarget, SMCE_zeroTarget, SMCE_oneScale, SMCE_zeroScale;
} orc_params;

void orc_params_default(orc_params* p);   /* defaults of E/NNNetwork.cpp:27-58 */

/* A sparse dataset view: "CSR with separate start/end" (E/NNTypes.h:213-225). */
typedef struct orc_csr {
    const uint64_t* sparseStart;    /* [uniqueExamples]                       */
    const uint64_t* sparseEnd;      /* [uniqueExamples]                       */
    const uint32_t* sparseIndex;    /* [nnz]                                   */
    const void*     sparseData;     /* [nnz] of dataType, NULL => Boolean      */
    int             dataType;       /* ORC_DT_*                                */
    const float*    dataWeight;     /* [uniqueExamples] or NULL                */
    const uint32_t* index;          /* [examples] or NULL (Indexed attribute)  */
    const float*    denoisingRandom;/* [nnz] or NULL                           */
} orc_csr;

/* ---- a14: bias init / add (E/kernels.cu:60-80, 564-584) ---- */
void orc_clear_unit(float* unit, const float* bias, uint32_t stride, uint32_t batch);
void orc_add_bias(float* unit, const float* bias, uint32_t stride, uint32_t batch);

/* ---- a1-a3: sparse Z forward (E/kernels.cu:662-1977) ----
 * Z[b,:] = beta*Z[b,:] + w_b * sum_j v_j * W[idx_j,:]; `denoised` selects the
 * *Denoised* family (drop when random<p, scale by q).  Rows with no nnz are
 * left untouched exactly as the reference kernels do (their while-loop body
 * never runs).  fp32 accumulation in CSR order. */
void orc_sparse_z(const orc_params* p, const orc_csr* d, uint32_t position, uint32_t batch,
                  uint32_t stride, const float* W, float* Z, float beta, int denoised);

/* ---- a4: transposed capacity table (host logic, E/NNTypes.cpp:1427-1570) ----
 * N = number of columns (width*height*length).  Writes transposedStart[N] and
 * returns the number of index slots needed (32-aligned running offset). */
uint32_t orc_transposed_capacity(const orc_csr* d, uint32_t examples, uint32_t uniqueExamples,
                                 uint32_t N, uint32_t batch, uint32_t* transposedStart);

/* ---- a5: batch CSR -> CSC counting scatter (E/kernels.cu:1980-2534) ----
 * transposedEnd must hold a copy of transposedStart on entry (the caller does
 * End<-Start, E/NNTypes.h:576).  Entries are emitted in ascending batch-row
 * order per column (the canonical order; the reference's order is arbitrary).
 * transposedData (may be NULL) receives w_b*v (q*w_b for the weighted denoised
 * Boolean kernel, E/kernels.cu:2154). */
void orc_sparse_transpose(const orc_params* p, const orc_csr* d, uint32_t position, uint32_t batch,
                          int denoised, uint32_t* transposedEnd, uint32_t* transposedIndex,
                          float* transposedData);

/* ---- a6: sparse weight gradient (E/kernels.cu:2537-2692) ----
 * dW[c,:] = beta*dW[c,:] + alpha*q * fix( sum_{e in col c} (tdata_e *) delta[row_e,:] )
 * with the sum taken in int64 fixed point, scale 2^30 (llrintf per term).
 * m = number of columns (rows of dW), n = stride of delta/dW.  transposedData
 * NULL => Boolean kernel.  Follows intent for the analog kernel's value load
 * ([tstart], not the reference's [start] typo at E/kernels.cu:2638). */
void orc_sparse_wgrad(const orc_params* p, float alpha, float beta, uint32_t m, uint32_t n,
                      const uint32_t* transposedStart, const uint32_t* transposedEnd,
                      const uint32_t* transposedIndex, const float* transposedData,
                      const float* delta, float* dW);

/* ---- a9: activations (E/kActivation.cu:46-236) ---- */
void orc_activation(int activation, float* data, uint32_t batch, uint32_t stride,
                    float slope, float alpha, float lambda);

/* ---- a7: sparse-target loss (E/kLoss.cu:595-691, 1749-1980, 2213-2352, 2566-2599) ----
 * Returns the summed (not averaged) error as a double; the reference returns
 * (float)(fixed-point sum * 2^-30).  activation only matters for CE/SMCE where
 * SoftMax selects the multinomial variants (dispatch: E/NNTypes.h:840-908). */
double orc_sparse_loss(const orc_params* p, const orc_csr* d, int errorFunction, int activation,
                       uint32_t position, uint32_t batch, uint32_t stride, const float* unit,
                       int sparseIgnoreZero);

/* ---- a8: sparse-target output delta (E/kDelta.cu:2193-2618, 6533-6608, 7182-7305) ---- */
void orc_sparse_output_delta(const orc_params* p, const orc_csr* d, int errorFunction,
                             int activation, uint32_t position, uint32_t batch, uint32_t stride,
                             const float* unit, float* delta, int sparseIgnoreZero,
                             float slope, float alpha, float lambda);

/* ---- a10: hidden backward elementwise (E/kDelta.cu:8979-9151) ---- */
void orc_sparseness_penalty(uint32_t batch, uint32_t stride, const float* unit, float* delta,
                            float p, float beta);
void orc_hadamard(int activation, uint64_t size, float scale, const float* unit, float* delta,
                  float slope, float alpha, float lambda);

/* ---- a11: dense GEMMs as NNLayer issues them (E/NNLayer.cpp:1073, 2223, 2274) ----
 * Row-major.  fwd: C[B][n] = beta*C + A[B][k]*W[k][n]
 *             dw : G[k][n] = beta*G + alpha * A[B][k]^T * D[B][n]
 *             dx : Dp[B][k]= beta*Dp + D[B][n] * W[k][n]^T            */
void orc_gemm_fwd(uint32_t B, uint32_t k, uint32_t n, const float* A, const float* W, float beta, float* C);
void orc_gemm_dw(uint32_t B, uint32_t k, uint32_t n, float alpha, const float* A, const float* D, float beta, float* G);
void orc_gemm_dx(uint32_t B, uint32_t k, uint32_t n, const float* D, const float* W, float beta, float* Dp);

/* ---- a12: optimizers (E/kernels.cu:2719-3199, dispatch E/NNWeight.cpp:718-851) ----
 * `t` is passed as the caller's _batches; the Adam kernels add 1 themselves. */
void orc_update_weights(int mode, float alpha, float lambda, float lambda1, float mu, float mu1,
                        float t, uint64_t size, float* v, const float* g, float* gv, float* w);
void orc_update_biases(int mode, float alpha, float mu, float mu1, float t, uint32_t batch,
                       uint32_t width, const float* delta, float* v, float* gv, float* bias);
double orc_regularization_error(float lambda, float lambda1, const float* w, uint64_t size);

/* ---- a13: top-K (E/kernels.cu:3201-4385; filter U/Filters.cpp:49-67) ----
 * Canonical rule: descending key, ties by ascending index.  Slots beyond the
 * number of candidates hold the reference's sentinel (-MAX_VALUE, 0)
 * (E/NNTypes.h:49, E/kernels.cu:3221-3224).  `filter` (may be NULL) is a CSR of
 * per-row column ids whose score is multiplied by 0.0f before selection
 * (NNRecsGenerator + Filters: score *= filterValue).  `inValue` (may be NULL)
 * is the caller-supplied payload of the 4-argument variants; when NULL the
 * payload is the column index. */
void orc_topk(const float* key, const uint32_t* inValue, uint32_t batch, uint32_t width, uint32_t k,
              const uint64_t* filterStart, const uint64_t* filterEnd, const uint32_t* filterIndex,
              float* outKey, uint32_t* outValue);

/* ---- a15: model-parallel shard ranges (E/NNLayer.cpp:108-112, E/NNWeight.cpp:435-457) ---- */
void orc_shard_range(uint32_t N, uint32_t rank, uint32_t nranks, uint32_t* minX, uint32_t* maxX);
int  orc_weight_outgoing_larger(uint32_t inputStride, uint32_t outputStride);

/* ---- whole-network restatement for sparse-in / sparse-out FC nets ----
 * Mirrors the body of NNNetwork::Train's minibatch loop
 * (E/NNNetwork.cpp:1601-1650): transposed build, forward, error, backward,
 * _batches++, UpdateWeights.  Used for end-to-end parity at small sizes and as
 * the timed CPU baseline. */
#define ORC_MAX_WEIGHTS 8
typedef struct orc_network {
    int        nWeights;                       /* L weight matrices, L+1 layers              */
    uint32_t   size[ORC_MAX_WEIGHTS + 1];      /* size[0]=sparse input width ... size[L]    */
    int        activation[ORC_MAX_WEIGHTS + 1];/* per layer (entry 0 unused)                */
    int        sparsePenalty[ORC_MAX_WEIGHTS + 1]; /* hidden layer has "Sparse": true        */
    int        errorFunction, trainingMode;
    int        denoising;                      /* input layer has Denoising attribute        */
    float      sparsenessPenalty_p, sparsenessPenalty_beta;
    orc_params params;
    uint32_t   maxBatch;
    uint64_t   batches;                        /* _batches                                   */
    float*     W[ORC_MAX_WEIGHTS];  float* b[ORC_MAX_WEIGHTS];
    float*     dW[ORC_MAX_WEIGHTS];
    float*     vW[ORC_MAX_WEIGHTS]; float* gvW[ORC_MAX_WEIGHTS];
    float*     vb[ORC_MAX_WEIGHTS]; float* gvb[ORC_MAX_WEIGHTS];
    float*     unit[ORC_MAX_WEIGHTS + 1];      /* [maxBatch][size[l]], l>=1                  */
    float*     delta[ORC_MAX_WEIGHTS + 1];
    /* transposed input matrix */
    uint32_t*  tStart; uint32_t* tEnd; uint32_t* tIndex; float* tData; uint32_t tCapacity;
    /* dropout of hidden layer l (training only): probability and the caller-supplied uniform randoms [maxBatch][size[l]] */
    float      pDropout[ORC_MAX_WEIGHTS + 1];
    const float* dropoutRandom[ORC_MAX_WEIGHTS + 1];
} orc_network;

/* NNLayer::CalculateDropout (E/NNLayer.cpp:1685-1708) with kCalculateDropout / kCalculateScaledBiasedDropout
 * (E/kernels.cu:4497-4537): unit = (r < p) ? target : scale * unit; Sigmoid drops to 0.5 unscaled, ELU / SELU use the
 * self-normalising affine form, everything else drops to 0 and rescales by 1 / (1 - p).                            */
void orc_dropout(int activation, float* unit, const float* random, uint32_t batch, uint32_t stride, float p, float eluAlpha, float seluLambda);

orc_network* orc_net_create(int nWeights, const uint32_t* sizes, const int* activations,
                            int errorFunction, int trainingMode, uint32_t maxBatch);
void   orc_net_destroy(orc_network* net);
/* Build the capacity table for the input dataset (a4) -- call once per (dataset,batch). */
void   orc_net_set_input(orc_network* net, const orc_csr* in, uint32_t examples, uint32_t uniqueExamples, uint32_t batch);
void   orc_net_forward(orc_network* net, const orc_csr* in, uint32_t position, uint32_t batch, int training);
/* returns (error_training); *reg receives the regularisation error */
double orc_net_train_step(orc_network* net, const orc_csr* in, const orc_csr* out, uint32_t position,
                          uint32_t batch, float alpha, float lambda, float lambda1, float mu, float mu1,
                          double* reg);
/* forward + loss only (what NNNetwork::Validate perturbs around) */
double orc_net_loss(orc_network* net, const orc_csr* in, const orc_csr* out, uint32_t position, uint32_t batch);
/* forward + backward, gradients left in dW / delta (no update) */
void   orc_net_backward(orc_network* net, const orc_csr* in, const orc_csr* out, uint32_t position, uint32_t batch);

int orc_num_threads(void);
/* torchrun exports OMP_NUM_THREADS=1: the CPU arm of bench.py sets the count back to the cores it may run on */
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
