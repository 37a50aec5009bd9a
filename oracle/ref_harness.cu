/*
 * ref_harness.cu -- C entry points onto the REFERENCE's own CUDA kernels.
 *
 * TEST / MEASUREMENT INFRASTRUCTURE ONLY.  This file is compiled together
 * with the reference's unmodified kernel translation units
 *   $(REF)/src/amazon/dsstne/engine/{kernels,kLoss,kDelta,kActivation}.cu
 * (read in place from /root/reference, never copied into this repo) into
 * oracle/_ref/libdsstne_refkernels.so by oracle/Makefile.  It supplies the
 * ~40 lines of GpuContext plumbing those files need (the reference's own
 * GpuTypes.cpp cannot be used: it requires MPI) and re-exports the host
 * launchers declared in E/kernels.h with C linkage and raw device pointers, so
 * tests can run the reference kernels on the B200 next to ours:
 *   - to pin the CPU oracle (oracle/dsstne_oracle.c) against the real thing;
 *   - to time "reference kernels recompiled for sm_100" as a GPU baseline.
 * Nothing in the product links or loads this library.
 */
#include "GpuTypes.h"
#include "NNTypes.h"
#include "kernels.h"

#include <limits>

static GpuContext* g_ctx = nullptr;

GpuContext::GpuContext() :
    _bECCSupport(false), _bCanMapHostMemory(false), _totalMemory(0), _totalCPUMemory(0), _totalGPUMemory(0),
    _bUnifiedMemory(false), _sm_version(SM_6X), _sm_major(10), _threadsPerBlock(SM_6X_THREADS_PER_BLOCK),
    _warpSize(32), _warpBits(5), _warpMask(31), _numprocs(1), _id(0), _device(0),
    _maxSparse(SM_6X_MAXSPARSE), _maxSparseAnalog(SM_6X_MAXSPARSEANALOG), _cuBLASHandle(nullptr), _RNG(nullptr),
    _cuDNNHandle(nullptr), _pNetwork(nullptr), _pbAccumulator(), _bCPUValidate(false), _acceptableError(0.0f),
    _bSingleNode(true), _bP2P(false)
{
    memset(&_data, 0, sizeof(_data));
}
GpuContext::~GpuContext() {}
void GpuContext::Shutdown() {}
void GpuContext::GetMemoryUsage(int* gpuMemory, int* cpuMemory) { *gpuMemory = 0; *cpuMemory = 0; }

struct GpuContext& getGpu()
{
    if (!g_ctx) g_ctx = new GpuContext();
    return *g_ctx;
}

static void push_constants()
{
    SetKernelsGpuData();
    SetKLossGpuData();
    SetKDeltaGpuData();
    SetKActivationGpuData();
}

extern "C" {

int ref_init()
{
    GpuContext& g = getGpu();
    if (!g._pbAccumulator) {
        g._pbAccumulator.reset(new GpuBuffer<unsigned long long int>((unsigned int)1, true));
        g._data._pAccumulator = g._pbAccumulator->_pDevData;
    }
    g._data._warpSize = 32; g._data._warpBits = 5; g._data._warpMask = 31;
    g._data._deltaBoost_one = 1.0f; g._data._deltaBoost_zero = 1.0f;
    g._data._SMCE_oneTarget = 0.9f; g._data._SMCE_zeroTarget = 0.1f;
    g._data._SMCE_oneScale = 1.0f; g._data._SMCE_zeroScale = 1.0f;
    g._data._bDenoising = false; g._data._denoising_p = 0.0f; g._data._denoising_q = 1.0f;
    g._data._bShuffleIndices = false; g._data._pShuffleIndex = nullptr;
    g._data._maxUint32_t = std::numeric_limits<uint32_t>::max();
    g._data._maxInt32_t = std::numeric_limits<int32_t>::max();
    g._data._maxUint64_t = std::numeric_limits<uint64_t>::max();
    g._data._maxInt64_t = std::numeric_limits<int64_t>::max();
    g._data._maxFloat = std::numeric_limits<float>::max();
    g._data._minFloat = std::numeric_limits<float>::min();
    push_constants();
    return (int)cudaDeviceSynchronize();
}

void ref_set_params(int bShuffle, unsigned int* pShuffleIndex, float denoising_p,
                    float deltaBoost_one, float deltaBoost_zero,
                    float oneTarget, float zeroTarget, float oneScale, float zeroScale)
{
    GpuContext& g = getGpu();
    g._data._bShuffleIndices = bShuffle != 0;
    g._data._pShuffleIndex = pShuffleIndex;
    g._data._bDenoising = denoising_p > 0.0f;
    g._data._denoising_p = denoising_p;
    g._data._denoising_q = 1.0f / (1.0f - denoising_p);
    g._data._deltaBoost_one = deltaBoost_one; g._data._deltaBoost_zero = deltaBoost_zero;
    g._data._SMCE_oneTarget = oneTarget; g._data._SMCE_zeroTarget = zeroTarget;
    g._data._SMCE_oneScale = oneScale; g._data._SMCE_zeroScale = zeroScale;
    push_constants();
}

int ref_sync() { return (int)cudaDeviceSynchronize(); }

void ref_clear_unit(float* u, float* b, uint32_t stride, uint32_t batch) { kClearUnit(u, b, stride, batch); }
void ref_add_bias(float* u, float* b, uint32_t stride, uint32_t batch) { kAddBias(u, b, stride, batch); }

/* ---- sparse Z ---- */
void ref_sparse_z(uint32_t position, uint32_t batch, uint32_t stride, float* W, uint32_t* pIndex,
                  uint64_t* s, uint64_t* e, uint32_t* idx, float* dw, float* data, float* rnd, float* unit, float beta)
{
    if (rnd) {
        if (data) {
            if (pIndex) kCalculateIndexedSparseAnalogDenoisedZ<float>(position, batch, stride, W, pIndex, s, e, idx, dw, data, rnd, unit, beta);
            else        kCalculateSparseAnalogDenoisedZ<float>(position, batch, stride, W, s, e, idx, dw, data, rnd, unit, beta);
        } else {
            if (pIndex) kCalculateIndexedSparseDenoisedZ(position, batch, stride, W, pIndex, s, e, idx, dw, rnd, unit, beta);
            else        kCalculateSparseDenoisedZ(position, batch, stride, W, s, e, idx, dw, rnd, unit, beta);
        }
    } else {
        if (data) {
            if (pIndex) kCalculateIndexedSparseAnalogZ<float>(position, batch, stride, W, pIndex, s, e, idx, dw, data, unit, beta);
            else        kCalculateSparseAnalogZ<float>(position, batch, stride, W, s, e, idx, dw, data, unit, beta);
        } else {
            if (pIndex) kCalculateIndexedSparseZ(position, batch, stride, W, pIndex, s, e, idx, dw, unit, beta);
            else        kCalculateSparseZ(position, batch, stride, W, s, e, idx, dw, unit, beta);
        }
    }
}

/* ---- transposed build (caller has already done End <- Start) ---- */
void ref_sparse_transpose(uint32_t position, uint32_t batch, uint32_t* pIndex, uint64_t* s, uint64_t* e, uint32_t* idx,
                          float* dw, float* data, float* rnd, uint32_t* tEnd, uint32_t* tIndex, float* tData)
{
    if (rnd) {
        if (data) {
            if (pIndex) kCalculateIndexedSparseTransposedAnalogDenoisedMatrix<float>(position, batch, pIndex, s, e, idx, dw, data, rnd, tEnd, tIndex, tData);
            else        kCalculateSparseTransposedAnalogDenoisedMatrix<float>(position, batch, s, e, idx, dw, data, rnd, tEnd, tIndex, tData);
        } else {
            if (pIndex) kCalculateIndexedSparseTransposedDenoisedMatrix(position, batch, pIndex, s, e, idx, dw, rnd, tEnd, tIndex, tData);
            else        kCalculateSparseTransposedDenoisedMatrix(position, batch, s, e, idx, dw, rnd, tEnd, tIndex, tData);
        }
    } else {
        if (data) {
            if (pIndex) kCalculateIndexedSparseTransposedAnalogMatrix<float>(position, batch, pIndex, s, e, idx, dw, data, tEnd, tIndex, tData);
            else        kCalculateSparseTransposedAnalogMatrix<float>(position, batch, s, e, idx, dw, data, tEnd, tIndex, tData);
        } else {
            if (pIndex) kCalculateIndexedSparseTransposedMatrix(position, batch, pIndex, s, e, idx, dw, tEnd, tIndex, tData);
            else        kCalculateSparseTransposedMatrix(position, batch, s, e, idx, dw, tEnd, tIndex, tData);
        }
    }
}

void ref_sparse_wgrad(float alpha, float beta, uint32_t m, uint32_t n, uint32_t* tStart, uint32_t* tEnd,
                      uint32_t* tIndex, float* tData, float* delta, float* dW)
{
    if (tData) kCalculateSparseTransposedAnalogWeightGradient(alpha, beta, m, n, tStart, tEnd, tIndex, tData, delta, dW);
    else       kCalculateSparseTransposedWeightGradient(alpha, beta, m, n, tStart, tEnd, tIndex, delta, dW);
}

/* ---- activation ---- */
void ref_activation(int activation, float* data, uint32_t batch, uint32_t stride, float slope, float alpha, float lambda)
{
    uint64_t size = (uint64_t)batch * stride;
    switch ((Activation)activation) {
    case Sigmoid:                 kCalculateSigmoidActivation(data, size); break;
    case Tanh:                    kCalculateTanhActivation(data, size); break;
    case RectifiedLinear:         kCalculateRELUActivation(data, size); break;
    case LeakyRectifiedLinear:    kCalculateLRELUActivation(data, size, slope); break;
    case ExponentialLinear:       kCalculateELUActivation(data, size, alpha); break;
    case ScaledExponentialLinear: kCalculateSELUActivation(data, size, alpha, lambda); break;
    case SoftMax:                 kCalculateSoftMaxActivation(data, batch, stride); break;
    default: break;
    }
}

/* ---- sparse-target loss (Boolean, non-indexed / indexed) ---- */
float ref_sparse_loss(int ef, int activation, uint32_t position, uint32_t batch, uint32_t stride, float* unit,
                      uint32_t* pIndex, uint64_t* s, uint64_t* e, uint32_t* idx, float* dw, int ignoreZero)
{
    bool iz = ignoreZero != 0;
    bool sm = ((Activation)activation == SoftMax);
    switch ((ErrorFunction)ef) {
    case L2:
        return pIndex ? kCalculateIndexedSparseL2Error(position, batch, stride, unit, pIndex, s, e, idx, dw, iz)
                      : kCalculateSparseL2Error(position, batch, stride, unit, s, e, idx, dw, iz);
    case CrossEntropy:
        if (sm) return pIndex ? kCalculateIndexedSparseMultinomialCrossEntropyError(position, batch, stride, unit, pIndex, s, e, idx, dw)
                              : kCalculateSparseMultinomialCrossEntropyError(position, batch, stride, unit, s, e, idx, dw);
        return pIndex ? kCalculateIndexedSparseCrossEntropyError(position, batch, stride, unit, pIndex, s, e, idx, dw, iz)
                      : kCalculateSparseCrossEntropyError(position, batch, stride, unit, s, e, idx, dw, iz);
    case ScaledMarginalCrossEntropy:
        if (sm) return pIndex ? kCalculateIndexedSparseMultinomialScaledMarginalCrossEntropyError(position, batch, stride, unit, pIndex, s, e, idx, dw)
                              : kCalculateSparseMultinomialScaledMarginalCrossEntropyError(position, batch, stride, unit, s, e, idx, dw);
        return pIndex ? kCalculateIndexedSparseScaledMarginalCrossEntropyError(position, batch, stride, unit, pIndex, s, e, idx, dw, iz)
                      : kCalculateSparseScaledMarginalCrossEntropyError(position, batch, stride, unit, s, e, idx, dw, iz);
    default: return 0.0f;
    }
}

/* ---- sparse-target output delta ---- */
void ref_sparse_output_delta(int ef, int activation, uint32_t position, uint32_t batch, uint32_t stride, float* unit,
                             float* delta, uint32_t* pIndex, uint64_t* s, uint64_t* e, uint32_t* idx, float* dw,
                             int ignoreZero, float slope, float alpha, float lambda)
{
    bool iz = ignoreZero != 0;
    Activation a = (Activation)activation;
    switch ((ErrorFunction)ef) {
    case L2:
        if (pIndex) kCalculateIndexedSparseOutputDelta(a, position, batch, stride, unit, delta, pIndex, s, e, idx, dw, iz, slope, alpha, lambda);
        else        kCalculateSparseOutputDelta(a, position, batch, stride, unit, delta, s, e, idx, dw, iz, slope, alpha, lambda);
        break;
    case CrossEntropy:
        if (pIndex) kCalculateIndexedSparseCrossEntropyOutputDelta(a, position, batch, stride, unit, delta, pIndex, s, e, idx, dw, iz);
        else        kCalculateSparseCrossEntropyOutputDelta(a, position, batch, stride, unit, delta, s, e, idx, dw, iz);
        break;
    case ScaledMarginalCrossEntropy:
        if (pIndex) kCalculateIndexedSparseScaledMarginalCrossEntropyOutputDelta(a, position, batch, stride, unit, delta, pIndex, s, e, idx, dw, iz);
        else        kCalculateSparseScaledMarginalCrossEntropyOutputDelta(a, position, batch, stride, unit, delta, s, e, idx, dw, iz);
        break;
    default: break;
    }
}

void ref_sparseness_penalty(uint32_t batch, uint32_t stride, float* unit, float* delta, float p, float beta)
{ kCalculateSparsenessPenalty(batch, stride, unit, delta, p, beta); }

void ref_hadamard(int activation, uint64_t size, float scale, float* unit, float* delta, float slope, float alpha, float lambda)
{ kCalculateHadamardProduct((Activation)activation, size, scale, unit, delta, slope, alpha, lambda); }

/* ---- optimizers ---- */
void ref_update_weights(int mode, float alpha, float lambda, float lambda1, float mu, float mu1, float t,
                        uint64_t size, float* v, float* g, float* gv, float* w)
{
    switch ((TrainingMode)mode) {
    case SGD:      kSGDUpdateWeights(alpha, lambda, lambda1, size, g, w); break;
    case Momentum: kMomentumUpdateWeights(alpha, lambda, lambda1, mu, size, v, g, w); break;
    case AdaGrad:  kAdaGradUpdateWeights(alpha, lambda, lambda1, size, v, g, w); break;
    case Nesterov: kNesterovUpdateWeights(alpha, lambda, lambda1, mu, size, v, g, w); break;
    case RMSProp:  kRMSPropUpdateWeights(alpha, lambda, lambda1, mu, size, v, g, w); break;
    case AdaDelta: kAdaDeltaUpdateWeights(lambda, lambda1, mu, size, v, g, gv, w); break;
    case Adam:     kAdamUpdateWeights(alpha, lambda, lambda1, mu, mu1, t, size, v, g, gv, w); break;
    }
}

void ref_update_biases(int mode, float alpha, float mu, float mu1, float t, uint32_t batch, uint32_t width,
                       float* delta, float* v, float* gv, float* bias)
{
    switch ((TrainingMode)mode) {
    case SGD:      kSGDUpdateBiases(alpha, batch, width, delta, bias); break;
    case Momentum: kMomentumUpdateBiases(alpha, mu, batch, width, delta, v, bias); break;
    case AdaGrad:  kAdaGradUpdateBiases(alpha, batch, width, delta, v, bias); break;
    case Nesterov: kNesterovUpdateBiases(alpha, mu, batch, width, delta, v, bias); break;
    case RMSProp:  kRMSPropUpdateBiases(alpha, mu, batch, width, delta, v, bias); break;
    case AdaDelta: kAdaDeltaUpdateBiases(mu, batch, width, delta, v, gv, bias); break;
    case Adam:     kAdamUpdateBiases(alpha, mu, mu1, t, batch, width, delta, v, gv, bias); break;
    }
}

float ref_regularization_error(float lambda, float lambda1, float* w, uint64_t size)
{ return kCalculateRegularizationError(lambda, lambda1, w, size); }

/* ---- top-K ---- */
void ref_topk3(float* scores, float* outKey, uint32_t* outValue, uint32_t batch, uint32_t width, uint32_t k)
{ kCalculateTopK(scores, outKey, outValue, batch, width, k); }

void ref_topk4(float* outKey, uint32_t* outValue, float* key, uint32_t* value, uint32_t batch, uint32_t width, uint32_t k)
{ kCalculateTopK(outKey, outValue, key, value, batch, width, k); }

}  // extern "C"
