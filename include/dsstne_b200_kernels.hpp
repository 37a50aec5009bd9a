/* dsstne_b200_kernels.hpp -- the reference's kernel launcher names (E/kernels.h, E = src/amazon/dsstne/engine)
 * re-declared as thin inline C++ wrappers over the C ABI of dsstne_b200.h.
 *
 * A maintainer of the reference who wants the B200 path under the EXISTING NNLayer / NNWeight / NNDataSet code
 * includes this header instead of E/kernels.h for the sparse fully-connected path, links libdsstne_b200.so instead of
 * kernels.o / kLoss.o / kDelta.o / kActivation.o, and calls dsb200k::bind(ctx) once after GpuContext::Startup.  Same
 * names, same argument order and meaning as the declarations cited on each wrapper; failures throw std::runtime_error
 * (the reference's launchers abort through LAUNCHERROR, E/GpuTypes.h:84-92).
 *
 * The hidden __constant__ cData parameters of the reference (shuffle table, denoising p, delta boost, SMCE targets,
 * E/GpuTypes.h:265-311) travel through dsb200_ctx_set_params -- the call that replaces SetKernelsGpuData /
 * SetKLossGpuData / SetKDeltaGpuData (E/GpuTypes.cpp:500-508).
 */
#pragma once
#include <stdexcept>
#include <string>

#include "dsstne_b200.h"

#ifndef NNFloat
typedef float NNFloat;
#endif

namespace dsb200k {

inline dsb200_ctx*& bound() { static dsb200_ctx* c = nullptr; return c; }
inline void bind(dsb200_ctx* ctx) { bound() = ctx; }
inline dsb200_ctx* ctx()
{
    if (!bound()) throw std::runtime_error("dsstne_b200: no context bound (dsb200k::bind)");
    return bound();
}
inline void check(int rc, const char* what)
{
    if (rc != 0) throw std::runtime_error(std::string(what) + " failed: " + dsb200_last_error(bound()));
}

/* NNDataSetEnums::DataType of a sparse value type (E/NNTypes.h:160-211) */
template <typename T> struct dtype_of;
template <> struct dtype_of<uint32_t> { enum { value = DSB200_DT_UINT }; };
template <> struct dtype_of<int32_t> { enum { value = DSB200_DT_INT }; };
template <> struct dtype_of<float> { enum { value = DSB200_DT_FLOAT }; };
template <> struct dtype_of<double> { enum { value = DSB200_DT_DOUBLE }; };
template <> struct dtype_of<unsigned char> { enum { value = DSB200_DT_UCHAR }; };
template <> struct dtype_of<char> { enum { value = DSB200_DT_CHAR }; };

inline dsb200_sparse view(const uint64_t* start, const uint64_t* end, const uint32_t* index, const NNFloat* weight,
                          const void* data = nullptr, int dt = DSB200_DT_FLOAT, const uint32_t* exIndex = nullptr,
                          const NNFloat* random = nullptr)
{
    dsb200_sparse s = {start, end, index, data, dt, weight, exIndex, random};
    return s;
}

}  // namespace dsb200k

/* E/kernels.h:36, :26 */
inline void kClearUnit(NNFloat* pUnit, NNFloat* pBias, uint32_t stride, uint32_t batch)
{ dsb200k::check(dsb200_clear_unit(dsb200k::ctx(), pUnit, pBias, stride, batch), "kClearUnit"); }
inline void kAddBias(NNFloat* pUnit, NNFloat* pBias, uint32_t stride, uint32_t batch)
{ dsb200k::check(dsb200_add_bias(dsb200k::ctx(), pUnit, pBias, stride, batch), "kAddBias"); }
/* E/kernels.h:45 */
inline void kAddBuffers(NNFloat* pDst, NNFloat* pSrc, uint64_t size)
{ dsb200k::check(dsb200_add_buffers(dsb200k::ctx(), pDst, pSrc, size), "kAddBuffers"); }

/* E/kernels.h:67-74 -- sparse-input forward */
inline void kCalculateSparseZ(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pWeight, uint64_t* pSparseStart, uint64_t* pSparseEnd,
                              uint32_t* pSparseIndex, NNFloat* pDataWeight, NNFloat* pUnit, NNFloat beta)
{
    dsb200_sparse s = dsb200k::view(pSparseStart, pSparseEnd, pSparseIndex, pDataWeight);
    dsb200k::check(dsb200_sparse_z(dsb200k::ctx(), &s, position, batch, stride, pWeight, pUnit, beta, 0), "kCalculateSparseZ");
}
inline void kCalculateIndexedSparseZ(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pWeight, uint32_t* pIndex, uint64_t* pSparseStart,
                                     uint64_t* pSparseEnd, uint32_t* pSparseIndex, NNFloat* pDataWeight, NNFloat* pUnit, NNFloat beta)
{
    dsb200_sparse s = dsb200k::view(pSparseStart, pSparseEnd, pSparseIndex, pDataWeight, nullptr, DSB200_DT_FLOAT, pIndex);
    dsb200k::check(dsb200_sparse_z(dsb200k::ctx(), &s, position, batch, stride, pWeight, pUnit, beta, 0), "kCalculateIndexedSparseZ");
}
template <typename T>
inline void kCalculateSparseAnalogZ(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pWeight, uint64_t* pSparseStart, uint64_t* pSparseEnd,
                                    uint32_t* pSparseIndex, NNFloat* pDataWeight, T* pSparseData, NNFloat* pUnit, NNFloat beta)
{
    dsb200_sparse s = dsb200k::view(pSparseStart, pSparseEnd, pSparseIndex, pDataWeight, pSparseData, dsb200k::dtype_of<T>::value);
    dsb200k::check(dsb200_sparse_z(dsb200k::ctx(), &s, position, batch, stride, pWeight, pUnit, beta, 0), "kCalculateSparseAnalogZ");
}
template <typename T>
inline void kCalculateIndexedSparseAnalogZ(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pWeight, uint32_t* pIndex, uint64_t* pSparseStart,
                                           uint64_t* pSparseEnd, uint32_t* pSparseIndex, NNFloat* pDataWeight, T* pSparseData, NNFloat* pUnit, NNFloat beta)
{
    dsb200_sparse s = dsb200k::view(pSparseStart, pSparseEnd, pSparseIndex, pDataWeight, pSparseData, dsb200k::dtype_of<T>::value, pIndex);
    dsb200k::check(dsb200_sparse_z(dsb200k::ctx(), &s, position, batch, stride, pWeight, pUnit, beta, 0), "kCalculateIndexedSparseAnalogZ");
}
inline void kCalculateSparseDenoisedZ(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pWeight, uint64_t* pSparseStart, uint64_t* pSparseEnd,
                                      uint32_t* pSparseIndex, NNFloat* pDataWeight, NNFloat* pRandom, NNFloat* pUnit, NNFloat beta)
{
    dsb200_sparse s = dsb200k::view(pSparseStart, pSparseEnd, pSparseIndex, pDataWeight, nullptr, DSB200_DT_FLOAT, nullptr, pRandom);
    dsb200k::check(dsb200_sparse_z(dsb200k::ctx(), &s, position, batch, stride, pWeight, pUnit, beta, 1), "kCalculateSparseDenoisedZ");
}
inline void kCalculateIndexedSparseDenoisedZ(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pWeight, uint32_t* pIndex, uint64_t* pSparseStart,
                                             uint64_t* pSparseEnd, uint32_t* pSparseIndex, NNFloat* pDataWeight, NNFloat* pRandom, NNFloat* pUnit, NNFloat beta)
{
    dsb200_sparse s = dsb200k::view(pSparseStart, pSparseEnd, pSparseIndex, pDataWeight, nullptr, DSB200_DT_FLOAT, pIndex, pRandom);
    dsb200k::check(dsb200_sparse_z(dsb200k::ctx(), &s, position, batch, stride, pWeight, pUnit, beta, 1), "kCalculateIndexedSparseDenoisedZ");
}
template <typename T>
inline void kCalculateSparseAnalogDenoisedZ(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pWeight, uint64_t* pSparseStart, uint64_t* pSparseEnd,
                                            uint32_t* pSparseIndex, NNFloat* pDataWeight, T* pSparseData, NNFloat* pRandom, NNFloat* pUnit, NNFloat beta)
{
    dsb200_sparse s = dsb200k::view(pSparseStart, pSparseEnd, pSparseIndex, pDataWeight, pSparseData, dsb200k::dtype_of<T>::value, nullptr, pRandom);
    dsb200k::check(dsb200_sparse_z(dsb200k::ctx(), &s, position, batch, stride, pWeight, pUnit, beta, 1), "kCalculateSparseAnalogDenoisedZ");
}

/* E/kernels.h:77-92 -- transposed matrix.  The reference's callers copy Start into End first (E/NNTypes.h:576);
 * these wrappers keep that contract (pass NULL as the capacity table => End is used as initialised by the caller). */
inline void kCalculateSparseTransposedMatrix(uint32_t position, uint32_t batch, uint64_t* pSparseStart, uint64_t* pSparseEnd, uint32_t* pSparseIndex,
                                             NNFloat* pDataWeight, uint32_t* pSparseTransposedEnd, uint32_t* pSparseTransposedIndex,
                                             NNFloat* pSparseTransposedData, uint32_t columns)
{
    dsb200_sparse s = dsb200k::view(pSparseStart, pSparseEnd, pSparseIndex, pDataWeight);
    dsb200k::check(dsb200_sparse_transpose(dsb200k::ctx(), &s, position, batch, 0, columns, nullptr, pSparseTransposedEnd, pSparseTransposedIndex,
                                           pSparseTransposedData), "kCalculateSparseTransposedMatrix");
}
inline void kCalculateSparseTransposedDenoisedMatrix(uint32_t position, uint32_t batch, uint64_t* pSparseStart, uint64_t* pSparseEnd, uint32_t* pSparseIndex,
                                                     NNFloat* pDataWeight, NNFloat* pRandom, uint32_t* pSparseTransposedEnd,
                                                     uint32_t* pSparseTransposedIndex, NNFloat* pSparseTransposedData, uint32_t columns)
{
    dsb200_sparse s = dsb200k::view(pSparseStart, pSparseEnd, pSparseIndex, pDataWeight, nullptr, DSB200_DT_FLOAT, nullptr, pRandom);
    dsb200k::check(dsb200_sparse_transpose(dsb200k::ctx(), &s, position, batch, 1, columns, nullptr, pSparseTransposedEnd, pSparseTransposedIndex,
                                           pSparseTransposedData), "kCalculateSparseTransposedDenoisedMatrix");
}
template <typename T>
inline void kCalculateSparseTransposedAnalogMatrix(uint32_t position, uint32_t batch, uint64_t* pSparseStart, uint64_t* pSparseEnd, uint32_t* pSparseIndex,
                                                   NNFloat* pDataWeight, T* pSparseData, uint32_t* pSparseTransposedEnd, uint32_t* pSparseTransposedIndex,
                                                   NNFloat* pSparseTransposedData, uint32_t columns)
{
    dsb200_sparse s = dsb200k::view(pSparseStart, pSparseEnd, pSparseIndex, pDataWeight, pSparseData, dsb200k::dtype_of<T>::value);
    dsb200k::check(dsb200_sparse_transpose(dsb200k::ctx(), &s, position, batch, 0, columns, nullptr, pSparseTransposedEnd, pSparseTransposedIndex,
                                           pSparseTransposedData), "kCalculateSparseTransposedAnalogMatrix");
}

/* E/kernels.h:81, :93 -- sparse weight gradient */
inline void kCalculateSparseTransposedWeightGradient(NNFloat alpha, NNFloat beta, uint32_t m, uint32_t n, uint32_t* pSparseTransposedStart,
                                                     uint32_t* pSparseTransposedEnd, uint32_t* pSparseTransposedIndex, NNFloat* pDelta, NNFloat* pWeightGradient)
{
    dsb200k::check(dsb200_sparse_wgrad(dsb200k::ctx(), alpha, beta, m, n, pSparseTransposedStart, pSparseTransposedEnd, pSparseTransposedIndex, nullptr,
                                       pDelta, pWeightGradient), "kCalculateSparseTransposedWeightGradient");
}
inline void kCalculateSparseTransposedAnalogWeightGradient(NNFloat alpha, NNFloat beta, uint32_t m, uint32_t n, uint32_t* pSparseTransposedStart,
                                                           uint32_t* pSparseTransposedEnd, uint32_t* pSparseTransposedIndex, NNFloat* pSparseTransposedData,
                                                           NNFloat* pDelta, NNFloat* pWeightGradient)
{
    dsb200k::check(dsb200_sparse_wgrad(dsb200k::ctx(), alpha, beta, m, n, pSparseTransposedStart, pSparseTransposedEnd, pSparseTransposedIndex,
                                       pSparseTransposedData, pDelta, pWeightGradient), "kCalculateSparseTransposedAnalogWeightGradient");
}

/* E/kernels.h:208-214 -- activations (in place) */
inline void kCalculateSigmoidActivation(NNFloat* pData, uint64_t size)
{ dsb200k::check(dsb200_activation(dsb200k::ctx(), DSB200_ACT_SIGMOID, pData, 1, (uint32_t)size, 0, 0, 0), "kCalculateSigmoidActivation"); }
inline void kCalculateTanhActivation(NNFloat* pData, uint64_t size)
{ dsb200k::check(dsb200_activation(dsb200k::ctx(), DSB200_ACT_TANH, pData, 1, (uint32_t)size, 0, 0, 0), "kCalculateTanhActivation"); }
inline void kCalculateRELUActivation(NNFloat* pData, uint64_t size)
{ dsb200k::check(dsb200_activation(dsb200k::ctx(), DSB200_ACT_RELU, pData, 1, (uint32_t)size, 0, 0, 0), "kCalculateRELUActivation"); }
inline void kCalculateLRELUActivation(NNFloat* pData, uint64_t size, NNFloat slope)
{ dsb200k::check(dsb200_activation(dsb200k::ctx(), DSB200_ACT_LRELU, pData, 1, (uint32_t)size, slope, 0, 0), "kCalculateLRELUActivation"); }
inline void kCalculateELUActivation(NNFloat* pData, uint64_t size, NNFloat alpha)
{ dsb200k::check(dsb200_activation(dsb200k::ctx(), DSB200_ACT_ELU, pData, 1, (uint32_t)size, 0, alpha, 0), "kCalculateELUActivation"); }
inline void kCalculateSELUActivation(NNFloat* pData, uint64_t size, NNFloat alpha, NNFloat lambda)
{ dsb200k::check(dsb200_activation(dsb200k::ctx(), DSB200_ACT_SELU, pData, 1, (uint32_t)size, 0, alpha, lambda), "kCalculateSELUActivation"); }
inline void kCalculateSoftMaxActivation(NNFloat* pData, uint32_t batch, uint32_t stride)
{ dsb200k::check(dsb200_activation(dsb200k::ctx(), DSB200_ACT_SOFTMAX, pData, batch, stride, 0, 0, 0), "kCalculateSoftMaxActivation"); }

/* E/kernels.h:113-124 -- losses over sparse targets (value returned after a stream sync, like the reference) */
#define DSB200K_LOSS(NAME, EF, ACT)                                                                                                              \
    inline NNFloat NAME(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pUnit, uint64_t* pSparseStart, uint64_t* pSparseEnd,         \
                        uint32_t* pSparseIndex, NNFloat* pDataWeight, bool bSparseIgnoreZero)                                                    \
    {                                                                                                                                            \
        dsb200_sparse s = dsb200k::view(pSparseStart, pSparseEnd, pSparseIndex, pDataWeight);                                                    \
        float v = 0.0f;                                                                                                                          \
        dsb200k::check(dsb200_sparse_loss(dsb200k::ctx(), &s, EF, ACT, position, batch, stride, pUnit, bSparseIgnoreZero ? 1 : 0, &v), #NAME);    \
        return v;                                                                                                                                \
    }
DSB200K_LOSS(kCalculateSparseL2Error, DSB200_ERR_L2, DSB200_ACT_SIGMOID)
DSB200K_LOSS(kCalculateSparseCrossEntropyError, DSB200_ERR_CROSS_ENTROPY, DSB200_ACT_SIGMOID)
DSB200K_LOSS(kCalculateSparseScaledMarginalCrossEntropyError, DSB200_ERR_SMCE, DSB200_ACT_SIGMOID)
#undef DSB200K_LOSS
inline NNFloat kCalculateSparseMultinomialCrossEntropyError(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pUnit, uint64_t* pSparseStart,
                                                            uint64_t* pSparseEnd, uint32_t* pSparseIndex, NNFloat* pDataWeight)
{
    dsb200_sparse s = dsb200k::view(pSparseStart, pSparseEnd, pSparseIndex, pDataWeight);
    float v = 0.0f;
    dsb200k::check(dsb200_sparse_loss(dsb200k::ctx(), &s, DSB200_ERR_CROSS_ENTROPY, DSB200_ACT_SOFTMAX, position, batch, stride, pUnit, 0, &v),
                   "kCalculateSparseMultinomialCrossEntropyError");
    return v;
}

/* E/kernels.h:174-187 -- output deltas over sparse targets */
/* `activation` is the reference's Activation enum (E/NNTypes.h:90-104; the DSB200_ACT_* values are the same numbers) */
#define DSB200K_DELTA(NAME, EF)                                                                                                                  \
    inline void NAME(int activation, uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pUnit, NNFloat* pDelta, uint64_t* pSparseStart,  \
                     uint64_t* pSparseEnd, uint32_t* pSparseIndex, NNFloat* pDataWeight, bool bSparseIgnoreZero, NNFloat slope = 0.0f,           \
                     NNFloat alpha = 0.0f, NNFloat lambda = 0.0f)                                                                                \
    {                                                                                                                                            \
        dsb200_sparse s = dsb200k::view(pSparseStart, pSparseEnd, pSparseIndex, pDataWeight);                                                    \
        dsb200k::check(dsb200_sparse_output_delta(dsb200k::ctx(), &s, EF, activation, position, batch, stride, pUnit, pDelta,                    \
                                                  bSparseIgnoreZero ? 1 : 0, slope, alpha, lambda), #NAME);                                      \
    }
DSB200K_DELTA(kCalculateSparseOutputDelta, DSB200_ERR_L2)
DSB200K_DELTA(kCalculateSparseCrossEntropyOutputDelta, DSB200_ERR_CROSS_ENTROPY)
DSB200K_DELTA(kCalculateSparseScaledMarginalCrossEntropyOutputDelta, DSB200_ERR_SMCE)
#undef DSB200K_DELTA

/* E/kernels.h:202, :205 */
inline void kCalculateSparsenessPenalty(uint32_t batch, uint32_t stride, NNFloat* pUnit, NNFloat* pDelta, NNFloat p, NNFloat beta)
{ dsb200k::check(dsb200_sparseness_penalty(dsb200k::ctx(), batch, stride, pUnit, pDelta, p, beta), "kCalculateSparsenessPenalty"); }
inline void kCalculateHadamardProduct(int activation, uint64_t size, NNFloat scale, NNFloat* pUnit, NNFloat* pDelta, NNFloat slope, NNFloat alpha,
                                      NNFloat lambda)
{ dsb200k::check(dsb200_hadamard(dsb200k::ctx(), activation, size, scale, pUnit, pDelta, slope, alpha, lambda), "kCalculateHadamardProduct"); }

/* E/kernels.h:127, :217-232 -- optimizers.  `t` of the Adam kernels is the caller's _batches, as in the reference. */
inline NNFloat kCalculateRegularizationError(NNFloat lambda, NNFloat lambda1, NNFloat* pWeight, uint64_t size)
{
    float v = 0.0f;
    dsb200k::check(dsb200_regularization_error(dsb200k::ctx(), lambda, lambda1, pWeight, size, &v), "kCalculateRegularizationError");
    return v;
}
inline void kSGDUpdateWeights(NNFloat alpha, NNFloat lambda, NNFloat lambda1, uint64_t size, NNFloat* pWeightGradient, NNFloat* pWeight)
{ dsb200k::check(dsb200_update_weights(dsb200k::ctx(), DSB200_SGD, alpha, lambda, lambda1, 0, 0, 0, size, nullptr, pWeightGradient, nullptr, pWeight), "kSGDUpdateWeights"); }
inline void kMomentumUpdateWeights(NNFloat alpha, NNFloat lambda, NNFloat lambda1, NNFloat mu, uint64_t size, NNFloat* pWeightVelocity,
                                   NNFloat* pWeightGradient, NNFloat* pWeight)
{ dsb200k::check(dsb200_update_weights(dsb200k::ctx(), DSB200_MOMENTUM, alpha, lambda, lambda1, mu, 0, 0, size, pWeightVelocity, pWeightGradient, nullptr, pWeight), "kMomentumUpdateWeights"); }
inline void kAdaGradUpdateWeights(NNFloat alpha, NNFloat lambda, NNFloat lambda1, uint64_t size, NNFloat* pWeightVelocity, NNFloat* pWeightGradient,
                                  NNFloat* pWeight)
{ dsb200k::check(dsb200_update_weights(dsb200k::ctx(), DSB200_ADAGRAD, alpha, lambda, lambda1, 0, 0, 0, size, pWeightVelocity, pWeightGradient, nullptr, pWeight), "kAdaGradUpdateWeights"); }
inline void kNesterovUpdateWeights(NNFloat alpha, NNFloat lambda, NNFloat lambda1, NNFloat mu, uint64_t size, NNFloat* pWeightVelocity,
                                   NNFloat* pWeightGradient, NNFloat* pWeight)
{ dsb200k::check(dsb200_update_weights(dsb200k::ctx(), DSB200_NESTEROV, alpha, lambda, lambda1, mu, 0, 0, size, pWeightVelocity, pWeightGradient, nullptr, pWeight), "kNesterovUpdateWeights"); }
inline void kRMSPropUpdateWeights(NNFloat alpha, NNFloat lambda, NNFloat lambda1, NNFloat mu, uint64_t size, NNFloat* pWeightVelocity,
                                  NNFloat* pWeightGradient, NNFloat* pWeight)
{ dsb200k::check(dsb200_update_weights(dsb200k::ctx(), DSB200_RMSPROP, alpha, lambda, lambda1, mu, 0, 0, size, pWeightVelocity, pWeightGradient, nullptr, pWeight), "kRMSPropUpdateWeights"); }
inline void kAdaDeltaUpdateWeights(NNFloat lambda, NNFloat lambda1, NNFloat mu, uint64_t size, NNFloat* pWeightVelocity, NNFloat* pWeightGradient,
                                   NNFloat* pWeightGradientVelocity, NNFloat* pWeight)
{ dsb200k::check(dsb200_update_weights(dsb200k::ctx(), DSB200_ADADELTA, 0, lambda, lambda1, mu, 0, 0, size, pWeightVelocity, pWeightGradient, pWeightGradientVelocity, pWeight), "kAdaDeltaUpdateWeights"); }
inline void kAdamUpdateWeights(NNFloat alpha, NNFloat lambda, NNFloat lambda1, NNFloat mu, NNFloat mu1, NNFloat t, uint64_t size, NNFloat* pWeightVelocity,
                               NNFloat* pWeightGradient, NNFloat* pWeightGradientVelocity, NNFloat* pWeight)
{ dsb200k::check(dsb200_update_weights(dsb200k::ctx(), DSB200_ADAM, alpha, lambda, lambda1, mu, mu1, t, size, pWeightVelocity, pWeightGradient, pWeightGradientVelocity, pWeight), "kAdamUpdateWeights"); }
inline void kSGDUpdateBiases(NNFloat alpha, uint32_t batch, uint32_t width, NNFloat* pDelta, NNFloat* pBias)
{ dsb200k::check(dsb200_update_biases(dsb200k::ctx(), DSB200_SGD, alpha, 0, 0, 0, batch, width, pDelta, nullptr, nullptr, pBias), "kSGDUpdateBiases"); }
inline void kMomentumUpdateBiases(NNFloat alpha, NNFloat mu, uint32_t batch, uint32_t width, NNFloat* pDelta, NNFloat* pBiasVelocity, NNFloat* pBias)
{ dsb200k::check(dsb200_update_biases(dsb200k::ctx(), DSB200_MOMENTUM, alpha, mu, 0, 0, batch, width, pDelta, pBiasVelocity, nullptr, pBias), "kMomentumUpdateBiases"); }
inline void kAdaGradUpdateBiases(NNFloat alpha, uint32_t batch, uint32_t width, NNFloat* pDelta, NNFloat* pBiasVelocity, NNFloat* pBias)
{ dsb200k::check(dsb200_update_biases(dsb200k::ctx(), DSB200_ADAGRAD, alpha, 0, 0, 0, batch, width, pDelta, pBiasVelocity, nullptr, pBias), "kAdaGradUpdateBiases"); }
inline void kNesterovUpdateBiases(NNFloat alpha, NNFloat mu, uint32_t batch, uint32_t width, NNFloat* pDelta, NNFloat* pBiasVelocity, NNFloat* pBias)
{ dsb200k::check(dsb200_update_biases(dsb200k::ctx(), DSB200_NESTEROV, alpha, mu, 0, 0, batch, width, pDelta, pBiasVelocity, nullptr, pBias), "kNesterovUpdateBiases"); }
inline void kRMSPropUpdateBiases(NNFloat alpha, NNFloat mu, uint32_t batch, uint32_t width, NNFloat* pDelta, NNFloat* pBiasVelocity, NNFloat* pBias)
{ dsb200k::check(dsb200_update_biases(dsb200k::ctx(), DSB200_RMSPROP, alpha, mu, 0, 0, batch, width, pDelta, pBiasVelocity, nullptr, pBias), "kRMSPropUpdateBiases"); }
inline void kAdaDeltaUpdateBiases(NNFloat mu, uint32_t batch, uint32_t width, NNFloat* pDelta, NNFloat* pBiasVelocity, NNFloat* pBiasGradientVelocity,
                                  NNFloat* pBias)
{ dsb200k::check(dsb200_update_biases(dsb200k::ctx(), DSB200_ADADELTA, 0, mu, 0, 0, batch, width, pDelta, pBiasVelocity, pBiasGradientVelocity, pBias), "kAdaDeltaUpdateBiases"); }
inline void kAdamUpdateBiases(NNFloat alpha, NNFloat mu, NNFloat mu1, NNFloat t, uint32_t batch, uint32_t width, NNFloat* pDelta, NNFloat* pBiasVelocity,
                              NNFloat* pBiasGradientVelocity, NNFloat* pBias)
{ dsb200k::check(dsb200_update_biases(dsb200k::ctx(), DSB200_ADAM, alpha, mu, mu1, t, batch, width, pDelta, pBiasVelocity, pBiasGradientVelocity, pBias), "kAdamUpdateBiases"); }

/* E/kernels.h:41-43 -- top-K.  3-arg: values are the column ids; 4-arg (uint32 values): caller-supplied ids. */
inline void kCalculateTopK(NNFloat* pOutputKey, NNFloat* pKey, uint32_t* pValue, uint32_t batch, uint32_t width, uint32_t k)
{ dsb200k::check(dsb200_topk(dsb200k::ctx(), pOutputKey, batch, width, k, nullptr, nullptr, nullptr, pKey, pValue), "kCalculateTopK"); }
inline void kCalculateTopK(NNFloat* pOutputKey, uint32_t* pOutputValue, NNFloat* pKey, uint32_t* pValue, uint32_t batch, uint32_t width, uint32_t k)
{ dsb200k::check(dsb200_topk_kv(dsb200k::ctx(), pOutputKey, pOutputValue, batch, width, k, pKey, pValue), "kCalculateTopK"); }
/* E/kernels.h:42 -- 4-arg variant with FLOAT values (a second score carried along with the key): the values travel as raw 32-bit words */
inline void kCalculateTopK(NNFloat* pOutputKey, NNFloat* pOutputValue, NNFloat* pKey, NNFloat* pValue, uint32_t batch, uint32_t width, uint32_t k)
{
    dsb200k::check(dsb200_topk_kv(dsb200k::ctx(), pOutputKey, reinterpret_cast<uint32_t*>(pOutputValue), batch, width, k, pKey,
                                  reinterpret_cast<uint32_t*>(pValue)), "kCalculateTopK");
}
