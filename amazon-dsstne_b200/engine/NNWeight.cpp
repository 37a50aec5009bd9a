// NNWeight.cpp -- weights between two fully-connected layers on the dsstne_b200 C ABI.
//
// Follows E/NNWeight.cpp: constructor + sharding rule (:411-470), Randomize (:500-558),
// RefreshState (:573-602), CalculateRegularizationError (:707-714), UpdateWeights (:718-851),
// Set/Get (:938-1100).  Differences by design:
//  * Randomize draws from a counter-based generator indexed by the GLOBAL element position, so a
//    model-parallel run holds exactly the slices of the single-GPU matrix (cuRAND's stream depends
//    on the rank and cannot be reproduced -- SURVEY 8c "parity unpinned");
//  * UpdateWeights can run the sparse input gradient and the optimizer rule as ONE kernel
//    (dsb200_sparse_wgrad_update) when BackPropagate deferred the gradient.
#include "NNWeight.h"

#include <algorithm>
#include <cmath>
#include <thread>
#include <vector>

#include "NNLayer.h"
#include "NNNetwork.h"

using namespace std;

NNWeight::NNWeight(NNLayer& inputLayer, NNLayer& outputLayer, bool bShared, bool bTransposed, bool bLocked, NNFloat maxNorm)
    : _inputLayer(inputLayer), _outputLayer(outputLayer), _bShared(bShared), _bTransposed(bTransposed), _transform(Linear),
      _bLocked(bLocked), _pSharedWeight(NULL), _sharingCount(1), _updateCount(0), _width(0), _height(0), _size(0), _biasSize(0),
      _localSize(0), _localBiasSize(0), _bOutgoingLarger(false), _norm(maxNorm), _bDeferredSparseGradient(false), _pDeferredDelta(NULL), _nBiasPartials(0)
{
    if (bShared || bTransposed) throw DsbEngineError("NNWeight: shared / transposed weights are outside the hot path");
    if (maxNorm > (NNFloat)0.0) throw DsbEngineError("NNWeight: WeightNorm is outside the hot path");
    inputLayer._vOutgoingLayer.push_back(&outputLayer);
    outputLayer._vIncomingLayer.push_back(&inputLayer);
    inputLayer._vOutgoingWeight.push_back(this);
    outputLayer._vIncomingWeight.push_back(this);
    // E/NNWeight.cpp:435-457: the larger side of the matrix is the one that is split across ranks
    const uint32_t outgoingSize = outputLayer._stride * 3;
    const uint32_t incomingSize = inputLayer._stride * 2;
    if (outgoingSize > incomingSize) {
        _bOutgoingLarger = true;
        inputLayer._vOutgoingLargerLayer.push_back(&outputLayer);
        inputLayer._vOutgoingLargerWeight.push_back(this);
        _width = outputLayer._localStride;
        _height = inputLayer._stride;
    } else {
        outputLayer._vIncomingLargerLayer.push_back(&inputLayer);
        outputLayer._vIncomingLargerWeight.push_back(this);
        _width = outputLayer._stride;
        _height = inputLayer._localStride;
    }
    _localSize = _width * _height;
    _localBiasSize = outputLayer._localStride;
    _size = (uint64_t)outputLayer._stride * inputLayer._stride;
    _biasSize = outputLayer._stride;
    _vWeight.resize(_localSize);
    _pbWeight.reset(new GpuBuffer<NNFloat>(_localSize));
    _pbWeightGradient.reset(new GpuBuffer<NNFloat>(_localSize));
    _vBias.resize(_localBiasSize);
    _pbBias.reset(new GpuBuffer<NNFloat>(_localBiasSize));
}

NNWeight::~NNWeight() {}

// ---- counter-based host generator (SplitMix64 finaliser), keyed by global element index ----
static inline uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
static inline float u01(uint64_t key, uint64_t i)          // (0, 1]
{
    return ((float)(uint32_t)(mix64(key + i * 0x9e3779b97f4a7c15ull) >> 40) + 1.0f) * (1.0f / 16777216.0f);
}

void NNWeight::Randomize()
{
    // formulas of E/NNWeight.cpp:500-558 (uniform u in (0,1]: w = scale*u - bias; Gaussian: N(0, sigma))
    // the key mixes in the names of the two layers: cuRAND's stream advances from one matrix to the next (E/NNWeight.cpp:500-558), so
    // equal-shaped matrices of a stacked network must not start out identical; the element index stays GLOBAL, so model-parallel
    // shards still hold exactly the slices of the single-GPU matrix
    uint64_t nameHash = 0xcbf29ce484222325ull;
    for (unsigned char ch : _inputLayer._name + "\x1f" + _outputLayer._name) nameHash = (nameHash ^ ch) * 0x100000001b3ull;
    const uint64_t key = mix64((uint64_t)getGpu()._seed * 0x9e3779b97f4a7c15ull + mix64((uint64_t)_inputLayer._stride * 1000003ull + _outputLayer._stride) + mix64(nameHash));
    const uint32_t inMin = _bOutgoingLarger ? 0 : _inputLayer._minX;
    const uint32_t outMin = _bOutgoingLarger ? _outputLayer._minX : 0;
    NNFloat scale = 0.0f, bias = 0.0f, sigma = 0.0f;
    bool gaussian = false, constant = false;
    switch (_outputLayer._weightInit) {
    case CaffeXavier: scale = _outputLayer._weightInitScale * 2.0f * sqrtf(3.0f / _outputLayer._stride); bias = 0.5f * scale; break;
    case Xavier:      scale = _outputLayer._weightInitScale * sqrtf(6.0f / (_outputLayer._stride + _inputLayer._stride)); bias = 0.5f * scale; break;
    case Uniform:     scale = 2.0f * _outputLayer._weightInitScale; bias = 0.5f * scale; break;
    case Gaussian:    gaussian = true; sigma = _outputLayer._weightInitScale; break;
    case UnitBall:    scale = _outputLayer._weightInitScale; bias = 0.0f; break;
    case SELU:        gaussian = true; sigma = 1.0f / _inputLayer._stride; break;
    case Constant:    constant = true; break;
    }
    // every element depends only on its global index, so rows are filled by independent host threads (the 1M x 1,024
    // matrices of BASELINE config 4 are 10^9 Box-Muller draws each)
    auto fillRows = [&](uint64_t r0, uint64_t r1) {
        for (uint64_t r = r0; r < r1; r++) {
            for (uint64_t c = 0; c < _width; c++) {
                const uint64_t g = (uint64_t)(inMin + r) * _outputLayer._stride + (outMin + c);  // global element index
                NNFloat w;
                if (constant) w = -_outputLayer._weightInitScale;      // kScaleAndBias(w = 0, scale 0, bias = scale) = 0 * w - scale (E/kernels.cu:48, E/NNWeight.cpp:550-554)
                else if (gaussian) {
                    const float u1 = u01(key, 2 * g), u2 = u01(key, 2 * g + 1);
                    w = sigma * sqrtf(-2.0f * logf(u1)) * cosf(6.28318530718f * u2);
                } else w = scale * u01(key, g) - bias;
                _vWeight[r * _width + c] = w;
            }
        }
    };
    const uint64_t work = (uint64_t)_height * _width;
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const uint64_t nThreads = work < (1ull << 20) ? 1 : std::min<uint64_t>({(uint64_t)hw, (uint64_t)_height, 64ull});
    if (nThreads <= 1) fillRows(0, _height);
    else {
        std::vector<std::thread> pool;
        for (uint64_t t = 0; t < nThreads; t++)
            pool.emplace_back(fillRows, (uint64_t)_height * t / nThreads, (uint64_t)_height * (t + 1) / nThreads);
        for (auto& th : pool) th.join();
    }
    _pbWeight->Upload(_vWeight.data());
    std::fill(_vBias.begin(), _vBias.end(), _outputLayer._biasInit);                             // E/NNWeight.cpp:556-557
    _pbBias->Upload(_vBias.data());
}

void NNWeight::RefreshState(NNNetwork* pNetwork, TrainingMode mode)
{
    (void)pNetwork;
    // E/NNWeight.cpp:573-602: velocity buffers exist only for the modes that use them
    if (mode != SGD) {
        if (!_pbWeightVelocity) _pbWeightVelocity.reset(new GpuBuffer<NNFloat>(_localSize));
        if (!_pbBiasVelocity) _pbBiasVelocity.reset(new GpuBuffer<NNFloat>(_localBiasSize));
        if (mode == AdaDelta || mode == Adam) {
            if (!_pbWeightGradientVelocity) _pbWeightGradientVelocity.reset(new GpuBuffer<NNFloat>(_localSize));
            if (!_pbBiasGradientVelocity) _pbBiasGradientVelocity.reset(new GpuBuffer<NNFloat>(_localBiasSize));
        } else { _pbWeightGradientVelocity.reset(); _pbBiasGradientVelocity.reset(); }
    } else {
        _pbWeightVelocity.reset(); _pbBiasVelocity.reset(); _pbWeightGradientVelocity.reset(); _pbBiasGradientVelocity.reset();
    }
}

void NNWeight::ClearVelocity()
{
    cudaStream_t s = getGpu().GetStream();
    if (_pbWeightVelocity) cudaMemsetAsync(_pbWeightVelocity->_pDevData, 0, _localSize * sizeof(NNFloat), s);
    if (_pbBiasVelocity) cudaMemsetAsync(_pbBiasVelocity->_pDevData, 0, _localBiasSize * sizeof(NNFloat), s);
    if (_pbWeightGradientVelocity) cudaMemsetAsync(_pbWeightGradientVelocity->_pDevData, 0, _localSize * sizeof(NNFloat), s);
    if (_pbBiasGradientVelocity) cudaMemsetAsync(_pbBiasGradientVelocity->_pDevData, 0, _localBiasSize * sizeof(NNFloat), s);
}

NNFloat NNWeight::CalculateRegularizationError(NNFloat lambda, NNFloat lambda1)
{
    if (_bShared) return 0;
    float e = 0.0f;
    getGpu().Check(dsb200_regularization_error(getGpu()._ctx, lambda, lambda1, _pbWeight->_pDevData, _localSize, &e), "dsb200_regularization_error");
    return e;
}

void NNWeight::UpdateWeights(TrainingMode mode, uint32_t batch, NNFloat alpha, NNFloat lambda, NNFloat lambda1, NNFloat mu, NNFloat mu1, NNFloat t)
{
    if (_bLocked) return;                                                                       // E/NNWeight.cpp:723-724
    dsb200_ctx* ctx = getGpu()._ctx;
    NNFloat* v = _pbWeightVelocity ? _pbWeightVelocity->_pDevData : NULL;
    NNFloat* gv = _pbWeightGradientVelocity ? _pbWeightGradientVelocity->_pDevData : NULL;
    if (_bDeferredDenseGradient) {
        // small dense layer: X^T * delta, the optimizer rule and the bias update in one launch (csrc/dense_small.cu); dW is never written
        _bDeferredDenseGradient = false;
        const NNFloat galpha = -(NNFloat)1.0 / (_sharingCount * (NNFloat)batch);                // E/NNLayer.cpp:2213
        getGpu().Check(dsb200_dense_update(ctx, (int)mode, batch, (uint32_t)_height, (uint32_t)_width, galpha, _pDeferredX, _pDeferredDelta, alpha, lambda, lambda1,
                                           mu, mu1, t, v, gv, _pbWeight->_pDevData, _pbBiasVelocity ? _pbBiasVelocity->_pDevData : NULL,
                                           _pbBiasGradientVelocity ? _pbBiasGradientVelocity->_pDevData : NULL, _pbBias->_pDevData), "dsb200_dense_update");
        return;
    }
    if (_bDeferredSparseGradient) {
        // sparse input gradient + optimizer rule in one kernel; same arithmetic as the two calls below
        const NNFloat galpha = -(NNFloat)1.0 / (_sharingCount * (NNFloat)batch);                // E/NNLayer.cpp:2213
        _inputLayer._pDataSet->CalculateSparseTransposedWeightGradientUpdate(mode, galpha, (uint32_t)_height, (uint32_t)_width, _pDeferredDelta,
                                                                             alpha, lambda, lambda1, mu, mu1, t, v, gv, _pbWeight->_pDevData);
        _bDeferredSparseGradient = false;
    } else {
        getGpu().Check(dsb200_update_weights(ctx, (int)mode, alpha, lambda, lambda1, mu, mu1, t, _localSize, v, _pbWeightGradient->_pDevData, gv,
                                             _pbWeight->_pDevData), "dsb200_update_weights");
    }
    // biases: column mean of the output layer's delta (E/NNWeight.cpp:760-794)
    if (_nBiasPartials > 0) {
        // the fused forward pass of the output layer already reduced delta to a few rows of column sums
        getGpu().Check(dsb200_update_biases_partials(ctx, (int)mode, alpha, mu, mu1, t, batch, (uint32_t)_localBiasSize, _pbBiasPartials->_pDevData, _nBiasPartials,
                                                     _pbBiasVelocity ? _pbBiasVelocity->_pDevData : NULL,
                                                     _pbBiasGradientVelocity ? _pbBiasGradientVelocity->_pDevData : NULL, _pbBias->_pDevData),
                       "dsb200_update_biases_partials");
        _nBiasPartials = 0;
        return;
    }
    getGpu().Check(dsb200_update_biases(ctx, (int)mode, alpha, mu, mu1, t, batch, (uint32_t)_localBiasSize, _outputLayer.GetDeltaBuffer(),
                                        _pbBiasVelocity ? _pbBiasVelocity->_pDevData : NULL,
                                        _pbBiasGradientVelocity ? _pbBiasGradientVelocity->_pDevData : NULL, _pbBias->_pDevData), "dsb200_update_biases");
}

// [4 * ceil(batch / 128)][local bias size] floats for dsb200_gemm_fwd_output_pass
NNFloat* NNWeight::BiasPartialsBuffer(uint32_t batch)
{
    const uint64_t need = (uint64_t)4 * ((batch + 127) / 128) * _localBiasSize;
    if (!_pbBiasPartials || _pbBiasPartials->_length < need) _pbBiasPartials.reset(new GpuBuffer<NNFloat>(need));
    return _pbBiasPartials->_pDevData;
}

bool NNWeight::CopyWeights(const NNWeight* pWeight)
{
    if (!pWeight || pWeight->_localSize != _localSize || pWeight->_localBiasSize != _localBiasSize) return false;
    _pbWeight->Copy(pWeight->_pbWeight->_pDevData);
    _pbBias->Copy(pWeight->_pbBias->_pDevData);
    return true;
}

bool NNWeight::SetWeights(const vector<NNFloat>& vWeight)
{
    if (vWeight.size() < _size) return false;                                                   // E/NNWeight.cpp:946-953
    const uint32_t inMin = _bOutgoingLarger ? 0 : _inputLayer._minX;
    const uint32_t outMin = _bOutgoingLarger ? _outputLayer._minX : 0;
    for (uint64_t r = 0; r < _height; r++)
        for (uint64_t c = 0; c < _width; c++)
            _vWeight[r * _width + c] = vWeight[(uint64_t)(inMin + r) * _outputLayer._stride + (outMin + c)];
    _pbWeight->Upload(_vWeight.data());
    return true;
}

bool NNWeight::SetBiases(const vector<NNFloat>& vBias)
{
    if (vBias.size() < _biasSize) return false;
    for (uint64_t c = 0; c < _localBiasSize; c++) _vBias[c] = vBias[_outputLayer._minX + c];
    _pbBias->Upload(_vBias.data());
    return true;
}

bool NNWeight::GetWeights(vector<NNFloat>& vWeight)
{
    vWeight.resize(_localSize);
    _pbWeight->Download(vWeight.data());
    _vWeight = vWeight;
    return true;
}

bool NNWeight::GetBiases(vector<NNFloat>& vBias)
{
    vBias.resize(_localBiasSize);
    _pbBias->Download(vBias.data());
    _vBias = vBias;
    return true;
}

bool NNWeight::GetGradients(vector<NNFloat>& vGradient)
{
    // with fusion on, the gradient of a sparse-input weight is consumed inside dsb200_sparse_wgrad_update and never written to
    // _pbWeightGradient: returning the buffer would hand out stale data
    if (getGpu()._pNetwork && getGpu()._pNetwork->FusionEnabled() &&
        ((_inputLayer._kind == NNLayer::Kind::Input && _inputLayer._bFastSparse) || NNLayer::SmallDense(_inputLayer._batch, _height, _width)))
        throw DsbEngineError("NNWeight::GetGradients: the gradient of weight " + _inputLayer._name + " -> " + _outputLayer._name +
                             " is fused into the optimizer step and never materialised; call NNNetwork::SetFusion(false) first");
    vGradient.resize(_localSize);
    _pbWeightGradient->Download(vGradient.data());
    return true;
}

bool NNWeight::GetDimensions(vector<uint64_t>& dimensions)
{
    dimensions.clear();
    dimensions.push_back(_width);
    dimensions.push_back(_height);
    return true;
}
