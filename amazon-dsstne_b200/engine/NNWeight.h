// NNWeight.h -- weight matrix between two fully-connected layers (E/NNWeight.h:18-124).
// Linear transform only; Convolution, shared and transposed (tied) weights are outside the hot
// path and rejected at construction.
#pragma once

#include "NNTypes.h"

class NNWeight {
public:
    enum Transform { Convolution, Linear };

private:
    friend class NNNetwork;
    friend class NNLayer;
    NNLayer&    _inputLayer;
    NNLayer&    _outputLayer;
    const bool  _bShared;
    const bool  _bTransposed;
    Transform   _transform;
    bool        _bLocked;
    NNWeight*   _pSharedWeight;
    uint32_t    _sharingCount;
    uint32_t    _updateCount;
    uint64_t    _width;            // output units held here (all, or the local slice when "outgoing larger")
    uint64_t    _height;           // input units held here (all, or the local slice when "incoming larger")
    uint64_t    _size, _biasSize, _localSize, _localBiasSize;
    bool        _bOutgoingLarger;  // sharding rule outcome (E/NNWeight.cpp:435-457)
    NNFloat     _norm;
    bool        _bDeferredSparseGradient;   // gradient will be produced inside the fused update
    NNFloat*    _pDeferredDelta;            // delta [batch][outputStride] the fused update reads
    bool        _bDeferredDenseGradient = false;   // small dense layer: gradient + optimizer + bias update run as ONE kernel in UpdateWeights
    const NNFloat* _pDeferredX = NULL;             // its input units [batch][_height]
    uint32_t    _nBiasPartials;             // > 0: the fused output-layer forward pass left this many rows of column sums of delta
    // true when UpdateWeights of this weight launches only kernels that use no scratch of the context (safe beside another update)
    bool UpdateTouchesNoScratch() const { return !_bLocked && !_bDeferredSparseGradient && (_bDeferredDenseGradient || _nBiasPartials > 0); }
    unique_ptr<GpuBuffer<NNFloat>> _pbBiasPartials;
    vector<NNFloat> _vWeight, _vBias;
    unique_ptr<GpuBuffer<NNFloat>> _pbWeight, _pbBias, _pbWeightGradient;
    unique_ptr<GpuBuffer<NNFloat>> _pbWeightVelocity, _pbBiasVelocity, _pbWeightGradientVelocity, _pbBiasGradientVelocity;

    NNWeight(NNLayer& inputLayer, NNLayer& outputLayer, bool bShared = false, bool bTransposed = false, bool bLocked = false, NNFloat maxNorm = 0.0f);
    ~NNWeight();
    NNFloat CalculateRegularizationError(NNFloat lambda, NNFloat lambda1);
    void ClearVelocity();
    void Randomize();
    void Lock() { _bLocked = true; }
    void Unlock() { _bLocked = false; }
    void RefreshState(NNNetwork* pNetwork, TrainingMode trainingMode);
    void UpdateWeights(TrainingMode trainingMode, uint32_t batch, NNFloat alpha, NNFloat lambda, NNFloat lambda1, NNFloat mu, NNFloat mu1, NNFloat t);
    NNFloat* GetWeightBuffer() { return _pbWeight ? _pbWeight->_pDevData : NULL; }
    NNFloat* GetWeightGradientBuffer() { return _pbWeightGradient ? _pbWeightGradient->_pDevData : NULL; }
    uint64_t GetBufferSize() { return _localSize; }
    NNFloat* BiasPartialsBuffer(uint32_t batch);

public:
    bool CopyWeights(const NNWeight* pWeight);
    // full (un-sharded) row-major [inputStride][outputStride] vectors; each rank keeps its slice
    bool SetWeights(const vector<NNFloat>& vWeight);
    bool SetBiases(const vector<NNFloat>& vBias);
    // local shard as held on this rank (full matrix when running on one GPU)
    bool GetWeights(vector<NNFloat>& vWeight);
    bool GetBiases(vector<NNFloat>& vBias);
    bool GetGradients(vector<NNFloat>& vGradient);
    bool GetDimensions(vector<uint64_t>& dimensions);
    bool SetNorm(NNFloat norm) { _norm = norm; return true; }
    bool IsOutgoingLarger() const { return _bOutgoingLarger; }
};

struct NNWeightDescriptor {
    string          _inputLayer, _outputLayer;
    uint64_t        _width, _height, _length, _depth, _breadth;
    vector<NNFloat> _vWeight, _vBias;
    bool            _bShared, _bTransposed, _bLocked;
    NNFloat         _norm;
    string          _sourceInputLayer, _sourceOutputLayer;
    NNWeightDescriptor() : _width(1), _height(1), _length(1), _depth(1), _breadth(1), _bShared(false), _bTransposed(false), _bLocked(false), _norm(0.0f) {}
};
