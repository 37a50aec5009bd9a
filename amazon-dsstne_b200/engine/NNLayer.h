// NNLayer.h -- fully-connected layer of the reference (E/NNLayer.h:19-327) on the dsstne_b200
// C ABI.  Same class / descriptor / method names; Convolutional and Pooling layer types, batch
// normalisation and skip connections are outside the hot path and are
// rejected when the network is built.
#pragma once

#include <cmath>
#include "NNTypes.h"

struct NNLayerDescriptor;

class NNLayer {
public:
    friend class NNNetwork;
    friend class NNWeight;
    enum Kind { Input, Hidden, Output, Target };
    enum Type { FullyConnected, Convolutional, Pooling };
    enum Attributes { None = 0x0, Sparse = 0x1, Denoising = 0x2, BatchNormalization = 0x4 };
    enum Parallelization { Data, Model, Serial };

private:
    const string            _name;
    const Kind              _kind;
    const Type              _type;
    const uint32_t          _attributes;
    string                  _dataSet;
    NNDataSetBase*          _pDataSet;
    vector<string>          _vSource;
    uint32_t                _Nx, _Ny, _Nz, _Nw;
    uint32_t                _stride;             // total units
    uint32_t                _localStride;        // units held by this rank (model parallel)
    uint32_t                _maxLocalStride;
    uint32_t                _batch;
    uint32_t                _deltaUpdateCount;
    uint32_t                _unitUpdateCount;
    uint32_t                _dimensions;
    uint32_t                _minX, _maxX;        // [Nx*r/P, Nx*(r+1)/P)  (E/NNLayer.cpp:108-112)
    WeightInitialization    _weightInit;
    NNFloat                 _weightInitScale;
    NNFloat                 _biasInit;
    NNFloat                 _RELUSlope, _ELUAlpha, _SELULambda;
    const Activation        _activation;
    const NNFloat           _pDropout;
    bool                    _bSparse;
    bool                    _bFastSparse;
    NNFloat                 _sparsenessPenalty_p, _sparsenessPenalty_beta;
    const bool              _bDenoising;
    NNFloat                 _weightNorm, _deltaNorm;
    Parallelization         _parallelization;
    bool                    _bDirty;
    // B200 fusion state
    bool                    _bActivationPending;  // unit buffer holds Z; the fused output pass will apply f(), loss, delta
    bool                    _bDeltaReady;         // output delta already produced by the fused pass
    bool                    _bHadamardDone;       // f'(x) already applied by the fused input-delta kernel of the layer above
    bool                    _bUnitsArePreActivation;  // the fused output pass did not store a = f(z): the unit buffer still holds z
    uint32_t                _preActivationBatch;
    bool                    _bForwardDeferred = false;    // engine option "fuse_output_gemm": the forward GEMM of this (output) layer has not run; the
                                                          // loss / delta pass runs it fused, or RunDeferredForward() runs it when the units are needed
    const NNFloat*          _pDeferredA = NULL;           // input of the deferred forward GEMM: the units of the layer below (all of them:
    uint32_t                _deferredK = 0;               // the gathered copy when model parallel) and their count
    void                    RunDeferredForward(bool applyActivation);
    // model-parallel exchange state (B200: one peer-memory kernel or one NCCL call per exchange, see Reduce / Gather)
    enum ExchangeSlot { UnitsGather = 0, UnitsReduce = 1, DeltaReduce = 2, DeltaGather = 3 };
    uint32_t                _exchangeIndex = 0;           // position in NNNetwork::_vLayer: slot = 4 * index + ExchangeSlot
    NNFloat*                _pGatheredUnits = NULL;       // all units of this layer on this rank, valid while _bUnitsGathered (one step)
    bool                    _bUnitsGathered = false;
    bool                    _bBiasActDone = false;        // the layer below already applied bias (+ activation) in its GEMM epilogue
    unique_ptr<GpuBuffer<NNFloat>> _pbGatheredUnits, _pbGatheredDelta;   // NCCL path only: the peer-memory path gathers into its arena slots

    vector<NNLayer*>        _vIncomingLayer;
    vector<NNWeight*>       _vIncomingWeight;
    vector<NNLayer*>        _vOutgoingLayer;
    vector<NNWeight*>       _vOutgoingWeight;
    vector<NNLayer*>        _vIncomingLargerLayer;
    vector<NNWeight*>       _vIncomingLargerWeight;
    vector<NNLayer*>        _vOutgoingLargerLayer;
    vector<NNWeight*>       _vOutgoingLargerWeight;
    vector<NNFloat>         _vUnit, _vDelta;
    unique_ptr<GpuBuffer<NNFloat>> _pbUnit;
    unique_ptr<GpuBuffer<NNFloat>> _pbDelta;
    unique_ptr<GpuBuffer<NNFloat>> _pbDropout;
    int32_t                 _priority;
    uint64_t                _dropoutCalls;        // masks drawn so far (stream id of the counter-based generator)

    NNLayer(NNLayerDescriptor& l, uint32_t batch);
    ~NNLayer();
    void Allocate(bool validate);
    void Deallocate();
    void SetBatch(uint32_t batch);
    void RefreshParallelization();
    void RefreshState(NNNetwork* pNetwork, TrainingMode trainingMode, bool validate);
    void LoadPredictionBatch(uint32_t position, uint32_t batch);
    void LoadTrainingBatch(uint32_t position, uint32_t batch);
    void LoadValidationBatch(uint32_t position, uint32_t batch);
    void ForwardPropagate(uint32_t position, uint32_t batch, bool bTraining = false);
    void ForwardPropagateFullyConnected(uint32_t position, uint32_t batch, bool bTraining);
    void CalculateActivation(uint32_t batch);
    void CalculateDropout(uint32_t batch);
    NNFloat CalculateError(uint32_t position, uint32_t batch, ErrorFunction ef);
    bool CalculateErrorAsync(uint32_t position, uint32_t batch, ErrorFunction ef, unsigned long long* pDevAccumulator);
    void BackPropagate(uint32_t position, uint32_t batch);
    void BackPropagateFullyConnected(uint32_t position, uint32_t batch);
    void CalculateOutputDelta(uint32_t position, uint32_t batch, ErrorFunction ef);
    void GenerateDenoisingData();
    void Reduce(uint32_t batch, uint32_t stride, NNFloat* pBuffer, uint32_t localStride, uint32_t updateCount, ExchangeSlot slot,
                const NNFloat* pBias = NULL, Activation activation = Linear);
    NNFloat* Gather(uint32_t batch, uint32_t stride, NNFloat* pBuffer, uint32_t localStride, ExchangeSlot slot);
    void ClearUpdates();
    bool FusedOutputEligible(ErrorFunction ef) const;
    // a dense weight small enough for the one-launch gradient + update kernel (dsb200_dense_update)
    static bool SmallDense(uint64_t batch, uint64_t k, uint64_t n) { return batch * k * n <= (1ull << 27) && batch <= 4096; }
    void MaterializeUnits();                              // apply the activation now if the fused training pass skipped storing it
    NNFloat* GetIncomingUnitBuffer() { return _pbUnit ? _pbUnit->_pDevData : NULL; }
    NNFloat* GetUnitBuffer() { return _pbUnit ? _pbUnit->_pDevData : NULL; }
    NNFloat* GetIncomingDeltaBuffer() { return _pbDelta ? _pbDelta->_pDevData : NULL; }
    NNFloat* GetDeltaBuffer() { return _pbDelta ? _pbDelta->_pDevData : NULL; }
    uint64_t GetBufferSize() { return (uint64_t)_batch * _stride; }

public:
    const string& GetName() const { return _name; }
    const string& GetDataSetName() const { return _dataSet; }
    Kind GetKind() const { return _kind; }
    Type GetType() const { return _type; }
    uint32_t GetAttributes() const { return _attributes; }
    NNDataSetBase* GetDataSet() const { return _pDataSet; }
    uint32_t GetNumDimensions() const { return _dimensions; }
    tuple<uint32_t, uint32_t, uint32_t, uint32_t> GetDimensions() const { return std::make_tuple(_Nx, _Ny, _Nz, _Nw); }
    tuple<uint32_t, uint32_t, uint32_t, uint32_t> GetLocalDimensions() const { return std::make_tuple(_maxX - _minX, _Ny, _Nz, _Nw); }
    uint32_t GetLocalStride() const { return _localStride; }
    bool GetUnits(vector<NNFloat>& vUnit);
    bool GetUnits(NNFloat* pUnit);
    bool SetUnits(const vector<NNFloat>& vUnit);
    bool GetDeltas(vector<NNFloat>& vDelta);
    bool GetDeltas(NNFloat* pDelta);
    bool SetDeltas(const vector<NNFloat>& vDelta);
};

struct NNLayerDescriptor {
    string                  _name;
    NNLayer::Kind           _kind;
    NNLayer::Type           _type;
    PoolingFunction         _poolingFunction;
    string                  _dataSet;
    vector<string>          _vSource;
    vector<string>          _vSkip;
    uint32_t                _Nx, _Ny, _Nz, _Nw;
    uint32_t                _dimensions;
    bool                    _bDimensionsProvided;
    WeightInitialization    _weightInit;
    NNFloat                 _weightInitScale;
    NNFloat                 _biasInit;
    NNFloat                 _weightNorm, _deltaNorm;
    NNFloat                 _pDropout;
    Activation              _activation;
    NNFloat                 _sparsenessPenalty_p, _sparsenessPenalty_beta;
    uint32_t                _attributes;
    NNFloat                 _RELUSlope, _ELUAlpha, _SELULambda;
    NNLayerDescriptor();
};
