// NNNetwork.h -- network graph + training / prediction loops of the reference (E/NNNetwork.h:20-293)
// for fully-connected networks with sparse inputs / sparse targets, on the dsstne_b200 C ABI.
//
// Same class, descriptor, method and loader names as the reference so callers (the `train` /
// `predict` drivers, language bindings) switch by relinking.  Not carried over: convolution /
// pooling layers, LRN / maxout, shared weights, ImportAutoEncoder (outside the hot path).
#pragma once

#include <map>

#include "NNLayer.h"
#include "NNTypes.h"
#include "NNWeight.h"

struct NNNetworkDescriptor;

class NNNetwork {
public:
    friend class NNLayer;
    friend class NNWeight;
    friend struct GpuContext;
    enum Kind { FeedForward, AutoEncoder };

private:
    friend NNNetwork* LoadNeuralNetworkJSON(const string& fname, const uint32_t batch, const vector<NNDataSetBase*>& vDataSet);
    friend NNNetwork* LoadNeuralNetworkJSONString(const string& json, const uint32_t batch, const vector<NNDataSetBase*>& vDataSet);
    friend NNNetwork* LoadNeuralNetworkNetCDF(const string& fname, const uint32_t batch);
    friend NNNetwork* CreateNeuralNetwork(NNNetworkDescriptor& nd, uint32_t batch);
    string                  _name;
    uint32_t                _batch;
    uint32_t                _position;
    bool                    _bExamplesFound;
    bool                    _bAllDataLoaded;
    uint32_t                _examples;
    const Kind              _kind;
    ErrorFunction           _errorFunction;
    TrainingMode            _trainingMode;
    Mode                    _mode;
    uint32_t                _epochs;
    uint32_t                _batches;
    float                   _decay;
    NNFloat                 _RELUSlope, _ELUAlpha, _SELULambda;
    bool                    _bSparsenessPenalty;
    NNFloat                 _sparsenessPenalty_p, _sparsenessPenalty_beta;
    bool                    _bDenoising;
    NNFloat                 _denoising_p;
    NNFloat                 _deltaBoost_one, _deltaBoost_zero;
    NNFloat                 _SMCE_oneTarget, _SMCE_zeroTarget, _SMCE_oneScale, _SMCE_zeroScale;
    bool                    _bShuffleIndices;
    uint32_t                _shuffleIndices;
    unique_ptr<GpuBuffer<uint32_t>> _pbShuffleIndex;
    vector<uint32_t>        _vShuffleIndex;
    uint64_t                _shuffleEpoch;
    string                  _checkpoint_name;
    int32_t                 _checkpoint_interval, _checkpoint_epochs;
    vector<NNLayer*>        _vLayer, _vInputLayer, _vOutputLayer;
    vector<NNWeight*>       _vWeight;
    vector<NNDataSetBase*>  _vData;
    vector<NNLayer*>        _vFPOrder, _vBPOrder;
    std::map<string, NNLayer*> _mLayer;
    bool                    _bDirty;
    bool                    _bClearVelocity;
    size_t                  _scratchBufferSize;
    unique_ptr<GpuBuffer<NNFloat>> _pbScratchBuffer;
    uint32_t                _maxStride;
    unique_ptr<GpuBuffer<NNFloat>> _pbP2PBuffer;          // full-width exchange buffer ("P2P send buffer")
    unique_ptr<GpuBuffer<unsigned long long>> _pbErrorAccumulator;   // device fixed-point loss + pinned shadow
    cudaEvent_t             _errorEvent;
    // engine option "step_trace" (diagnostics): where a TrainStep spends host and device time (dsb200_engine_step_trace)
    struct StepTrace {
        static const int kRing = 64;
        bool on = false;
        cudaEvent_t start[kRing] = {}, loss[kRing] = {}, end[kRing] = {};
        uint64_t steps = 0;
        double host[6] = {0, 0, 0, 0, 0, 0};           // seconds: preparation launches, forward launches, loss launches, backward launches, wait for the loss, update launches
    } _trace;
    cudaStream_t            _sideStream;                  // regularisation error runs here, beside the forward pass
    cudaEvent_t             _forkEvent, _joinEvent, _prepEvent, _regEvent = NULL, _lossPassEvent = NULL;
    bool                    _bTransposePending = false;   // the side stream is still building this step's transposed matrix (_joinEvent): WaitForTransposed()
    cudaEvent_t             _updateEvent = NULL;          // end of the weight updates that ran on the side stream (UpdateWeights)
    size_t                  _validateMaxSamples = 256;    // Validate(): elements checked per weight matrix / bias vector
    bool                    _bStepReadsRecorded = false;   // TrainStep: LaunchError records GpuContext::_dataConsumedEvent for the streaming loader
    bool                    _bBatchPrepared = false;      // TrainStep launched LoadBatch on the side stream: PredictTrainingBatch must not repeat it
    bool                    _verbose;
    bool                    _bRegularizationLaunched;     // LaunchError finds the regularisation kernels already in flight
    bool                    _bFusion;                     // B200 fusions on (default) / off (kernel-by-kernel, like the reference)
    bool                    _bPeerMemoryExchange = false; // model parallel: the peer-memory arena is mapped on every rank (AllocatePeerBuffers)
    uint64_t                _peerSlotFloats = 0;
    // divergence-brake state that survives across Train calls made one step at a time
    NNFloat                 _movingAverage;
    uint32_t                _brakeSteps, _initSteps;

public:
    ~NNNetwork();
    void ClearDataSets();
    void LoadDataSets(vector<NNDataSetBase*>& vData);
    void Randomize();
    bool Validate();
    void SetValidateSamples(size_t n) { _validateMaxSamples = n ? n : 1; }
    float Train(uint32_t epochs = 1, NNFloat alpha = (NNFloat)0.1, NNFloat lambda = (NNFloat)0.001, NNFloat lambda1 = (NNFloat)0.0,
                NNFloat mu = (NNFloat)0.1, NNFloat mu1 = 0.0);
    // one minibatch of Train's loop body at `position` (B200 addition: lets a caller time / drive single steps)
    const std::vector<uint32_t>& ShuffleIndexVector() const { return _vShuffleIndex; }
    void WaitForTransposed();                             // main stream waits for the side stream's transposed matrix, once per step, at its first consumer
    void SetStepTrace(bool on);
    int StepTraceReport(double* out, int cap);            // 6 host means (us) + 3 device means (us): start->loss, loss->end, end->next start
    float TrainStep(uint32_t position, NNFloat alpha, NNFloat lambda, NNFloat lambda1, NNFloat mu, NNFloat mu1, NNFloat* pRegularization = NULL);
    void PredictBatch(uint32_t layers = 0);
    void CalculateTopK(const string& layer, uint32_t k, GpuBuffer<NNFloat>* pbKey, GpuBuffer<uint32_t>* pbValue);
    // top-K with the exclusion filter applied on the device (replaces the host round trip of U/NNRecsGenerator.cpp:132-150)
    void CalculateTopKFiltered(const string& layer, uint32_t k, NNDataSetBase* pFilter, GpuBuffer<NNFloat>* pbKey, GpuBuffer<uint32_t>* pbValue);
    // model parallel: every rank's local top-K (local ids + this rank's first unit) is all-gathered and merged, so that every rank
    // holds the top-K of the WHOLE layer with global ids -- the role of the MPI gather + host merge of U/NNRecsGenerator.cpp:150-244
    void CalculateTopKGlobal(const string& layer, uint32_t k, NNDataSetBase* pFilter, GpuBuffer<NNFloat>* pbKey, GpuBuffer<uint32_t>* pbValue);
    bool LockWeights(const string& inputLayer, const string& outputLayer);
    bool UnlockWeights(const string& inputLayer, const string& outputLayer);
    void SetBatch(uint32_t batch);
    void SetPosition(uint32_t position);
    bool SetDecay(NNFloat decay);
    void SetTrainingMode(TrainingMode mode);
    void SetShuffleIndices(bool bShuffleIndices);
    void SetClearVelocity(bool bClear) { _bClearVelocity = bClear; }
    void SetFusion(bool bFusion) { _bFusion = bFusion; }
    bool FusionEnabled() const { return _bFusion; }
    void MarkDirty() { _bDirty = true; }
    bool PeerMemoryExchange() const { return _bPeerMemoryExchange; }     // exchange steps on the dsb200_p2p_* kernels (else NCCL)
    bool SaveNetCDF(const string& fname);
    unsigned int GetBatch() const { return _batch; }
    uint32_t GetExamples() const { return _examples; }
    uint32_t GetPosition() const { return _position; }
    ErrorFunction GetErrorFunction() const { return _errorFunction; }
    NNWeight* GetWeight(const string& inputLayer, const string& outputLayer) const;
    uint64_t GetBufferSize(const string& layer) const;
    NNLayer* GetLayer(const string& layer) const;
    vector<string> GetLayers() const;
    const string& GetName() const { return _name; }
    tuple<NNFloat> GetDecay() const { return std::make_tuple(_decay); }
    tuple<NNFloat, NNFloat> GetSparsenessPenalty() const { return std::make_tuple(_sparsenessPenalty_p, _sparsenessPenalty_beta); }
    tuple<NNFloat> GetDenoising() const { return std::make_tuple(_denoising_p); }
    tuple<NNFloat, NNFloat> GetDeltaBoost() const { return std::make_tuple(_deltaBoost_one, _deltaBoost_zero); }
    tuple<NNFloat, NNFloat, NNFloat, NNFloat> GetSMCE() const { return std::make_tuple(_SMCE_oneTarget, _SMCE_zeroTarget, _SMCE_oneScale, _SMCE_zeroScale); }
    tuple<bool> GetShuffleIndices() const { return std::make_tuple(_bShuffleIndices); }
    tuple<string, int32_t> GetCheckPoint() const { return std::make_tuple(_checkpoint_name, _checkpoint_interval); }
    bool GetDebugLevel() const { return _verbose; }
    NNFloat* GetUnitBuffer(const string& layer);
    NNFloat* GetDeltaBuffer(const string& layer);
    NNFloat* GetWeightBuffer(const string& inputLayer, const string& outputLayer);
    NNFloat* GetScratchBuffer(size_t size = 0);
    NNFloat* GetP2PSendBuffer();
    bool P2P_Allreduce(NNFloat* pBuffer, size_t size);
    bool SetSparsenessPenalty(NNFloat p = 0.0f, NNFloat beta = 0.0f);
    bool SetDenoising(NNFloat p = 0.0f);
    bool SetDeltaBoost(NNFloat one = 1.0f, NNFloat zero = 1.0f);
    bool SetSMCE(NNFloat oneTarget = 0.9f, NNFloat zeroTarget = 0.1f, NNFloat oneScale = 1.0f, NNFloat zeroScale = 1.0f);
    bool SetCheckpoint(string name, int32_t interval);
    void SetDebugLevel(bool verbose) { _verbose = verbose; }
    unsigned long long* GetErrorAccumulator() { return _pbErrorAccumulator->_pDevData; }
    NNFloat ReadErrorAccumulator();
    void AddBuffers(NNFloat* pDst, NNFloat* pSrc, uint64_t size);

private:
    void CalculatePropagationOrder();
    void AllocatePeerBuffers();
    void LoadBatch();
    void PredictTrainingBatch(uint32_t layers = 0);
    void RefreshShuffleBuffers();
    void ShuffleIndices();
    tuple<NNFloat, NNFloat> CalculateError(NNFloat lambda, NNFloat lambda1);
    void LaunchError(NNFloat lambda, NNFloat lambda1);     // asynchronous part of CalculateError
    void LaunchBatchPreparation(NNFloat lambda, NNFloat lambda1);
    void LaunchRegularization(NNFloat lambda, NNFloat lambda1, bool fork);
    void ClearUpdates();
    void BackPropagate();
    void UpdateWeights(NNFloat alpha, NNFloat lambda, NNFloat lambda1, NNFloat mu, NNFloat mu1);
    NNNetwork(NNNetworkDescriptor& nd, uint32_t batch = DefaultBatch);
    void RefreshState();
};

struct NNNetworkDescriptor {
    string                      _name;
    NNNetwork::Kind             _kind;
    ErrorFunction               _errorFunction;
    vector<NNLayerDescriptor>   _vLayerDescriptor;
    vector<NNWeightDescriptor>  _vWeightDescriptor;
    bool                        _bShuffleIndices;
    NNFloat                     _decay;
    NNFloat                     _RELUSlope, _ELUAlpha, _SELULambda;
    bool                        _bSparsenessPenalty;
    NNFloat                     _sparsenessPenalty_p, _sparsenessPenalty_beta;
    bool                        _bDenoising;
    NNFloat                     _denoising_p;
    NNFloat                     _deltaBoost_one, _deltaBoost_zero;
    NNFloat                     _SMCE_oneTarget, _SMCE_zeroTarget, _SMCE_oneScale, _SMCE_zeroScale;
    string                      _checkpoint_name;
    int32_t                     _checkpoint_interval, _checkpoint_epochs;
    NNNetworkDescriptor();
};

// name and dimensions of a data set: all the JSON loader needs from it (auto-sized layers)
struct NNDataSetShape { string _name; uint32_t _width, _height, _length, _dimensions; };
// host-only stages of LoadNeuralNetworkJSON (no GPU context needed)
NNNetworkDescriptor ParseNeuralNetworkJSON(const string& json, const vector<NNDataSetShape>& vDataSet);
string DescribeNeuralNetworkJSON(const string& json, const vector<NNDataSetShape>& vDataSet);
NNNetwork* CreateNeuralNetwork(NNNetworkDescriptor& nd, uint32_t batch = DefaultBatch);
NNNetwork* LoadNeuralNetworkNetCDF(const string& fname, const uint32_t batch = DefaultBatch);
NNNetwork* LoadNeuralNetworkJSON(const string& fname, const uint32_t batch = DefaultBatch,
                                 const vector<NNDataSetBase*>& vDataSet = vector<NNDataSetBase*>());
NNNetwork* LoadNeuralNetworkJSONString(const string& json, const uint32_t batch = DefaultBatch,
                                       const vector<NNDataSetBase*>& vDataSet = vector<NNDataSetBase*>());
