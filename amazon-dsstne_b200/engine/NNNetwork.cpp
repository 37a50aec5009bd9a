// NNNetwork.cpp -- graph, training and prediction loops for fully-connected networks.
//
// Follows E/NNNetwork.cpp: constructor (:404-700), LoadDataSets (:1001-1170), RefreshState (:909-1000),
// PredictBatch / PredictTrainingBatch (:1376-1470), Train (:1536-1694), CalculateError (:1708-1744),
// BackPropagate (:1747-1768), UpdateWeights (:1770-1790), CalculateTopK (:1792-1822).
// B200-first differences:
//  * the per-minibatch loss is a fixed-point device accumulator copied to pinned memory behind an
//    event; the host launches BackPropagate BEFORE it waits for the value, so the GPU never idles on
//    the read-back (the reference blocks in a synchronous Download every batch, E/kLoss.cu:2349);
//    skipping BackPropagate while the divergence brake skips the update (E/NNNetwork.cpp:1641-1650) has
//    no observable effect, so it always runs and only the update is conditional;
//  * model parallelism: per-layer exchange through NCCL (NNLayer::Reduce / Gather), the two loss
//    scalars through an int64 all-reduce of the fixed-point words instead of MPI_Allreduce on doubles;
//  * shuffle indices come from a host Fisher-Yates driven by the counter-based generator (identical
//    on every rank, so no broadcast).
#include "NNNetwork.h"

#include <algorithm>
#include <chrono>
#include <cfloat>
#include <iostream>
#include <sstream>

using namespace std;

NNNetworkDescriptor::NNNetworkDescriptor()
    : _kind(NNNetwork::Kind::FeedForward), _errorFunction(ErrorFunction::CrossEntropy), _bShuffleIndices(true), _decay(0.0f),
      _RELUSlope(1.0f), _ELUAlpha(1.0f), _SELULambda(1.050701f), _bSparsenessPenalty(false), _sparsenessPenalty_p(0.0f),
      _sparsenessPenalty_beta(0.0f), _bDenoising(false), _denoising_p(0.0f), _deltaBoost_one(1.0f), _deltaBoost_zero(1.0f),
      _SMCE_oneTarget(0.9f), _SMCE_zeroTarget(0.1f), _SMCE_oneScale(1.0f), _SMCE_zeroScale(1.0f), _checkpoint_name("checkpoint"),
      _checkpoint_interval(0), _checkpoint_epochs(0)
{
}

NNNetwork* CreateNeuralNetwork(NNNetworkDescriptor& nd, uint32_t batch) { return new NNNetwork(nd, batch); }

NNNetwork::NNNetwork(NNNetworkDescriptor& d, uint32_t batch)
    : _name(d._name), _batch(batch), _position(0), _bExamplesFound(false), _bAllDataLoaded(true), _examples(0), _kind(d._kind),
      _errorFunction(d._errorFunction), _trainingMode(SGD), _mode(Prediction), _epochs(0), _batches(0), _decay(d._decay),
      _RELUSlope(d._RELUSlope), _ELUAlpha(d._ELUAlpha), _SELULambda(d._SELULambda), _bSparsenessPenalty(d._bSparsenessPenalty),
      _sparsenessPenalty_p(d._sparsenessPenalty_p), _sparsenessPenalty_beta(d._sparsenessPenalty_beta), _bDenoising(d._bDenoising),
      _denoising_p(d._denoising_p), _deltaBoost_one(d._deltaBoost_one), _deltaBoost_zero(d._deltaBoost_zero),
      _SMCE_oneTarget(d._SMCE_oneTarget), _SMCE_zeroTarget(d._SMCE_zeroTarget), _SMCE_oneScale(d._SMCE_oneScale),
      _SMCE_zeroScale(d._SMCE_zeroScale), _bShuffleIndices(d._bShuffleIndices), _shuffleIndices(0), _shuffleEpoch(0),
      _checkpoint_name(d._checkpoint_name), _checkpoint_interval(d._checkpoint_interval), _checkpoint_epochs(0), _bDirty(true),
      _bClearVelocity(true), _scratchBufferSize(0), _maxStride(0), _errorEvent(NULL), _sideStream(NULL), _forkEvent(NULL), _joinEvent(NULL), _prepEvent(NULL), _verbose(false), _bRegularizationLaunched(false), _bFusion(true),
      _movingAverage(0.0f), _brakeSteps(0), _initSteps(100)
{
    if (!getGpu()._ctx) throw DsbEngineError("NNNetwork: getGpu().Startup() has not been called (no GPU context; there is no CPU fallback)");
    for (auto l : d._vLayerDescriptor) {
        if (std::isnan(l._RELUSlope)) l._RELUSlope = _RELUSlope;            // layer inherits the network defaults (E/NNNetwork.cpp:3683-3690)
        if (std::isnan(l._ELUAlpha)) l._ELUAlpha = _ELUAlpha;
        if (std::isnan(l._SELULambda)) l._SELULambda = _SELULambda;
        if (_mLayer.count(l._name)) throw DsbEngineError("NNNetwork: duplicate layer name " + l._name);
        _vLayer.push_back(new NNLayer(l, batch));
        _mLayer[_vLayer.back()->_name] = _vLayer.back();
        if (_vLayer.back()->_kind == NNLayer::Kind::Input) _vInputLayer.push_back(_vLayer.back());
        else if (_vLayer.back()->_kind == NNLayer::Kind::Output) _vOutputLayer.push_back(_vLayer.back());
    }
    for (auto& wd : d._vWeightDescriptor) {                                  // E/NNNetwork.cpp:540-600
        if (!_mLayer.count(wd._inputLayer) || !_mLayer.count(wd._outputLayer))
            throw DsbEngineError("NNNetwork: weight between unknown layers " + wd._inputLayer + " -> " + wd._outputLayer);
        NNWeight* pWeight = new NNWeight(*_mLayer[wd._inputLayer], *_mLayer[wd._outputLayer], wd._bShared, wd._bTransposed, wd._bLocked, wd._norm);
        _vWeight.push_back(pWeight);
        if (wd._vWeight.empty() || wd._vBias.empty()) pWeight->Randomize();
        if (!wd._vWeight.empty()) pWeight->SetWeights(wd._vWeight);
        if (!wd._vBias.empty()) pWeight->SetBiases(wd._vBias);
    }
    CalculatePropagationOrder();
    _pbErrorAccumulator.reset(new GpuBuffer<unsigned long long>(4, true));
    RTERROR(cudaEventCreateWithFlags(&_errorEvent, cudaEventDisableTiming), "NNNetwork: cudaEventCreate");
    RTERROR(cudaEventCreateWithFlags(&_forkEvent, cudaEventDisableTiming), "NNNetwork: cudaEventCreate");
    RTERROR(cudaEventCreateWithFlags(&_joinEvent, cudaEventDisableTiming), "NNNetwork: cudaEventCreate");
    RTERROR(cudaEventCreateWithFlags(&_prepEvent, cudaEventDisableTiming), "NNNetwork: cudaEventCreate");
    RTERROR(cudaEventCreateWithFlags(&_regEvent, cudaEventDisableTiming), "NNNetwork: cudaEventCreate");
    RTERROR(cudaEventCreateWithFlags(&_lossPassEvent, cudaEventDisableTiming), "NNNetwork: cudaEventCreate");
    RTERROR(cudaEventCreateWithFlags(&_updateEvent, cudaEventDisableTiming), "NNNetwork: cudaEventCreate");
    RTERROR(cudaStreamCreateWithFlags(&_sideStream, cudaStreamNonBlocking), "NNNetwork: cudaStreamCreate");
}

NNNetwork::~NNNetwork()
{
    if (getGpu()._pNetwork == this) getGpu()._pNetwork = NULL;
    for (auto w : _vWeight) delete w;
    for (auto l : _vLayer) delete l;
    if (_errorEvent) cudaEventDestroy(_errorEvent);
    if (_forkEvent) cudaEventDestroy(_forkEvent);
    if (_joinEvent) cudaEventDestroy(_joinEvent);
    if (_prepEvent) cudaEventDestroy(_prepEvent);
    if (_regEvent) cudaEventDestroy(_regEvent);
    if (_lossPassEvent) cudaEventDestroy(_lossPassEvent);
    if (_updateEvent) cudaEventDestroy(_updateEvent);
    if (_sideStream) cudaStreamDestroy(_sideStream);
}

// Kahn's algorithm over the layer graph, ties broken by declaration order (the reference assigns
// priorities by longest path, E/NNNetwork.cpp:2345-2457; for feed-forward graphs the orders agree)
void NNNetwork::CalculatePropagationOrder()
{
    _vFPOrder.clear(); _vBPOrder.clear();
    map<NNLayer*, size_t> pending;
    for (auto l : _vLayer) pending[l] = l->_vIncomingLayer.size();
    vector<NNLayer*> done;
    while (done.size() < _vLayer.size()) {
        bool progressed = false;
        for (auto l : _vLayer) {
            if (l->_priority >= 0 || pending[l] != 0) continue;
            l->_priority = (int32_t)done.size();
            done.push_back(l);
            for (auto o : l->_vOutgoingLayer) pending[o]--;
            progressed = true;
        }
        if (!progressed) throw DsbEngineError("NNNetwork: the layer graph has a cycle");
    }
    _vFPOrder = done;
    _vBPOrder.assign(done.rbegin(), done.rend());
}

void GpuContext::SetNeuralNetwork(NNNetwork* pNetwork)
{
    // E/GpuTypes.cpp:475-498: publish the network's hyper-parameters to the kernels
    _pNetwork = pNetwork;
    _data.bShuffleIndices = pNetwork->_bShuffleIndices && (pNetwork->_mode == Training);
    _data.pShuffleIndex = pNetwork->_pbShuffleIndex ? pNetwork->_pbShuffleIndex->_pDevData : NULL;
    _data.denoising_p = pNetwork->_denoising_p;
    _data.denoising_q = 1.0f / (1.0f - pNetwork->_denoising_p);
    _data.deltaBoost_one = pNetwork->_deltaBoost_one;
    _data.deltaBoost_zero = pNetwork->_deltaBoost_zero;
    _data.SMCE_oneTarget = pNetwork->_SMCE_oneTarget;
    _data.SMCE_zeroTarget = pNetwork->_SMCE_zeroTarget;
    _data.SMCE_oneScale = pNetwork->_SMCE_oneScale;
    _data.SMCE_zeroScale = pNetwork->_SMCE_zeroScale;
    CopyConstants();
}

void NNNetwork::ClearDataSets()
{
    _examples = 0; _bExamplesFound = false;
    for (auto l : _vInputLayer) l->_pDataSet = NULL;
    for (auto l : _vOutputLayer) l->_pDataSet = NULL;
    _bDirty = true;
}

void NNNetwork::LoadDataSets(vector<NNDataSetBase*>& vData)
{
    _bAllDataLoaded = false;
    auto attach = [&](NNLayer* l, bool isInput) {
        for (auto d : vData) {
            if (l->_dataSet.compare(d->_name) != 0) continue;
            if (l->_Nx < d->_width || l->_Ny < d->_height || l->_Nz < d->_length) {
                stringstream msg;
                msg << "NNNetwork::LoadDataSets: Data element mismatch (" << l->_Nx << ", " << l->_Ny << ", " << l->_Nz << ") layer " << l->_name
                    << " versus (" << d->_width << ", " << d->_height << ", " << d->_length << ") data set " << d->_name;
                throw DsbEngineError(msg.str());
            }
            if (!_bExamplesFound) { _examples = d->_examples; _bExamplesFound = true; }
            if (d->_examples != _examples) throw DsbEngineError("NNNetwork::LoadDataSets: Mismatched examples count in dataset " + d->_name);
            l->_pDataSet = d;
            if (isInput) l->_bSparse = d->_attributes & NNDataSetEnums::Sparse;      // input layer takes the sparseness of its data
            l->_bDirty = true;
            break;
        }
    };
    for (auto l : _vInputLayer) attach(l, true);
    for (auto l : _vOutputLayer) attach(l, false);
    _vData = vData;
    _bDirty = true;
    _bAllDataLoaded = true;
    for (auto l : _vInputLayer) if (!l->_pDataSet) _bAllDataLoaded = false;
    for (auto l : _vOutputLayer) if (!l->_pDataSet) _bAllDataLoaded = false;
}

void NNNetwork::Randomize()
{
    for (auto w : _vWeight) w->Randomize();
}

void NNNetwork::SetBatch(uint32_t batch)
{
    if (batch != _batch) {
        _batch = batch;
        for (auto l : _vLayer) l->SetBatch(batch);
        _bDirty = true;
    }
}

void NNNetwork::SetPosition(uint32_t position)
{
    if (_bExamplesFound && position >= _examples) throw DsbEngineError("NNNetwork::SetPosition: Invalid position setting");
    _position = position;
}

bool NNNetwork::SetDecay(NNFloat decay) { if (decay < 0.0f) return false; _decay = decay; return true; }

void NNNetwork::SetTrainingMode(TrainingMode mode)
{
    if (_trainingMode != mode) { _trainingMode = mode; _bDirty = true; }
}

void NNNetwork::SetShuffleIndices(bool bShuffleIndices)
{
    if (_bShuffleIndices != bShuffleIndices) { _bShuffleIndices = bShuffleIndices; _bDirty = true; }
}

bool NNNetwork::SetSparsenessPenalty(NNFloat p, NNFloat beta)
{
    if (p < 0.0f || p > 1.0f) return false;
    _sparsenessPenalty_p = p; _sparsenessPenalty_beta = beta; _bSparsenessPenalty = (beta > 0.0f); _bDirty = true;
    return true;
}

bool NNNetwork::SetDenoising(NNFloat p)
{
    if (p < 0.0f || p >= 1.0f) return false;
    if (_denoising_p != p) { _denoising_p = p; _bDenoising = (p > 0.0f); _bDirty = true; }
    return true;
}

bool NNNetwork::SetDeltaBoost(NNFloat one, NNFloat zero)
{
    if (one < 0.0f || zero < 0.0f) return false;
    _deltaBoost_one = one; _deltaBoost_zero = zero; _bDirty = true;
    return true;
}

bool NNNetwork::SetSMCE(NNFloat oneTarget, NNFloat zeroTarget, NNFloat oneScale, NNFloat zeroScale)
{
    if (oneTarget < 0.0f || oneTarget > 1.0f || zeroTarget < 0.0f || zeroTarget > 1.0f || oneScale < 0.0f || zeroScale < 0.0f) return false;
    _SMCE_oneTarget = oneTarget; _SMCE_zeroTarget = zeroTarget; _SMCE_oneScale = oneScale; _SMCE_zeroScale = zeroScale; _bDirty = true;
    return true;
}

bool NNNetwork::SetCheckpoint(string name, int32_t interval)
{
    _checkpoint_name = name; _checkpoint_interval = interval;
    return true;
}

NNLayer* NNNetwork::GetLayer(const string& layer) const
{
    auto it = _mLayer.find(layer);
    return it == _mLayer.end() ? NULL : it->second;
}

vector<string> NNNetwork::GetLayers() const
{
    vector<string> v;
    for (auto l : _vLayer) v.push_back(l->_name);
    return v;
}

NNWeight* NNNetwork::GetWeight(const string& inputLayer, const string& outputLayer) const
{
    for (auto w : _vWeight)
        if (w->_inputLayer._name == inputLayer && w->_outputLayer._name == outputLayer) return w;
    return NULL;
}

uint64_t NNNetwork::GetBufferSize(const string& layer) const { NNLayer* l = GetLayer(layer); return l ? l->GetBufferSize() : 0; }
NNFloat* NNNetwork::GetUnitBuffer(const string& layer)
{
    NNLayer* l = GetLayer(layer);
    if (l) l->MaterializeUnits();
    return l ? l->GetUnitBuffer() : NULL;
}
NNFloat* NNNetwork::GetDeltaBuffer(const string& layer) { NNLayer* l = GetLayer(layer); return l ? l->GetDeltaBuffer() : NULL; }
NNFloat* NNNetwork::GetWeightBuffer(const string& inputLayer, const string& outputLayer)
{
    NNWeight* w = GetWeight(inputLayer, outputLayer);
    return w ? w->GetWeightBuffer() : NULL;
}

bool NNNetwork::LockWeights(const string& inputLayer, const string& outputLayer)
{
    NNWeight* w = GetWeight(inputLayer, outputLayer);
    if (!w) return false;
    w->Lock();
    return true;
}

bool NNNetwork::UnlockWeights(const string& inputLayer, const string& outputLayer)
{
    NNWeight* w = GetWeight(inputLayer, outputLayer);
    if (!w) return false;
    w->Unlock();
    return true;
}

NNFloat* NNNetwork::GetScratchBuffer(size_t size)
{
    if (size > _scratchBufferSize) {
        _pbScratchBuffer.reset(new GpuBuffer<NNFloat>(size));
        _scratchBufferSize = size;
    }
    return _pbScratchBuffer ? _pbScratchBuffer->_pDevData : NULL;
}

NNFloat* NNNetwork::GetP2PSendBuffer() { return _pbP2PBuffer ? _pbP2PBuffer->_pDevData : NULL; }

bool NNNetwork::P2P_Allreduce(NNFloat* pBuffer, size_t size)
{
    getGpu().Check(dsb200_all_reduce(getGpu()._ctx, pBuffer, size), "dsb200_all_reduce");
    return true;
}

void NNNetwork::AddBuffers(NNFloat* pDst, NNFloat* pSrc, uint64_t size)
{
    getGpu().Check(dsb200_add_buffers(getGpu()._ctx, pDst, pSrc, size), "dsb200_add_buffers");
}

// E/NNNetwork.cpp:2666-2714 allocates two IPC-exported peer buffers per rank for its copy ring.  Here: one local full-width
// buffer for the partial sums a reduce-scatter starts from, plus -- engine option "p2p_exchange", on by default -- the
// peer-mapped arena of csrc/comm.cu with four slots per layer (gathered units / reduced units / reduced delta / gathered delta).
void NNNetwork::AllocatePeerBuffers()
{
    if (getGpu()._numprocs <= 1) return;
    _maxStride = 0;
    for (auto w : _vWeight) {
        const uint32_t stride = w->_bOutgoingLarger ? w->_inputLayer._stride : w->_outputLayer._stride;
        _maxStride = max(_maxStride, stride);
    }
    const uint64_t need = (uint64_t)_maxStride * _batch;
    if (!_pbP2PBuffer || _pbP2PBuffer->_length < need) _pbP2PBuffer.reset(new GpuBuffer<NNFloat>(need));
    for (size_t i = 0; i < _vLayer.size(); i++) _vLayer[i]->_exchangeIndex = (uint32_t)i;
    const uint64_t slotFloats = (uint64_t)_batch * (_maxStride + (uint32_t)getGpu()._numprocs);
    if (!getGpu()._bP2PExchange) { _bPeerMemoryExchange = false; return; }
    if (_bPeerMemoryExchange && _peerSlotFloats >= slotFloats) return;           // mapped already and large enough
    // collective: every rank makes this call at the same point of RefreshState with the same sizes; any rank failing makes all fall back
    const int rc = dsb200_p2p_setup(getGpu()._ctx, (uint32_t)(4 * _vLayer.size()), slotFloats);
    _bPeerMemoryExchange = rc == 0;
    _peerSlotFloats = _bPeerMemoryExchange ? slotFloats : 0;
    if (rc != 0 && rc != DSB200_EUNSUPPORTED) getGpu().Check(rc, "dsb200_p2p_setup");
    if (!_bPeerMemoryExchange && getGpu()._id == 0) printf("NNNetwork: peer-memory exchange unavailable, the exchange steps use NCCL\n");
}

void NNNetwork::RefreshState()
{
    if (!_bAllDataLoaded) {
        _bAllDataLoaded = true;
        if (_mode != Prediction) for (auto l : _vOutputLayer) if (!l->_pDataSet) _bAllDataLoaded = false;
        for (auto l : _vInputLayer) if (!l->_pDataSet) _bAllDataLoaded = false;
    }
    if (!_bAllDataLoaded) throw DsbEngineError("NNNetwork::RefreshState: Attempt to use neural network " + _name + " without providing data sets for all input and output layers");
    // denoising_p > 0 turns the Denoising attribute on for sparse input layers at load time (E/NNNetwork.cpp:3705-3720)
    for (auto l : _vLayer) if (l->_bDirty) l->RefreshState(this, _trainingMode, _mode == Validation);
    for (auto w : _vWeight) w->RefreshState(this, _trainingMode);
    if (_bShuffleIndices && _mode == Training) RefreshShuffleBuffers();
    AllocatePeerBuffers();
    // split-row workspace of the sparse-Z kernel: worst case one partial row per 64 nnz plus one per example
    uint32_t maxBatchNnz = 0;
    for (auto l : _vInputLayer) {
        if (!l->_pDataSet || l->_vOutgoingLayer.empty()) continue;
        l->_pDataSet->GenerateSparseTransposedMatrix(_batch, l);
        uint32_t stride = 0;
        for (auto o : l->_vOutgoingLayer) stride = max(stride, o->_stride);
        const size_t items = (size_t)l->_pDataSet->_maxBatchNnz / 64 + _batch + 1;
        getGpu().Check(dsb200_ctx_reserve(getGpu()._ctx, _batch, items * stride), "dsb200_ctx_reserve");
        // the sparse gradient kernel sizes its heavy-column work lists from the most entries one batch can hold
        maxBatchNnz = max(maxBatchNnz, l->_pDataSet->_maxBatchNnz);
    }
    // the gradient kernel sums in fixed point, so the order inside a transposed column does not matter: skip the
    // canonical-order pass in the training loop (kernel-level callers keep it on by default)
    getGpu().Check(dsb200_ctx_set_option(getGpu()._ctx, "transpose_sort", 0), "dsb200_ctx_set_option");
    if (maxBatchNnz) getGpu().Check(dsb200_ctx_set_option(getGpu()._ctx, "wgrad_max_entries", (int)min<uint32_t>(maxBatchNnz + 64, 0x7fffffffu)), "dsb200_ctx_set_option");
    getGpu().SetNeuralNetwork(this);
    _bDirty = false;
}

void NNNetwork::RefreshShuffleBuffers()
{
    if (_shuffleIndices != _examples || !_pbShuffleIndex) {
        _shuffleIndices = _examples;
        _vShuffleIndex.resize(_examples);
        for (uint32_t i = 0; i < _examples; i++) _vShuffleIndex[i] = i;
        _pbShuffleIndex.reset(new GpuBuffer<uint32_t>(_examples));
        _pbShuffleIndex->Upload(_vShuffleIndex.data());
    }
}

static inline uint64_t mix64(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

// E/NNNetwork.cpp:874-907 restarts from the identity every epoch, sorts all `_examples` indices by cuRAND keys on rank 0 and
// broadcasts the result; the permutation itself is "parity unpinned" (it is whatever cuRAND's stream gives).  Here:
// the same restart from the identity, then Fisher-Yates over all indices on the host from the counter-based generator keyed by
// (seed, epoch) -- identical on every rank without a broadcast, and a function of the epoch number alone.
void NNNetwork::ShuffleIndices()
{
    RefreshShuffleBuffers();
    for (uint32_t i = 0; i < _examples; i++) _vShuffleIndex[i] = i;
    const uint64_t key = mix64((uint64_t)getGpu()._seed ^ mix64(0x5f3759dfull + _shuffleEpoch++));
    for (uint32_t i = _examples; i > 1; i--) {
        const uint32_t j = (uint32_t)(mix64(key + (uint64_t)i * 0x9e3779b97f4a7c15ull) % i);
        swap(_vShuffleIndex[i - 1], _vShuffleIndex[j]);
    }
    _pbShuffleIndex->Upload(_vShuffleIndex.data());
}

void NNNetwork::ClearUpdates()
{
    for (auto w : _vWeight) { w->_updateCount = 0; w->_bDeferredSparseGradient = false; w->_bDeferredDenseGradient = false; w->_nBiasPartials = 0; }
    for (auto l : _vLayer) l->ClearUpdates();
}

void NNNetwork::LoadBatch()
{
    uint32_t batch = _batch;
    if (_position + batch > _examples) batch = _examples - _position;
    for (auto l : _vInputLayer) {
        switch (_mode) {
        case Prediction: l->LoadPredictionBatch(_position, batch); break;
        case Training:   l->LoadTrainingBatch(_position, batch); break;
        case Validation: l->LoadValidationBatch(_position, batch); break;
        default: throw DsbEngineError("NNNetwork::LoadBatch: unsupported mode");
        }
    }
}

void NNNetwork::PredictBatch(uint32_t layers)
{
    if (layers > _vLayer.size()) return;
    if (_mode != Prediction || getGpu()._pNetwork != this) { _mode = Prediction; _bDirty = true; }      // (another network may have been published since)
    if (_bDirty) RefreshState();
    uint32_t batch = _batch;
    if (_position + batch > _examples) batch = _examples - _position;
    for (auto d : _vData) d->WaitForUpload(getGpu().GetStream());
    getGpu()._bDataConsumedValid = false;              // this pass reads the data sets without recording where it stops
    _bStepReadsRecorded = false;
    ClearUpdates();
    LoadBatch();
    for (auto l : _vFPOrder) l->ForwardPropagate(_position, batch, false);
}

void NNNetwork::PredictTrainingBatch(uint32_t layers)
{
    if (layers > _vLayer.size()) return;
    if (_bDirty) RefreshState();
    uint32_t batch = _batch;
    if (_position + batch > _examples) batch = _examples - _position;
    if (!_bBatchPrepared) LoadBatch();                    // TrainStep has it running on the side stream already
    for (auto l : _vFPOrder) l->ForwardPropagate(_position, batch, true);
}

// Everything of a training step that depends only on the data batch or on the weights as the last update left them runs on the
// side stream BESIDE the forward pass, in the order the main stream needs it:
//   _prepEvent  the target bitmap of the fused output-layer forward and the hi / lo copies of the output weights for the input-delta
//               kernel -- the loss pass waits for it ~40 us into the step;
//   _regEvent   the regularisation error -- the loss read-back waits for it;
//   _joinEvent  the transposed sparse matrix of the input batch (for a streamed batch including its capacity table, built on the
//               device) -- only the sparse weight gradient at the end of the step reads it: WaitForTransposed().
// (Round-2 step trace, tools/e2e_probe.py: with one event for all of it the loss left the device 140 us into the step, 40 us
// after the forward kernels were through, 162 us for a streamed batch.)
void NNNetwork::LaunchBatchPreparation(NNFloat lambda, NNFloat lambda1)
{
    uint32_t batch = _batch;
    if (_position + batch > _examples) batch = _examples - _position;
    cudaStream_t s = getGpu().GetStream();
    dsb200_ctx* ctx = getGpu()._ctx;
    RTERROR(cudaEventRecord(_forkEvent, s), "LaunchBatchPreparation fork");
    RTERROR(cudaStreamWaitEvent(_sideStream, _forkEvent, 0), "LaunchBatchPreparation fork wait");
    getGpu().Check(dsb200_ctx_set_stream(ctx, _sideStream), "dsb200_ctx_set_stream");
    for (auto l : _vOutputLayer) {
        if (!getGpu()._bFuseOutputGemm || !l->FusedOutputEligible(_errorFunction) || l->_vIncomingLayer.size() != 1) continue;
        if (l->_activation != Sigmoid || !l->_pDataSet || !(l->_pDataSet->_attributes & NNDataSetEnums::Boolean)) continue;
        dsb200_sparse v = l->_pDataSet->View();
        getGpu().Check(dsb200_gemm_fwd_output_prepare(ctx, &v, _position, batch, l->_localStride), "dsb200_gemm_fwd_output_prepare");
        NNLayer* in = l->_vIncomingLayer[0];
        NNWeight* w = l->_vIncomingWeight[0];
        if (in->_kind == NNLayer::Kind::Hidden && (getGpu()._numprocs == 1 || w->_bOutgoingLarger))
            getGpu().Check(dsb200_gemm_dx_prepare(ctx, batch, in->_stride, l->_localStride, w->_pbWeight->_pDevData), "dsb200_gemm_dx_prepare");
    }
    RTERROR(cudaEventRecord(_prepEvent, _sideStream), "LaunchBatchPreparation prep event");
    const bool any = lambda != (NNFloat)0.0 || lambda1 != (NNFloat)0.0;
    unsigned long long* acc = _pbErrorAccumulator->_pDevData;
    if (any)
        for (auto w : _vWeight)
            if (!w->_bShared)
                getGpu().Check(dsb200_regularization_error_async(ctx, lambda, lambda1, w->_pbWeight->_pDevData, w->_localSize, acc + 1),
                               "dsb200_regularization_error_async");
    RTERROR(cudaEventRecord(_regEvent, _sideStream), "LaunchBatchPreparation regularisation event");
    LoadBatch();
    getGpu().Check(dsb200_ctx_set_stream(ctx, s), "dsb200_ctx_set_stream");
    RTERROR(cudaEventRecord(_joinEvent, _sideStream), "LaunchBatchPreparation join");
    _bBatchPrepared = true;
    _bTransposePending = true;
}

void NNNetwork::WaitForTransposed()
{
    if (!_bTransposePending) return;
    RTERROR(cudaStreamWaitEvent(getGpu().GetStream(), _joinEvent, 0), "WaitForTransposed");
    _bTransposePending = false;
}

NNFloat NNNetwork::ReadErrorAccumulator()
{
    RTERROR(cudaMemcpyAsync(_pbErrorAccumulator->_pSysData, _pbErrorAccumulator->_pDevData, sizeof(unsigned long long),
                            cudaMemcpyDeviceToHost, getGpu().GetStream()), "ReadErrorAccumulator copy");
    RTERROR(cudaStreamSynchronize(getGpu().GetStream()), "ReadErrorAccumulator sync");
    return (NNFloat)((double)(long long)_pbErrorAccumulator->_pSysData[0] * (1.0 / 1073741824.0));
}

// asynchronous half of CalculateError: loss kernels (fused with activation + delta where possible) into the
// device accumulator, cross-rank sum of the fixed-point words, copy to pinned memory, record the event
void NNNetwork::LaunchError(NNFloat lambda, NNFloat lambda1)
{
    uint32_t batch = _batch;
    if (_position + batch > _examples) batch = _examples - _position;
    cudaStream_t s = getGpu().GetStream();
    unsigned long long* acc = _pbErrorAccumulator->_pDevData;        // [0] training error, [1] regularisation error
    if (!_bRegularizationLaunched) {
        RTERROR(cudaMemsetAsync(acc, 0, 2 * sizeof(unsigned long long), s), "LaunchError memset");
        LaunchRegularization(lambda, lambda1, false);
    }
    if (_bRegularizationLaunched) RTERROR(cudaStreamWaitEvent(s, _prepEvent, 0), "LaunchError prep wait");   // target bitmap, W copies
    _bRegularizationLaunched = false;
    for (auto l : _vOutputLayer) l->CalculateErrorAsync(_position, batch, _errorFunction, acc);
    if (_bStepReadsRecorded) {
        // every reader of the data sets' CSR buffers in this step has been launched -- sparse Z and the loss / delta pass on this stream,
        // the target bitmap and the transposed matrix on the side stream: the event is recorded where both streams are past them, and
        // the next batch may be uploaded beside the backward pass (NNDataSet::BeginUpload)
        getGpu().CopyStream();
        if (_bTransposePending) {
            RTERROR(cudaEventRecord(_lossPassEvent, s), "LaunchError loss pass event");
            RTERROR(cudaStreamWaitEvent(_sideStream, _lossPassEvent, 0), "LaunchError loss pass wait");
            RTERROR(cudaEventRecord(getGpu()._dataConsumedEvent, _sideStream), "LaunchError consumed event");
        } else {
            RTERROR(cudaEventRecord(getGpu()._dataConsumedEvent, s), "LaunchError consumed event");
        }
        getGpu()._bDataConsumedValid = true;
    }
    RTERROR(cudaStreamWaitEvent(s, _regEvent, 0), "LaunchError regularisation join");
    if (getGpu()._numprocs > 1)
        getGpu().Check(dsb200_all_reduce_u64(getGpu()._ctx, acc, 2), "dsb200_all_reduce_u64");
    RTERROR(cudaMemcpyAsync(_pbErrorAccumulator->_pSysData, acc, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s), "LaunchError copy");
    RTERROR(cudaEventRecord(_errorEvent, s), "LaunchError event");
}

// Regularisation error of every weight shard (E/NNNetwork.cpp:1724-1730) into the second fixed-point word of the
// accumulator (kCalculateRegularizationError, E/kernels.cu:2736-2744) -- no blocking Download per weight matrix.  The
// weights do not change between the end of one update and the next, so with `fork` the kernels run on a side stream
// BESIDE the forward pass (they read 2 x 14 MB of weights for BASELINE config 2) and join before the read-back.
void NNNetwork::LaunchRegularization(NNFloat lambda, NNFloat lambda1, bool fork)
{
    cudaStream_t s = getGpu().GetStream();
    unsigned long long* acc = _pbErrorAccumulator->_pDevData;
    const bool any = lambda != (NNFloat)0.0 || lambda1 != (NNFloat)0.0;
    if (fork && any) {
        RTERROR(cudaEventRecord(_forkEvent, s), "LaunchRegularization fork");
        RTERROR(cudaStreamWaitEvent(_sideStream, _forkEvent, 0), "LaunchRegularization fork wait");
        getGpu().Check(dsb200_ctx_set_stream(getGpu()._ctx, _sideStream), "dsb200_ctx_set_stream");
    }
    if (any)
        for (auto w : _vWeight)
            if (!w->_bShared)
                getGpu().Check(dsb200_regularization_error_async(getGpu()._ctx, lambda, lambda1, w->_pbWeight->_pDevData, w->_localSize, acc + 1),
                               "dsb200_regularization_error_async");
    if (fork && any) {
        getGpu().Check(dsb200_ctx_set_stream(getGpu()._ctx, s), "dsb200_ctx_set_stream");
        RTERROR(cudaEventRecord(_regEvent, _sideStream), "LaunchRegularization join");
    } else {
        RTERROR(cudaEventRecord(_regEvent, s), "LaunchRegularization join");
    }
}

tuple<NNFloat, NNFloat> NNNetwork::CalculateError(NNFloat lambda, NNFloat lambda1)
{
    LaunchError(lambda, lambda1);
    RTERROR(cudaEventSynchronize(_errorEvent), "CalculateError event sync");
    const NNFloat error_training = (NNFloat)((double)(long long)_pbErrorAccumulator->_pSysData[0] * (1.0 / 1073741824.0));
    const NNFloat error_regularization = (NNFloat)((double)(long long)_pbErrorAccumulator->_pSysData[1] * (1.0 / 1073741824.0));
    return make_tuple(error_training, error_regularization);
}

void NNNetwork::BackPropagate()
{
    uint32_t batch = _batch;
    if (_position + batch > _examples) batch = _examples - _position;
    for (auto l : _vBPOrder) {
        switch (l->_kind) {
        case NNLayer::Kind::Output:
            l->CalculateOutputDelta(_position, batch, _errorFunction);
            l->BackPropagate(_position, batch);
            break;
        case NNLayer::Kind::Hidden:
            l->BackPropagate(_position, batch);
            break;
        default: break;
        }
    }
}

void NNNetwork::UpdateWeights(NNFloat alpha, NNFloat lambda, NNFloat lambda1, NNFloat mu, NNFloat mu1)
{
    uint32_t batch = _batch;
    if (_position + batch > _examples) batch = _examples - _position;
    // The updates of different weights are independent.  With the fusions on, the sparse-input weight's update (gradient + optimizer
    // in one latency-bound kernel that leaves most issue slots idle) stays on the main stream and every update that touches no
    // scratch of the context -- one-launch small dense layers, an output weight whose bias gradient came out of the fused forward --
    // runs beside it on the side stream (round-2 launch list: 37 us of such updates in front of a 37 us sparse update).  The next
    // step's side-stream work (operand copies of the updated weights, regularisation) is ordered behind them by the stream itself;
    // the main stream joins before it leaves this function.
    bool side = false;
    cudaStream_t s = getGpu().GetStream();
    dsb200_ctx* ctx = getGpu()._ctx;
    vector<char> onSide(_vWeight.size(), 0);              // decided up front: the flags behind it are one-shot, cleared by the update itself
    if (_bFusion && _sideStream) {
        int nSide = 0, nMain = 0;
        for (size_t i = 0; i < _vWeight.size(); i++) { onSide[i] = _vWeight[i]->UpdateTouchesNoScratch() ? 1 : 0; if (onSide[i]) nSide++; else nMain++; }
        side = nSide > 0 && nMain > 0;
    }
    if (side) {
        RTERROR(cudaEventRecord(_forkEvent, s), "UpdateWeights fork");
        RTERROR(cudaStreamWaitEvent(_sideStream, _forkEvent, 0), "UpdateWeights fork wait");
        getGpu().Check(dsb200_ctx_set_stream(ctx, _sideStream), "dsb200_ctx_set_stream");
        for (int64_t i = (int64_t)_vWeight.size() - 1; i >= 0; i--)
            if (onSide[i]) _vWeight[i]->UpdateWeights(_trainingMode, batch, alpha, lambda, lambda1, mu, mu1, (NNFloat)_batches);
        RTERROR(cudaEventRecord(_updateEvent, _sideStream), "UpdateWeights side done");
        getGpu().Check(dsb200_ctx_set_stream(ctx, s), "dsb200_ctx_set_stream");
        for (int64_t i = (int64_t)_vWeight.size() - 1; i >= 0; i--)
            if (!onSide[i]) _vWeight[i]->UpdateWeights(_trainingMode, batch, alpha, lambda, lambda1, mu, mu1, (NNFloat)_batches);
        RTERROR(cudaStreamWaitEvent(s, _updateEvent, 0), "UpdateWeights join");
        return;
    }
    for (int64_t i = (int64_t)_vWeight.size() - 1; i >= 0; i--)
        _vWeight[i]->UpdateWeights(_trainingMode, batch, alpha, lambda, lambda1, mu, mu1, (NNFloat)_batches);
}

void NNNetwork::SetStepTrace(bool on)
{
    if (on && !_trace.start[0])
        for (int i = 0; i < StepTrace::kRing; i++) {
            RTERROR(cudaEventCreate(&_trace.start[i]), "step trace event");
            RTERROR(cudaEventCreate(&_trace.loss[i]), "step trace event");
            RTERROR(cudaEventCreate(&_trace.end[i]), "step trace event");
        }
    _trace.on = on; _trace.steps = 0;
    for (double& h : _trace.host) h = 0;
}

int NNNetwork::StepTraceReport(double* out, int cap)
{
    if (!out || cap < 9 || !_trace.steps) return 0;
    RTERROR(cudaDeviceSynchronize(), "step trace sync");
    for (int i = 0; i < 6; i++) out[i] = _trace.host[i] / (double)_trace.steps * 1e6;
    const uint64_t n = min<uint64_t>(_trace.steps, StepTrace::kRing);
    double fwd = 0, bwd = 0, gap = 0; uint64_t gaps = 0;
    for (uint64_t k = 0; k < n; k++) {
        const uint64_t step = _trace.steps - n + k;
        const int i = (int)(step % StepTrace::kRing);
        float ms = 0;
        RTERROR(cudaEventElapsedTime(&ms, _trace.start[i], _trace.loss[i]), "step trace elapsed"); fwd += ms;
        RTERROR(cudaEventElapsedTime(&ms, _trace.loss[i], _trace.end[i]), "step trace elapsed"); bwd += ms;
        if (k + 1 < n) { RTERROR(cudaEventElapsedTime(&ms, _trace.end[i], _trace.start[(i + 1) % StepTrace::kRing]), "step trace elapsed"); gap += ms; gaps++; }
    }
    out[6] = fwd / (double)n * 1e3; out[7] = bwd / (double)n * 1e3; out[8] = gaps ? gap / (double)gaps * 1e3 : 0;
    return 9;
}

// Body of the minibatch loop of NNNetwork::Train (E/NNNetwork.cpp:1601-1650) for the batch at `position`.
float NNNetwork::TrainStep(uint32_t position, NNFloat alpha, NNFloat lambda, NNFloat lambda1, NNFloat mu, NNFloat mu1, NNFloat* pRegularization)
{
    if (_mode != Training || getGpu()._pNetwork != this) { _mode = Training; _bDirty = true; }      // (another network may have been published since)
    if (_bDirty) RefreshState();
    SetPosition(position);
    ClearUpdates();
    using clk = std::chrono::steady_clock;
    const bool trace = _trace.on;
    const int ti = (int)(_trace.steps % StepTrace::kRing);
    clk::time_point t0, t1;
    auto lap = [&](int slot) { if (trace) { t1 = clk::now(); _trace.host[slot] += std::chrono::duration<double>(t1 - t0).count(); t0 = t1; } };
    if (trace) { RTERROR(cudaEventRecord(_trace.start[ti], getGpu().GetStream()), "step trace"); t0 = clk::now(); }
    for (auto d : _vData) d->WaitForUpload(getGpu().GetStream());      // a batch streamed in on the copy stream since the last step
    getGpu()._bDataConsumedValid = false;                            // until this step's readers are on their way (LaunchError)
    _bStepReadsRecorded = _bFusion;                                  // only the fused step waits for the side stream before the loss pass
    if (_bFusion) {                                  // batch preparation + regularisation error on the side stream while the forward pass runs
        RTERROR(cudaMemsetAsync(_pbErrorAccumulator->_pDevData, 0, 2 * sizeof(unsigned long long), getGpu().GetStream()), "TrainStep memset");
        LaunchBatchPreparation(lambda, lambda1);
        _bRegularizationLaunched = true;
    }
    lap(0);
    PredictTrainingBatch();
    _bBatchPrepared = false;
    lap(1);
    LaunchError(lambda, lambda1);                    // loss (+ output delta) kernels, join, and the async read-back
    if (trace) RTERROR(cudaEventRecord(_trace.loss[ti], getGpu().GetStream()), "step trace");
    lap(2);
    BackPropagate();                                 // queued behind them; does not depend on the host seeing the loss
    lap(3);
    RTERROR(cudaEventSynchronize(_errorEvent), "TrainStep event sync");
    lap(4);
    const NNFloat error_training = (NNFloat)((double)(long long)_pbErrorAccumulator->_pSysData[0] * (1.0 / 1073741824.0));
    const NNFloat error_regularization = (NNFloat)((double)(long long)_pbErrorAccumulator->_pSysData[1] * (1.0 / 1073741824.0));
    if (pRegularization) *pRegularization = error_regularization;

    // divergence brake, E/NNNetwork.cpp:1617-1639
    NNFloat step_alpha = (_decay <= 0.0) ? alpha : alpha * ((NNFloat)1.0 / ((NNFloat)1.0 + _decay * ((NNFloat)_batches)));
    _movingAverage = 0.9 * _movingAverage + 0.1 * error_training;
    if (_initSteps == 0) {
        if (error_training > 2.0 * _movingAverage) {
            _brakeSteps = 25;
            if (getGpu()._id == 0) printf("NNNetwork::Train: Detected network divergence, attempting recovery.\n");
        }
    } else _initSteps--;
    if (_brakeSteps > 0) { step_alpha *= (NNFloat)0.1; _brakeSteps--; }
    WaitForTransposed();                             // the sparse weight gradient inside the update is its first (and only) reader
    if (_brakeSteps < 24) {
        _batches++;                                  // before the update: Adam reads it (E/NNNetwork.cpp:1646-1649)
        UpdateWeights(step_alpha, lambda, lambda1, mu, mu1);
    }
    if (trace) { RTERROR(cudaEventRecord(_trace.end[ti], getGpu().GetStream()), "step trace"); lap(5); _trace.steps++; }
    return error_training;
}

float NNNetwork::Train(uint32_t epochs, NNFloat alpha, NNFloat lambda, NNFloat lambda1, NNFloat mu, NNFloat mu1)
{
    if (_mode != Training || getGpu()._pNetwork != this) { _mode = Training; _bDirty = true; }      // (another network may have been published since)
    if (_bDirty) RefreshState();
    if (_trainingMode != SGD && _bClearVelocity) {
        for (auto w : _vWeight) w->ClearVelocity();
        _batches = 0;
    }
    NNFloat total_error_training = 0, total_error_regularization = 0;
    NNFloat average_error_training = (NNFloat)FLT_MAX, average_error_regularization = 0;
    _movingAverage = 0; _brakeSteps = 0; _initSteps = 100;

    for (uint32_t epoch = 0; epoch < epochs; epoch++) {
        auto const start = std::chrono::steady_clock::now();
        total_error_training = 0; total_error_regularization = 0;
        if (_bDenoising) for (auto l : _vInputLayer) if (l->_bDenoising) l->GenerateDenoisingData();
        if (_bShuffleIndices) ShuffleIndices();
        for (uint32_t pos = 0; pos < GetExamples(); pos += GetBatch()) {
            NNFloat reg = 0;
            const NNFloat error_training = TrainStep(pos, alpha, lambda, lambda1, mu, mu1, &reg);
            uint32_t minibatch = GetBatch();
            if (_examples - pos < minibatch) minibatch = _examples - pos;
            total_error_training += error_training;
            total_error_regularization += reg * minibatch;
            if (_verbose && getGpu()._id == 0)
                printf("NNNetwork::Train: Minibatch@%u, average error %f, (%f training, %f regularization), alpha %f\n", pos,
                       error_training / minibatch + reg, error_training / minibatch, reg, alpha);
        }
        getGpu().Synchronize();
        auto const end = std::chrono::steady_clock::now();
        average_error_training = total_error_training / GetExamples();
        average_error_regularization = total_error_regularization / GetExamples();
        if (getGpu()._id == 0)
            printf("NNNetwork::Train: Epoch %d, average error %f, average training error %f, average regularization error %f, elapsed time %fs\n",
                   ++_epochs, average_error_training + average_error_regularization, average_error_training, average_error_regularization,
                   std::chrono::duration<double>(end - start).count());
        if (_checkpoint_interval > 0) {
            _checkpoint_epochs++;
            if (_checkpoint_epochs >= _checkpoint_interval) {
                string filename = _checkpoint_name + to_string(_epochs) + ".nc";
                if (getGpu()._id == 0) printf("NNNetwork::Train: saving checkpoint %s\n", filename.c_str());
                SaveNetCDF(filename);
                _checkpoint_epochs = 0;
            }
        }
    }
    return average_error_training + average_error_regularization;
}

// NNNetwork::Validate (E/NNNetwork.cpp:2459-2633): finite-difference check of every weight and bias gradient through the SAME kernels
// training uses -- here on the sparse path as it is (the reference's forces nothing either: its test networks are sparse, tst/test_data/
// validate_*.json).  delta = 1e-3 scaled by 1 / batch, threshold 20 * delta, SGD, no regularisation as in the reference, but with the
// CENTRED difference its own comment asks for (E/NNNetwork.cpp:2472-2474) and the loss read in 2^30 fixed point instead of fp32; fusions are switched off for the duration so that the weight gradients are materialised.  A matrix with more than
// _validateMaxSamples elements is checked on an evenly spaced sample (the reference walks every element of its 2 x 2 test networks).
bool NNNetwork::Validate()
{
    bool result = true;
    const NNFloat delta = (NNFloat)0.001, alpha = (NNFloat)1.0, lambda = (NNFloat)0.0, lambda1 = (NNFloat)0.0, mu = (NNFloat)0.0, mu1 = (NNFloat)0.0;
    const NNFloat epsilon = delta * 20.f;
    if (getGpu()._numprocs > 1) {
        cout << "NNNetwork::Validate: Do not call this method from a multi-process run, just don't, mmkay?" << endl;
        return false;
    }
    const bool fusion = _bFusion;
    const TrainingMode trainingMode = _trainingMode;
    _bFusion = false;
    SetTrainingMode(SGD);
    if (_mode != Validation || getGpu()._pNetwork != this) { _mode = Validation; _bDirty = true; }      // (another network may have been published since)
    if (_bDirty) RefreshState();
    cout << "Validating network weights and biases with epsilon error threshold of " << epsilon << endl;
    uint32_t batch = _batch;
    SetPosition(0);
    if (_position + batch > _examples) batch = _examples - _position;
    auto forwardError = [&]() {
        ClearUpdates();
        LoadBatch();
        for (auto l : _vFPOrder) l->ForwardPropagate(_position, batch, false);
        CalculateError(lambda, lambda1);
        // the loss as the kernels accumulated it (2^30 fixed point), NOT rounded to fp32: a probe moves it by ~delta * gradient = 1e-4,
        // below the fp32 resolution of a loss summed over a whole batch (the reference's 2 x 2 test networks never get there)
        return (double)(long long)_pbErrorAccumulator->_pSysData[0] * (1.0 / 1073741824.0) + (double)(long long)_pbErrorAccumulator->_pSysData[1] * (1.0 / 1073741824.0);
    };
    const double initialError = forwardError();
    cout << "initialError " << initialError << endl;
    BackPropagate();
    vector<vector<NNFloat>> vWeightGradient, vBiasGradient;
    for (auto w : _vWeight) {
        vWeightGradient.push_back(vector<NNFloat>(w->_localSize));
        w->_pbWeight->Download(w->_vWeight.data());
        w->_pbBias->Download(w->_vBias.data());
        w->_pbWeightGradient->Download(vWeightGradient.back().data());
    }
    // the bias gradient is not stored: take it from one SGD step with alpha = 1 and put the parameters back (E/NNNetwork.cpp:2548-2573)
    UpdateWeights(alpha, lambda, lambda1, mu, mu1);
    for (auto w : _vWeight) {
        vector<NNFloat> after(w->_localBiasSize);
        w->_pbBias->Download(after.data());
        for (size_t b = 0; b < after.size(); b++) after[b] -= w->_vBias[b];
        vBiasGradient.push_back(after);
        w->_pbWeight->Upload(w->_vWeight.data());
        w->_pbBias->Upload(w->_vBias.data());
    }
    for (size_t id = 0; id < _vWeight.size(); id++) {
        NNWeight* w = _vWeight[id];
        cout << "Validating weights between layer " << w->_inputLayer._name << " and " << w->_outputLayer._name << endl;
        const size_t nW = w->_localSize, stepW = max<size_t>(1, nW / _validateMaxSamples);
        for (size_t i = 0; i < nW; i += stepW) {
            const NNFloat g = vWeightGradient[id][i];
            NNFloat dEdW = 0;
            // A probe of delta moves the loss by ~1e-3 * gradient, and the fp32 rounding of the per-element losses moves it by ~1e-5 on a
            // loss of ~1e3 (oneScale 30): at the reference's delta the estimate carries noise of about half the threshold.  A miss is
            // therefore re-measured with an 8x wider probe (noise / 8, truncation error still ~1e-6) before it counts as a failure.
            for (int attempt = 0; attempt < 2; attempt++) {
                const NNFloat h = (attempt ? 8 : 1) * delta / (batch * w->_sharingCount);      // the gradient carries -1 / (sharing * batch)
                const NNFloat up = w->_vWeight[i] + h, down = w->_vWeight[i] - h;
                RTERROR(cudaMemcpy(w->_pbWeight->_pDevData + i, &up, sizeof(NNFloat), cudaMemcpyHostToDevice), "Validate poke");
                const double errorUp = forwardError();
                RTERROR(cudaMemcpy(w->_pbWeight->_pDevData + i, &down, sizeof(NNFloat), cudaMemcpyHostToDevice), "Validate poke");
                const double errorDown = forwardError();
                RTERROR(cudaMemcpy(w->_pbWeight->_pDevData + i, &w->_vWeight[i], sizeof(NNFloat), cudaMemcpyHostToDevice), "Validate restore");
                dEdW = (NNFloat)((errorUp - errorDown) / ((double)(up - down) * batch * w->_sharingCount));
                if (fabs(dEdW + g) <= epsilon) break;
            }
            if (fabs(dEdW + g) > epsilon) {
                cout << "Failed Weight " << i << " exceeds error threshold: " << dEdW << " vs " << g << endl;
                result = false;
            }
        }
        const size_t nB = w->_localBiasSize, stepB = max<size_t>(1, nB / _validateMaxSamples);
        for (size_t i = 0; i < nB; i += stepB) {
            const NNFloat g = vBiasGradient[id][i];
            NNFloat dEdb = 0;
            for (int attempt = 0; attempt < 2; attempt++) {
                const NNFloat h = (attempt ? 8 : 1) * delta / batch;
                const NNFloat up = w->_vBias[i] + h, down = w->_vBias[i] - h;
                RTERROR(cudaMemcpy(w->_pbBias->_pDevData + i, &up, sizeof(NNFloat), cudaMemcpyHostToDevice), "Validate poke");
                const double errorUp = forwardError();
                RTERROR(cudaMemcpy(w->_pbBias->_pDevData + i, &down, sizeof(NNFloat), cudaMemcpyHostToDevice), "Validate poke");
                const double errorDown = forwardError();
                RTERROR(cudaMemcpy(w->_pbBias->_pDevData + i, &w->_vBias[i], sizeof(NNFloat), cudaMemcpyHostToDevice), "Validate restore");
                dEdb = (NNFloat)((errorUp - errorDown) / ((double)(up - down) * batch));
                if (fabs(dEdb + g) <= epsilon) break;
            }
            if (fabs(dEdb + g) > epsilon) {
                cout << "Failed Bias " << i << " exceeds error threshold: " << dEdb << " vs " << g << endl;
                result = false;
            }
        }
    }
    _bFusion = fusion;
    SetTrainingMode(trainingMode);
    _mode = Training; _bDirty = true;
    return result;
}

// E/NNNetwork.cpp:1792-1822.  The reference caps k at 128 here; the kernel behind this takes k <= 1024.
void NNNetwork::CalculateTopK(const string& layer, uint32_t k, GpuBuffer<NNFloat>* pbKey, GpuBuffer<uint32_t>* pbValue)
{
    CalculateTopKFiltered(layer, k, NULL, pbKey, pbValue);
}

// SURVEY 8e "Top-K, P > 1": local top-K on [batch][N/P] (K' = K is sufficient for exactness) -> all-gather of (score, global id)
// [batch][K] -> final top-K over the P * K candidates of every row, on every rank.  Order: descending score, ties by ascending
// global id -- the gathered row is rank-major and every rank's list is already in that order, so the positional tie-break of
// dsb200_topk_kv is the global one.
void NNNetwork::CalculateTopKGlobal(const string& layer, uint32_t k, NNDataSetBase* pFilter, GpuBuffer<NNFloat>* pbKey, GpuBuffer<uint32_t>* pbValue)
{
    NNLayer* pLayer = GetLayer(layer);
    if (!pLayer) { if (getGpu()._id == 0) printf("NNNetwork::CalculateTopKGlobal: Unknown layer %s.\n", layer.c_str()); return; }
    const uint32_t P = (uint32_t)getGpu()._numprocs;
    if (P <= 1 || pLayer->_localStride == pLayer->_stride) { CalculateTopKFiltered(layer, k, pFilter, pbKey, pbValue); return; }
    uint32_t batch = _batch;
    if (_position + batch > _examples) batch = _examples - _position;
    if (!pbKey || !pbValue || pbKey->_length < (size_t)batch * k || pbValue->_length < (size_t)batch * k)
        throw DsbEngineError("NNNetwork::CalculateTopKGlobal: output buffers are too small");
    const size_t local = (size_t)batch * k, full = local * P;
    GpuBuffer<NNFloat> localKey(local), fullKey(full);
    GpuBuffer<uint32_t> localValue(local), fullValue(full);
    CalculateTopKFiltered(layer, k, pFilter, &localKey, &localValue);
    dsb200_ctx* ctx = getGpu()._ctx;
    getGpu().Check(dsb200_topk_offset(ctx, localValue._pDevData, local, pLayer->_minX), "dsb200_topk_offset");
    // [batch][k] of rank r lands in columns [k r, k (r + 1)) of [batch][P k]; ids travel as raw 32-bit words
    getGpu().Check(dsb200_all_gather(ctx, batch, P * k, localKey._pDevData, fullKey._pDevData), "dsb200_all_gather (top-K scores)");
    getGpu().Check(dsb200_all_gather(ctx, batch, P * k, reinterpret_cast<const float*>(localValue._pDevData),
                                     reinterpret_cast<float*>(fullValue._pDevData)), "dsb200_all_gather (top-K ids)");
    getGpu().Check(dsb200_topk_kv(ctx, fullKey._pDevData, fullValue._pDevData, batch, P * k, k, pbKey->_pDevData, pbValue->_pDevData), "dsb200_topk_kv");
    getGpu().Check(dsb200_ctx_sync(ctx), "dsb200_ctx_sync");          // the temporaries above are freed on return
}

void NNNetwork::CalculateTopKFiltered(const string& layer, uint32_t k, NNDataSetBase* pFilter, GpuBuffer<NNFloat>* pbKey, GpuBuffer<uint32_t>* pbValue)
{
    NNLayer* pLayer = GetLayer(layer);
    if (!pLayer) { if (getGpu()._id == 0) printf("NNNetwork::CalculateTopK: Unknown layer %s.\n", layer.c_str()); return; }
    pLayer->MaterializeUnits();
    uint32_t batch = _batch;
    if (_position + batch > _examples) batch = _examples - _position;
    if (!pbKey || !pbValue || pbKey->_length < (size_t)batch * k || pbValue->_length < (size_t)batch * k)
        throw DsbEngineError("NNNetwork::CalculateTopK: output buffers are too small");
    const uint64_t* fs = NULL; const uint64_t* fe = NULL; const uint32_t* fi = NULL;
    if (pFilter) {
        // the filter dataset is addressed by absolute example: offset the start / end tables to the batch
        fs = pFilter->_pbSparseStart->_pDevData + _position;
        fe = pFilter->_pbSparseEnd->_pDevData + _position;
        fi = pFilter->_pbSparseIndex->_pDevData;
    }
    getGpu().Check(dsb200_topk(getGpu()._ctx, pLayer->GetUnitBuffer(), batch, pLayer->_localStride, k, fs, fe, fi, pbKey->_pDevData, pbValue->_pDevData),
                   "dsb200_topk");
}
