// NNLayer.cpp -- fully-connected layer forward / backward on the dsstne_b200 C ABI.
//
// Control flow follows NNLayer::ForwardPropagateFullyConnected (E/NNLayer.cpp:994-1422),
// NNLayer::BackPropagateFullyConnected (E/NNLayer.cpp:2121-2640), CalculateError / CalculateOutputDelta
// (E/NNLayer.cpp:1710-1806), LoadTrainingBatch (E/NNLayer.cpp:910-948) and Reduce / Gather
// (E/NNLayer.cpp:2702-2826).  What changes underneath:
//  * every kernel is a dsb200_* call; the SGEMMs go through dsb200_gemm_{fwd,dw,dx};
//  * Reduce / Gather are one NCCL reduce-scatter / all-gather each (no peer-buffer ring, no
//    per-stage cudaDeviceSynchronize + MPI_Barrier);
//  * three fusions the reference does not have: bias + sparse Z + activation for a sparse input
//    layer, activation + loss + delta for a sparse-target output layer, and (in NNWeight) sparse
//    gradient + optimizer step.
#include "NNLayer.h"

#include <algorithm>
#include <sstream>

#include "NNNetwork.h"
#include "NNWeight.h"

using namespace std;

NNLayerDescriptor::NNLayerDescriptor()
    : _kind(NNLayer::Kind::Hidden), _type(NNLayer::Type::FullyConnected), _poolingFunction(None), _Nx(1), _Ny(1), _Nz(1), _Nw(1),
      _dimensions(1), _bDimensionsProvided(true), _weightInit(Xavier), _weightInitScale((NNFloat)1.0), _biasInit((NNFloat)0.0),
      _weightNorm((NNFloat)0.0), _deltaNorm((NNFloat)0.0), _pDropout((NNFloat)0.0), _activation(Activation::Sigmoid),
      _sparsenessPenalty_p((NNFloat)0.0), _sparsenessPenalty_beta((NNFloat)0.0), _attributes(NNLayer::Attributes::None),
      _RELUSlope(NAN), _ELUAlpha(NAN), _SELULambda(NAN)
{
}

NNLayer::NNLayer(NNLayerDescriptor& d, uint32_t batch)
    : _name(d._name), _kind(d._kind), _type(d._type), _attributes(d._attributes), _dataSet(d._dataSet), _pDataSet(NULL),
      _vSource(d._vSource), _Nx(d._Nx), _Ny(d._Ny), _Nz(d._Nz), _Nw(d._Nw), _batch(batch), _deltaUpdateCount(0), _unitUpdateCount(0),
      _dimensions(d._dimensions), _weightInit(d._weightInit), _weightInitScale(d._weightInitScale), _biasInit(d._biasInit),
      _RELUSlope(d._RELUSlope), _ELUAlpha(d._ELUAlpha), _SELULambda(d._SELULambda), _activation(d._activation), _pDropout(d._pDropout),
      _bSparse(d._attributes & NNLayer::Attributes::Sparse), _bFastSparse(false), _sparsenessPenalty_p(d._sparsenessPenalty_p),
      _sparsenessPenalty_beta(d._sparsenessPenalty_beta), _bDenoising(d._attributes & NNLayer::Attributes::Denoising),
      _weightNorm(d._weightNorm), _deltaNorm(d._deltaNorm), _parallelization(Serial), _bDirty(true), _bActivationPending(false),
      _bDeltaReady(false), _bHadamardDone(false), _bUnitsArePreActivation(false), _preActivationBatch(0), _priority(-1), _dropoutCalls(0)
{
    if (_type != FullyConnected) throw DsbEngineError("NNLayer: layer " + _name + ": only FullyConnected layers are on the dsstne_b200 hot path");
    if (_attributes & BatchNormalization) throw DsbEngineError("NNLayer: layer " + _name + ": batch normalisation is outside the hot path");
    if (!d._vSkip.empty()) throw DsbEngineError("NNLayer: layer " + _name + ": skip connections are outside the hot path");
    if (_pDropout >= (NNFloat)1.0) throw DsbEngineError("NNLayer: layer " + _name + ": pDropout must be below 1");
    _stride = _Nx * _Ny * _Nz * _Nw;
    _parallelization = Model;
    // E/NNLayer.cpp:108-112
    _minX = (uint32_t)(((size_t)_Nx * (size_t)getGpu()._id) / (size_t)getGpu()._numprocs);
    _maxX = (uint32_t)(((size_t)_Nx * (size_t)(getGpu()._id + 1)) / (size_t)getGpu()._numprocs);
    _localStride = (_maxX - _minX) * _Ny * _Nz * _Nw;
    _maxLocalStride = (((size_t)_Nx + getGpu()._numprocs - 1) / (size_t)getGpu()._numprocs) * _Ny * _Nz * _Nw;
}

NNLayer::~NNLayer() { Deallocate(); }

void NNLayer::Deallocate()
{
    _pbUnit.reset(); _pbDelta.reset(); _pbDropout.reset();
}

void NNLayer::Allocate(bool validate)
{
    Deallocate();
    const uint64_t size = (uint64_t)_maxLocalStride * _batch;
    // sparse input layers with the fast path never materialise their units (E/NNLayer.cpp:594-612)
    if (!(_bSparse && _bFastSparse && _kind == Input) || validate) {
        _vUnit.resize(size);
        _pbUnit.reset(new GpuBuffer<NNFloat>(size));
    }
    if (_kind != Input) {
        _vDelta.resize(size);
        _pbDelta.reset(new GpuBuffer<NNFloat>(size));
    }
}

void NNLayer::SetBatch(uint32_t batch)
{
    if (batch != _batch) { _batch = batch; _bDirty = true; }
}

void NNLayer::RefreshParallelization() { _parallelization = Model; }

void NNLayer::RefreshState(NNNetwork* pNetwork, TrainingMode trainingMode, bool validate)
{
    (void)pNetwork; (void)trainingMode;
    if (_bDirty) {
        _bFastSparse = false;
        if (_kind == Input && _pDataSet != NULL && _bSparse) {
            if (_pDataSet->_sparseDensity > (NNFloat)0.1)        // E/NNLayer.cpp:807
                throw DsbEngineError("NNLayer::RefreshState: sparse density of " + _name + " is above 0.1; the dense input path is outside the hot path");
            _bFastSparse = true;
        }
        if (getGpu()._numprocs > 1) RefreshParallelization();
        Allocate(validate);
        if (_kind != Hidden && _pDataSet != NULL) _pDataSet->Shard(NNDataSetEnums::Model);     // E/NNLayer.cpp:856-866
        _bDirty = false;
    }
    if (_kind == Input && _pDataSet) _pDataSet->SetDenoising(_bDenoising);
}

void NNLayer::ClearUpdates()
{
    _unitUpdateCount = 0; _deltaUpdateCount = 0; _bActivationPending = false; _bDeltaReady = false; _bForwardDeferred = false;
    _bUnitsGathered = false; _bBiasActDone = false;
}

void NNLayer::LoadPredictionBatch(uint32_t position, uint32_t batch)
{
    (void)position; (void)batch;
    if (_kind == Input && !_bFastSparse) throw DsbEngineError("NNLayer::LoadPredictionBatch: dense input layers are outside the hot path");
}

void NNLayer::LoadTrainingBatch(uint32_t position, uint32_t batch)
{
    if (_kind != Input) return;
    if (!_bFastSparse) throw DsbEngineError("NNLayer::LoadTrainingBatch: dense input layers are outside the hot path");
    if (_bDenoising) _pDataSet->CalculateSparseTransposedDenoisedMatrix(position, batch, this);   // E/NNLayer.cpp:920-924
    else             _pDataSet->CalculateSparseTransposedMatrix(position, batch, this);
}

void NNLayer::LoadValidationBatch(uint32_t position, uint32_t batch) { LoadTrainingBatch(position, batch); }

void NNLayer::GenerateDenoisingData()
{
    if (_pDataSet) _pDataSet->GenerateDenoisingData();
}

void NNLayer::ForwardPropagate(uint32_t position, uint32_t batch, bool bTraining)
{
    ForwardPropagateFullyConnected(position, batch, bTraining);
}

bool NNLayer::FusedOutputEligible(ErrorFunction ef) const
{
    // dsb200_output_pass: sigmoid / linear / relu-family activations with L2, CE or SMCE on sparse targets
    if (_kind != Output || !_pDataSet || !(_pDataSet->_attributes & NNDataSetEnums::Sparse)) return false;
    if (_pDataSet->_attributes & NNDataSetEnums::SparseIgnoreZero) return false;
    if (!getGpu()._pNetwork || !getGpu()._pNetwork->FusionEnabled()) return false;
    if (_activation == SoftMax || _pDropout > (NNFloat)0.0) return false;
    if (ef == CrossEntropy || ef == ScaledMarginalCrossEntropy) return _activation == Sigmoid;
    return ef == L2 && (_activation == Sigmoid || _activation == Linear || _activation == Tanh || _activation == RectifiedLinear);
}

void NNLayer::CalculateActivation(uint32_t batch)
{
    getGpu().Check(dsb200_activation(getGpu()._ctx, (int)_activation, GetUnitBuffer(), batch, _localStride, _RELUSlope, _ELUAlpha, _SELULambda),
                   "dsb200_activation");
}

// NNLayer::CalculateDropout (E/NNLayer.cpp:1685-1708): the mask of call number c of this layer is drawn from the
// counter-based generator with stream (layer order << 40 | c); see dsb200_dropout.
void NNLayer::CalculateDropout(uint32_t batch)
{
    const uint64_t stream = 0x4000000000000000ull | ((uint64_t)(uint32_t)_priority << 40) | (_dropoutCalls++ & 0xffffffffffull);
    getGpu().Check(dsb200_dropout(getGpu()._ctx, (int)_activation, GetUnitBuffer(), batch, _localStride, _stride, _minX * _Ny * _Nz * _Nw, _pDropout,
                                  _ELUAlpha, _SELULambda, (uint64_t)getGpu()._seed, stream), "dsb200_dropout");
}

// the deferred forward GEMM of an output layer, run on its own (the units are wanted, or the fused kernel declined the combination)
void NNLayer::RunDeferredForward(bool applyActivation)
{
    getGpu().Check(dsb200_gemm_fwd_bias_act(getGpu()._ctx, _preActivationBatch, _deferredK, _localStride, _pDeferredA,
                                            _vIncomingWeight[0]->_pbWeight->_pDevData, _vIncomingWeight[0]->_pbBias->_pDevData,
                                            applyActivation ? (int)_activation : (int)Linear, GetIncomingUnitBuffer(), _RELUSlope, _ELUAlpha, _SELULambda),
                   "dsb200_gemm_fwd_bias_act (deferred)");
    _bForwardDeferred = false;
}

void NNLayer::MaterializeUnits()
{
    if (_bForwardDeferred) {                         // the fused kernel never wrote z or a: run the layer now, activation included
        RunDeferredForward(true);
        _bUnitsArePreActivation = false;
        return;
    }
    if (!_bUnitsArePreActivation) return;
    _bUnitsArePreActivation = false;
    CalculateActivation(_preActivationBatch);
}

void NNLayer::ForwardPropagateFullyConnected(uint32_t position, uint32_t batch, bool bTraining)
{
    _bUnitsArePreActivation = false;                 // the unit buffer is about to be rewritten
    if (getGpu()._numprocs == 1) _bForwardDeferred = false;   // model parallel: the layer BELOW sets it in its own forward pass (it owns the GEMM)
    dsb200_ctx* ctx = getGpu()._ctx;
    NNNetwork* net = getGpu()._pNetwork;
    const bool deferActivation = bTraining && net && FusedOutputEligible(net->GetErrorFunction());
    if (getGpu()._numprocs == 1) {
        if (_kind == Input) return;
        if (_vIncomingLayer.empty()) throw DsbEngineError("NNLayer::ForwardPropagate: layer " + _name + " has no incoming layers");
        bool activated = false;
        // fused: bias + sparse Z + activation when the single source is a fast sparse input layer
        if (_vIncomingLayer.size() == 1 && _vIncomingLayer[0]->_bFastSparse && net && net->FusionEnabled() && !deferActivation &&
            (_activation == Sigmoid || _activation == Tanh || _activation == RectifiedLinear || _activation == Linear)) {
            NNLayer* in = _vIncomingLayer[0];
            in->_pDataSet->CalculateSparseZBiasActivation(position, batch, _stride, _vIncomingWeight[0]->_pbWeight->_pDevData,
                                                          _vIncomingWeight[0]->_pbBias->_pDevData, _activation, GetIncomingUnitBuffer(),
                                                          bTraining && in->_bDenoising);
            activated = true;
        } else if (_vIncomingLayer.size() == 1 && !_vIncomingLayer[0]->_bFastSparse && net && net->FusionEnabled() &&
                   (_activation == Sigmoid || _activation == Tanh || _activation == RectifiedLinear || _activation == Linear ||
                    _activation == LeakyRectifiedLinear || _activation == ExponentialLinear || _activation == ScaledExponentialLinear)) {
            // dense layer with a single source: bias + GEMM (+ activation unless the fused loss pass applies it) in one
            // call -- one tcgen05 kernel in the tensor-core GEMM modes (E/NNLayer.cpp:1009 + 1073 + 1157)
            NNLayer* in = _vIncomingLayer[0];
            if (deferActivation && getGpu()._bFuseOutputGemm && _activation == Sigmoid && _pDataSet && (_pDataSet->_attributes & NNDataSetEnums::Boolean)) {
                // nothing runs now -- CalculateErrorAsync runs GEMM + activation + loss + delta as one kernel
                _bForwardDeferred = true;
                _preActivationBatch = batch;
                _pDeferredA = in->GetUnitBuffer();
                _deferredK = in->_stride;
            } else {
                getGpu().Check(dsb200_gemm_fwd_bias_act(ctx, batch, in->_stride, _localStride, in->GetUnitBuffer(), _vIncomingWeight[0]->_pbWeight->_pDevData,
                                                        _vIncomingWeight[0]->_pbBias->_pDevData, deferActivation ? (int)Linear : (int)_activation,
                                                        GetIncomingUnitBuffer(), _RELUSlope, _ELUAlpha, _SELULambda), "dsb200_gemm_fwd_bias_act");
            }
            activated = !deferActivation;
        } else {
            // E/NNLayer.cpp:1002-1044: units start as the (sum of the) incoming biases
            getGpu().Check(dsb200_clear_unit(ctx, GetIncomingUnitBuffer(), _vIncomingWeight[0]->_pbBias->_pDevData, _stride, batch), "dsb200_clear_unit");
            for (size_t i = 1; i < _vIncomingLayer.size(); i++)
                getGpu().Check(dsb200_add_bias(ctx, GetIncomingUnitBuffer(), _vIncomingWeight[i]->_pbBias->_pDevData, _stride, batch), "dsb200_add_bias");
            const NNFloat sgemm_beta = (NNFloat)1.0;
            for (size_t i = 0; i < _vIncomingLayer.size(); i++) {
                NNLayer* in = _vIncomingLayer[i];
                NNFloat* pWeight = _vIncomingWeight[i]->_pbWeight->_pDevData;
                if (in->_bFastSparse) {                                                     // E/NNLayer.cpp:1046-1055
                    if (bTraining && in->_bDenoising) in->_pDataSet->CalculateSparseDenoisedZ(position, batch, _stride, pWeight, GetIncomingUnitBuffer(), sgemm_beta);
                    else                              in->_pDataSet->CalculateSparseZ(position, batch, _stride, pWeight, GetIncomingUnitBuffer(), sgemm_beta);
                } else {                                                                    // E/NNLayer.cpp:1057-1086
                    getGpu().Check(dsb200_gemm_fwd(ctx, batch, in->_stride, _localStride, in->GetUnitBuffer(), pWeight, sgemm_beta, GetIncomingUnitBuffer()), "dsb200_gemm_fwd");
                }
            }
        }
        if (!activated) {
            if (deferActivation) _bActivationPending = true;        // applied by the fused loss/delta pass
            else CalculateActivation(batch);                        // E/NNLayer.cpp:1157
        }
        if (bTraining && _pDropout > (NNFloat)0.0 && !_bActivationPending) CalculateDropout(batch);   // E/NNLayer.cpp:1160-1161
        return;
    }

    // ---------------------------------------------------------------- model parallel (E/NNLayer.cpp:1169-1422)
    if (_kind != Input) {
        bool biasActDone = _bBiasActDone;                                         // the feeding layer's GEMM epilogue did both
        _bBiasActDone = false;
        if (!_vIncomingLargerLayer.empty()) {
            // local partial products over this rank's input slice, then reduce-scatter over the units
            NNFloat sgemm_beta = (NNFloat)0.0;
            NNFloat* pSend = net->GetP2PSendBuffer();
            for (size_t i = 0; i < _vIncomingLargerLayer.size(); i++) {
                NNLayer* in = _vIncomingLargerLayer[i];
                NNFloat* pWeight = _vIncomingLargerWeight[i]->_pbWeight->_pDevData;
                if (in->_bFastSparse) {
                    // the sparse-Z kernels leave rows without non-zeros untouched (as the reference's do), and a column
                    // shard has many such rows: start from zeros and accumulate instead of relying on beta = 0
                    if (sgemm_beta == (NNFloat)0.0) {
                        RTERROR(cudaMemsetAsync(pSend, 0, (size_t)batch * _stride * sizeof(NNFloat), getGpu().GetStream()), "NNLayer::ForwardPropagate memset");
                        sgemm_beta = (NNFloat)1.0;
                    }
                    if (bTraining && in->_bDenoising) in->_pDataSet->CalculateSparseDenoisedZ(position, batch, _stride, pWeight, pSend, sgemm_beta);
                    else                              in->_pDataSet->CalculateSparseZ(position, batch, _stride, pWeight, pSend, sgemm_beta);
                } else {
                    getGpu().Check(dsb200_gemm_fwd(ctx, batch, in->_localStride, _stride, in->GetUnitBuffer(), pWeight, sgemm_beta, pSend), "dsb200_gemm_fwd");
                }
                sgemm_beta = (NNFloat)1.0;
            }
            // kAddBias + activation (E/NNLayer.cpp:1247-1340) ride on the exchange when this is the layer's only contribution
            const bool fuse = net->FusionEnabled() && _unitUpdateCount == 0 && _vIncomingLayer.size() == 1 && _activation != SoftMax && !biasActDone;
            Reduce(batch, _stride, GetIncomingUnitBuffer(), _localStride, _unitUpdateCount, UnitsReduce,
                   fuse ? _vIncomingWeight[0]->_pbBias->_pDevData : NULL, (fuse && !deferActivation) ? _activation : Linear);
            biasActDone = biasActDone || fuse;
            _unitUpdateCount++;
        }
        if (_bForwardDeferred) {
            _bActivationPending = true;                                           // CalculateErrorAsync runs GEMM + bias + activation + loss + delta as one kernel
        } else {
            if (!biasActDone) {
                for (size_t i = 0; i < _vIncomingLayer.size(); i++)                          // E/NNLayer.cpp:1247-1280
                    getGpu().Check(dsb200_add_bias(ctx, GetIncomingUnitBuffer(), _vIncomingWeight[i]->_pbBias->_pDevData, _localStride, batch), "dsb200_add_bias");
            }
            if (deferActivation) _bActivationPending = true;
            else if (!biasActDone) CalculateActivation(batch);
        }
        if (bTraining && _pDropout > (NNFloat)0.0 && !_bActivationPending) CalculateDropout(batch);   // E/NNLayer.cpp:1340-1341
    }
    // circulate activations to the outgoing larger layers (E/NNLayer.cpp:1340-1421)
    if (!_vOutgoingLargerLayer.empty()) {
        if (_bFastSparse)
            throw DsbEngineError("NNLayer::ForwardPropagate: sparse input layer " + _name + " feeding a wider layer is not supported model-parallel "
                                 "(the reference shards the dataset by columns but keeps full-height weights here, E/NNWeight.cpp:438 'BUG?')");
        _pGatheredUnits = Gather(batch, _stride, GetUnitBuffer(), _localStride, UnitsGather);
        _bUnitsGathered = bTraining;                                              // the backward pass of this step reuses it (E/NNLayer.cpp:2323 gathers again)
        for (size_t i = 0; i < _vOutgoingLargerLayer.size(); i++) {
            NNLayer* out = _vOutgoingLargerLayer[i];
            NNWeight* w = _vOutgoingLargerWeight[i];
            const bool only = net->FusionEnabled() && out->_vIncomingLayer.size() == 1 && out->_unitUpdateCount == 0 && out->_activation != SoftMax;
            const bool outDefer = bTraining && out->FusedOutputEligible(net->GetErrorFunction());
            if (only && outDefer && getGpu()._bFuseOutputGemm && out->_activation == Sigmoid && out->_pDataSet && (out->_pDataSet->_attributes & NNDataSetEnums::Boolean)) {
                // the output layer's forward GEMM is deferred into its loss / delta pass (dsb200_gemm_fwd_output_pass)
                out->_bForwardDeferred = true;
                out->_preActivationBatch = batch;
                out->_pDeferredA = _pGatheredUnits;
                out->_deferredK = _stride;
            } else if (only) {
                // bias + GEMM (+ activation unless the fused loss pass applies it) in one call
                getGpu().Check(dsb200_gemm_fwd_bias_act(ctx, batch, _stride, out->_localStride, _pGatheredUnits, w->_pbWeight->_pDevData, w->_pbBias->_pDevData,
                                                        outDefer ? (int)Linear : (int)out->_activation, out->GetIncomingUnitBuffer(),
                                                        out->_RELUSlope, out->_ELUAlpha, out->_SELULambda), "dsb200_gemm_fwd_bias_act");
                out->_bBiasActDone = true;
            } else {
                const NNFloat sgemm_beta = (out->_unitUpdateCount == 0) ? (NNFloat)0.0 : (NNFloat)1.0;
                getGpu().Check(dsb200_gemm_fwd(ctx, batch, _stride, out->_localStride, _pGatheredUnits, w->_pbWeight->_pDevData, sgemm_beta,
                                               out->GetIncomingUnitBuffer()), "dsb200_gemm_fwd");
            }
            out->_unitUpdateCount++;
        }
    }
}

NNFloat NNLayer::CalculateError(uint32_t position, uint32_t batch, ErrorFunction ef)
{
    if (_kind != Output) throw DsbEngineError("NNLayer::CalculateError: Attempt to calculate error on non-output layer " + _name);
    if (_bActivationPending) {
        // fused path asked for a synchronous value: run the fused pass into the network accumulator and read it back
        NNNetwork* net = getGpu()._pNetwork;
        unsigned long long* acc = net->GetErrorAccumulator();
        RTERROR(cudaMemsetAsync(acc, 0, sizeof(unsigned long long), getGpu().GetStream()), "CalculateError memset");
        CalculateErrorAsync(position, batch, ef, acc);
        return net->ReadErrorAccumulator();
    }
    switch (ef) {                                                                             // E/NNLayer.cpp:1720-1756
    case L2: return _pDataSet->CalculateL2Error(position, batch, _localStride, GetUnitBuffer());
    case CrossEntropy:
        return (_activation == SoftMax) ? _pDataSet->CalculateMultinomialCrossEntropyError(position, batch, _localStride, GetUnitBuffer())
                                        : _pDataSet->CalculateCrossEntropyError(position, batch, _localStride, GetUnitBuffer());
    case ScaledMarginalCrossEntropy:
        return (_activation == SoftMax) ? _pDataSet->CalculateMultinomialScaledMarginalCrossEntropyError(position, batch, _localStride, GetUnitBuffer())
                                        : _pDataSet->CalculateScaledMarginalCrossEntropyError(position, batch, _localStride, GetUnitBuffer());
    default: throw DsbEngineError("NNLayer::CalculateError: error function outside the hot path (L2, CrossEntropy, ScaledMarginalCrossEntropy)");
    }
}

bool NNLayer::CalculateErrorAsync(uint32_t position, uint32_t batch, ErrorFunction ef, unsigned long long* pDevAccumulator)
{
    if (_kind != Output) throw DsbEngineError("NNLayer::CalculateError: Attempt to calculate error on non-output layer " + _name);
    if (_bActivationPending && _bForwardDeferred) {
        // forward GEMM + activation + loss + delta as ONE tcgen05 kernel (csrc/gemm_stream.cu); the units are not produced at all
        // (MaterializeUnits re-runs the layer if somebody asks for them).  The kernel also leaves the column sums of delta, i.e.
        // the bias gradient, so NNWeight::UpdateWeights does not read delta again.
        NNWeight* w = _vIncomingWeight[0];
        dsb200_sparse v = _pDataSet->View();
        uint32_t nPartials = 0;
        const int rc = dsb200_gemm_fwd_output_pass(getGpu()._ctx, &v, (int)ef, (int)_activation, position, batch, _deferredK, _localStride, _pDeferredA,
                                                   w->_pbWeight->_pDevData, w->_pbBias->_pDevData, NULL, GetIncomingDeltaBuffer(), pDevAccumulator,
                                                   w->BiasPartialsBuffer(batch), &nPartials);
        if (rc == 0) {
            w->_nBiasPartials = nPartials;
            _bActivationPending = false;             // _bForwardDeferred stays set: the unit buffer holds nothing of this batch
            _bDeltaReady = true;
            return true;
        }
        if (rc != DSB200_EUNSUPPORTED) getGpu().Check(rc, "dsb200_gemm_fwd_output_pass");
        RunDeferredForward(false);                   // the kernel declined the combination: z now, then the usual fused pass below
    }
    if (_bActivationPending) {
        // ONE pass: a = f(z) in place, loss into the accumulator, delta written -- replaces kCalculate*Activation,
        // the Raw + NonZero error kernels and the Raw + NonZero delta kernels (six passes over [batch][N])
        // training only needs the delta of a sparse-target output layer: the activations are not stored (a third of this
        // pass's HBM traffic); anyone who reads the units afterwards gets them materialised first (MaterializeUnits)
        _pDataSet->CalculateFusedOutput(ef, _activation, position, batch, _localStride, GetUnitBuffer(), GetIncomingDeltaBuffer(), pDevAccumulator, false);
        _bUnitsArePreActivation = true;
        _preActivationBatch = batch;
        _bActivationPending = false;
        _bDeltaReady = true;
        return true;
    }
    return _pDataSet->CalculateErrorAsync(ef, _activation, position, batch, _localStride, GetUnitBuffer(), pDevAccumulator);
}

void NNLayer::CalculateOutputDelta(uint32_t position, uint32_t batch, ErrorFunction ef)
{
    if (_kind != Output) throw DsbEngineError("NNLayer::CalculateOutputDelta: Attempt to calculate output delta on non-output layer " + _name);
    if (_bDeltaReady) return;
    switch (ef) {                                                                             // E/NNLayer.cpp:1772-1790
    case CrossEntropy:
        _pDataSet->CalculateCrossEntropyOutputDelta(_activation, position, batch, _localStride, GetUnitBuffer(), GetIncomingDeltaBuffer()); break;
    case ScaledMarginalCrossEntropy:
        _pDataSet->CalculateScaledMarginalCrossEntropyOutputDelta(_activation, position, batch, _localStride, GetUnitBuffer(), GetIncomingDeltaBuffer()); break;
    case L2:
        _pDataSet->CalculateOutputDelta(_activation, position, batch, _localStride, GetUnitBuffer(), GetIncomingDeltaBuffer(), _RELUSlope, _ELUAlpha, _SELULambda); break;
    default: throw DsbEngineError("NNLayer::CalculateOutputDelta: error function outside the hot path");
    }
    if (_deltaNorm > (NNFloat)0.0) throw DsbEngineError("NNLayer::CalculateOutputDelta: DeltaNorm is outside the hot path");
}

void NNLayer::BackPropagate(uint32_t position, uint32_t batch) { BackPropagateFullyConnected(position, batch); }

void NNLayer::BackPropagateFullyConnected(uint32_t position, uint32_t batch)
{
    (void)position;
    dsb200_ctx* ctx = getGpu()._ctx;
    NNNetwork* net = getGpu()._pNetwork;
    auto hiddenLocal = [&]() {
        // E/NNLayer.cpp:2127-2138 (and :2451-2462): sparseness penalty, then f'(x) (dropout scale 1)
        if (_kind != Hidden) return;
        if (_bSparse && net->_bSparsenessPenalty) {
            const NNFloat p = (_sparsenessPenalty_p > (NNFloat)0.0) ? _sparsenessPenalty_p : net->_sparsenessPenalty_p;
            const NNFloat beta = (_sparsenessPenalty_beta > (NNFloat)0.0) ? _sparsenessPenalty_beta : net->_sparsenessPenalty_beta;
            getGpu().Check(dsb200_sparseness_penalty(ctx, batch, _localStride, GetUnitBuffer(), GetIncomingDeltaBuffer(), p, beta), "dsb200_sparseness_penalty");
        }
        const NNFloat scale = (NNFloat)1.0 / ((NNFloat)1.0 - _pDropout);
        if (_bHadamardDone) _bHadamardDone = false;              // applied by dsb200_gemm_dx_hadamard of the layer above
        else
            getGpu().Check(dsb200_hadamard(ctx, (int)_activation, (uint64_t)batch * _localStride, scale, GetUnitBuffer(), GetIncomingDeltaBuffer(),
                                           _RELUSlope, _ELUAlpha, _SELULambda), "dsb200_hadamard");
        if (_deltaNorm > (NNFloat)0.0) throw DsbEngineError("NNLayer::BackPropagate: DeltaNorm is outside the hot path");
    };

    if (getGpu()._numprocs == 1) {
        hiddenLocal();
        for (size_t i = 0; i < _vIncomingLayer.size(); i++) {                                // E/NNLayer.cpp:2187-2300
            NNLayer* in = _vIncomingLayer[i];
            NNWeight* w = _vIncomingWeight[i];
            if (!w->_bLocked) {
                const NNFloat sgemm_alpha = -(NNFloat)1.0 / (w->_sharingCount * (NNFloat)batch);
                const NNFloat sgemm_beta = (w->_updateCount == 0) ? (NNFloat)0.0 : (NNFloat)1.0;
                if (in->_kind == Input && in->_bFastSparse) {
                    if (net->FusionEnabled() && net->_mode == Training && sgemm_beta == (NNFloat)0.0 && w->_sharingCount == 1 && (_stride % 4 == 0)) {
                        // produced inside NNWeight::UpdateWeights, fused with the optimizer: dW is never written
                        w->_bDeferredSparseGradient = true;
                        w->_pDeferredDelta = GetDeltaBuffer();
                    } else {
                        net->WaitForTransposed();
                        in->_pDataSet->CalculateSparseTransposedWeightGradient(sgemm_alpha, sgemm_beta, in->_localStride, _localStride,
                                                                               GetDeltaBuffer(), w->_pbWeightGradient->_pDevData);
                    }
                } else if (net->FusionEnabled() && net->_mode == Training && sgemm_beta == (NNFloat)0.0 && w->_sharingCount == 1 &&
                           SmallDense(batch, in->_localStride, _localStride)) {
                    // small dense layer: gradient + optimizer + bias update as ONE launch from NNWeight::UpdateWeights
                    w->_bDeferredDenseGradient = true;
                    w->_pDeferredX = in->GetUnitBuffer();
                    w->_pDeferredDelta = GetDeltaBuffer();
                } else {
                    getGpu().Check(dsb200_gemm_dw(ctx, batch, in->_localStride, _localStride, sgemm_alpha, in->GetUnitBuffer(), GetDeltaBuffer(),
                                                  sgemm_beta, w->_pbWeightGradient->_pDevData), "dsb200_gemm_dw");
                }
                w->_updateCount++;
            }
            if (in->_kind != Input) {
                const NNFloat sgemm_beta = (in->_deltaUpdateCount == 0) ? (NNFloat)0.0 : (NNFloat)1.0;
                // small dense layer below with this layer as its only consumer, no sparseness penalty in between: the input
                // delta and its Hadamard product with f'(x) in one launch (E/NNLayer.cpp:2274 + 2137)
                const bool fuse = net->FusionEnabled() && in->_kind == Hidden && in->_vOutgoingLayer.size() == 1 && sgemm_beta == (NNFloat)0.0 &&
                                  !(in->_bSparse && net->_bSparsenessPenalty) && in->_deltaNorm <= (NNFloat)0.0;
                if (fuse) {
                    getGpu().Check(dsb200_gemm_dx_hadamard(ctx, batch, in->_localStride, _localStride, GetDeltaBuffer(), w->_pbWeight->_pDevData,
                                                           (int)in->_activation, (NNFloat)1.0 / ((NNFloat)1.0 - in->_pDropout), in->GetUnitBuffer(),
                                                           in->GetIncomingDeltaBuffer(), in->_RELUSlope, in->_ELUAlpha, in->_SELULambda),
                                   "dsb200_gemm_dx_hadamard");
                    in->_bHadamardDone = true;
                } else
                    getGpu().Check(dsb200_gemm_dx(ctx, batch, in->_localStride, _localStride, GetDeltaBuffer(), w->_pbWeight->_pDevData, sgemm_beta,
                                                  in->GetIncomingDeltaBuffer()), "dsb200_gemm_dx");
                in->_deltaUpdateCount++;
            }
        }
        return;
    }

    // ---------------------------------------------------------------- model parallel (E/NNLayer.cpp:2316-2626)
    NNFloat* pSend = net->GetP2PSendBuffer();
    if (!_vOutgoingLargerLayer.empty()) {
        // X(L) on every rank: every rank needs all of it for dW(L->L+1) of its output slice.  The forward pass of this step
        // gathered it already (the reference gathers again, E/NNLayer.cpp:2323)
        NNFloat* pX = _bUnitsGathered ? _pGatheredUnits : Gather(batch, _stride, GetUnitBuffer(), _localStride, UnitsGather);
        for (size_t i = 0; i < _vOutgoingLargerLayer.size(); i++) {
            NNLayer* out = _vOutgoingLargerLayer[i];
            NNWeight* w = _vOutgoingLargerWeight[i];
            if (w->_bLocked) continue;
            const NNFloat sgemm_alpha = -(NNFloat)1.0 / (w->_sharingCount * (NNFloat)batch);
            const NNFloat sgemm_beta = (w->_updateCount == 0) ? (NNFloat)0.0 : (NNFloat)1.0;
            if (net->FusionEnabled() && net->_mode == Training && sgemm_beta == (NNFloat)0.0 && w->_sharingCount == 1 && SmallDense(batch, _stride, out->_localStride)) {
                w->_bDeferredDenseGradient = true;                                // one launch in NNWeight::UpdateWeights; X(L) stays in its exchange slot
                w->_pDeferredX = pX;
                w->_pDeferredDelta = out->GetDeltaBuffer();
            } else
                getGpu().Check(dsb200_gemm_dw(ctx, batch, _stride, out->_localStride, sgemm_alpha, pX, out->GetDeltaBuffer(), sgemm_beta,
                                              w->_pbWeightGradient->_pDevData), "dsb200_gemm_dw");
            w->_updateCount++;
        }
        if (_kind != Input) {
            // partial delta(L) over this rank's output slices, then reduce-scatter
            NNFloat sgemm_beta = (NNFloat)0.0;
            for (size_t i = 0; i < _vOutgoingLargerLayer.size(); i++) {
                NNLayer* out = _vOutgoingLargerLayer[i];
                getGpu().Check(dsb200_gemm_dx(ctx, batch, _stride, out->_localStride, out->GetDeltaBuffer(), _vOutgoingLargerWeight[i]->_pbWeight->_pDevData,
                                              sgemm_beta, pSend), "dsb200_gemm_dx");
                sgemm_beta = (NNFloat)1.0;
            }
            Reduce(batch, _stride, GetIncomingDeltaBuffer(), _localStride, _deltaUpdateCount, DeltaReduce);
            _deltaUpdateCount++;
        }
    }
    hiddenLocal();
    if (!_vIncomingLargerLayer.empty()) {
        // delta(L) on every rank: dW and delta of the incoming larger layers need every unit of it.  It stays where the exchange
        // delivered it (an arena slot / a buffer of this layer) until the next step, so a deferred consumer can rely on it
        NNFloat* pD = Gather(batch, _stride, GetDeltaBuffer(), _localStride, DeltaGather);
        for (size_t i = 0; i < _vIncomingLargerLayer.size(); i++) {
            NNLayer* in = _vIncomingLargerLayer[i];
            NNWeight* w = _vIncomingLargerWeight[i];
            if (!w->_bLocked) {
                const NNFloat sgemm_alpha = -(NNFloat)1.0 / (w->_sharingCount * (NNFloat)batch);
                const NNFloat sgemm_beta = (w->_updateCount == 0) ? (NNFloat)0.0 : (NNFloat)1.0;
                if (in->_kind == Input && in->_bFastSparse) {
                    if (net->FusionEnabled() && net->_mode == Training && sgemm_beta == (NNFloat)0.0 && w->_sharingCount == 1 && (_stride % 4 == 0)) {
                        w->_bDeferredSparseGradient = true;                       // produced inside NNWeight::UpdateWeights, fused with the optimizer
                        w->_pDeferredDelta = pD;
                    } else {
                        net->WaitForTransposed();
                        in->_pDataSet->CalculateSparseTransposedWeightGradient(sgemm_alpha, sgemm_beta, in->_localStride, _stride, pD, w->_pbWeightGradient->_pDevData);
                    }
                } else {
                    getGpu().Check(dsb200_gemm_dw(ctx, batch, in->_localStride, _stride, sgemm_alpha, in->GetUnitBuffer(), pD, sgemm_beta,
                                                  w->_pbWeightGradient->_pDevData), "dsb200_gemm_dw");
                }
                w->_updateCount++;
            }
            if (in->_kind != Input) {
                const NNFloat beta2 = (in->_deltaUpdateCount == 0) ? (NNFloat)0.0 : (NNFloat)1.0;
                getGpu().Check(dsb200_gemm_dx(ctx, batch, in->_localStride, _stride, pD, w->_pbWeight->_pDevData, beta2, in->GetIncomingDeltaBuffer()), "dsb200_gemm_dx");
                in->_deltaUpdateCount++;
            }
        }
    }
}

// NNLayer::Reduce (E/NNLayer.cpp:2702-2761): the full [batch][stride] partial sums sit in the send buffer; this rank ends up with
// the sum over the ranks of its unit slice, optionally with bias and activation applied (the peer-memory kernel does both while it
// sums; on NCCL they are separate launches).
void NNLayer::Reduce(uint32_t batch, uint32_t stride, NNFloat* pBuffer, uint32_t localStride, uint32_t updateCount, ExchangeSlot slot,
                     const NNFloat* pBias, Activation activation)
{
    NNNetwork* net = getGpu()._pNetwork;
    if (getGpu()._numprocs == 1) return;
    dsb200_ctx* ctx = getGpu()._ctx;
    NNFloat* dst = pBuffer;
    if (updateCount > 0) dst = net->GetScratchBuffer((size_t)batch * localStride);
    if (net->PeerMemoryExchange()) {
        const bool epi = updateCount == 0;
        getGpu().Check(dsb200_p2p_reduce_scatter(ctx, 4 * _exchangeIndex + (uint32_t)slot, batch, stride, net->GetP2PSendBuffer(), dst, epi ? pBias : NULL,
                                                 epi ? (int)activation : (int)Linear, _RELUSlope, _ELUAlpha, _SELULambda), "dsb200_p2p_reduce_scatter");
        if (updateCount > 0) net->AddBuffers(pBuffer, dst, (uint64_t)batch * localStride);
        if (!epi && (pBias || activation != Linear)) throw DsbEngineError("NNLayer::Reduce: bias / activation on an accumulating exchange");
        return;
    }
    getGpu().Check(dsb200_reduce_scatter(ctx, batch, stride, net->GetP2PSendBuffer(), dst), "dsb200_reduce_scatter");
    if (updateCount > 0) { net->AddBuffers(pBuffer, dst, (uint64_t)batch * localStride); return; }
    if (pBias) getGpu().Check(dsb200_add_bias(ctx, pBuffer, pBias, localStride, batch), "dsb200_add_bias");
    if (activation != Linear)
        getGpu().Check(dsb200_activation(ctx, (int)activation, pBuffer, batch, localStride, _RELUSlope, _ELUAlpha, _SELULambda), "dsb200_activation");
}

// NNLayer::Gather (E/NNLayer.cpp:2764-2826): local slices -> the full [batch][stride] on every rank; returns where it is.
NNFloat* NNLayer::Gather(uint32_t batch, uint32_t stride, NNFloat* pBuffer, uint32_t localStride, ExchangeSlot slot)
{
    (void)localStride;
    if (getGpu()._numprocs == 1) return pBuffer;
    NNNetwork* net = getGpu()._pNetwork;
    dsb200_ctx* ctx = getGpu()._ctx;
    if (net->PeerMemoryExchange()) {
        const uint32_t s = 4 * _exchangeIndex + (uint32_t)slot;
        getGpu().Check(dsb200_p2p_all_gather(ctx, s, batch, stride, pBuffer), "dsb200_p2p_all_gather");
        return dsb200_p2p_slot(ctx, s);
    }
    unique_ptr<GpuBuffer<NNFloat>>& buf = (slot == DeltaGather) ? _pbGatheredDelta : _pbGatheredUnits;
    const uint64_t need = (uint64_t)_batch * stride;
    if (!buf || buf->_length < need) buf.reset(new GpuBuffer<NNFloat>(need));
    getGpu().Check(dsb200_all_gather(ctx, batch, stride, pBuffer, buf->_pDevData), "dsb200_all_gather");
    return buf->_pDevData;
}

bool NNLayer::GetUnits(vector<NNFloat>& vUnit)
{
    if (!_pbUnit) return false;
    MaterializeUnits();
    vUnit.resize(_pbUnit->_length);
    _pbUnit->Download(vUnit.data());
    return true;
}
bool NNLayer::GetUnits(NNFloat* pUnit) { if (!_pbUnit) return false; MaterializeUnits(); _pbUnit->Download(pUnit); return true; }
bool NNLayer::SetUnits(const vector<NNFloat>& vUnit)
{
    if (!_pbUnit || vUnit.size() < _pbUnit->_length) return false;
    _pbUnit->Upload(vUnit.data());
    return true;
}
bool NNLayer::GetDeltas(vector<NNFloat>& vDelta)
{
    if (!_pbDelta) return false;
    vDelta.resize(_pbDelta->_length);
    _pbDelta->Download(vDelta.data());
    return true;
}
bool NNLayer::GetDeltas(NNFloat* pDelta) { if (!_pbDelta) return false; _pbDelta->Download(pDelta); return true; }
bool NNLayer::SetDeltas(const vector<NNFloat>& vDelta)
{
    if (!_pbDelta || vDelta.size() < _pbDelta->_length) return false;
    _pbDelta->Upload(vDelta.data());
    return true;
}
