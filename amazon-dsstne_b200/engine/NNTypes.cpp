// NNTypes.cpp -- NNDataSet<T> for sparse datasets on the dsstne_b200 C ABI.
//
// Follows the behaviour of E/NNTypes.cpp for sparse (Boolean / analog, weighted, indexed) data:
// constructors :501-581, LoadSparseData :610-650, LoadIndexedData / LoadDataWeight :730-770,
// CalculateSparseDatapointCounts :1427-1492, GenerateSparseTransposedMatrix :1520-1570,
// Shard(Model) :1845-1955, and the attribute -> kernel dispatch of E/NNTypes.h:527-1182.
// Differences by design: Shard() slices locally (every rank holds the host copy; no MPI_Send /
// MPI_Recv), denoising randoms come from a counter-based generator in the kernel library, and the
// loss entry points have asynchronous / fused siblings (see NNTypes.h "B200 additions").
#include "NNTypes.h"

#include <algorithm>
#include <sstream>

#include "NNLayer.h"
#include "NNNetwork.h"

using namespace std;

NNDataSetBase::NNDataSetBase()
    : _name(""), _dataType(NNDataSetEnums::Float), _attributes(0), _examples(0), _uniqueExamples(0), _localExamples(0),
      _dimensions(0), _width(0), _height(0), _length(0), _stride(0), _sharding(NNDataSetEnums::None), _minX(0), _maxX(0),
      _sparseDataSize(0), _sparseDensity(0), _sparseTransposedIndices(0), _bDenoising(false), _bDirty(true),
      _bStreaming(false), _bIndexed(false), _batch(0), _denoisingEpoch(0), _maxBatchNnz(0)
{
}

NNDataSetBase::NNDataSetBase(const string& name, NNDataSetEnums::DataType dataType, uint32_t examples, uint32_t uniqueExamples,
                             const NNDataSetDimensions& dim)
    : NNDataSetBase()
{
    _name = name; _dataType = dataType; _examples = examples; _uniqueExamples = uniqueExamples; _localExamples = examples;
    _dimensions = dim._dimensions; _width = dim._width; _height = dim._height; _length = dim._length;
}

template <typename T>
NNDataSet<T>::NNDataSet(uint32_t examples, NNFloat sparseDensity, const NNDataSetDimensions& dim, bool isWeighted, const string& name)
    : NNDataSet(examples, examples, (size_t)(((double)dim._width * dim._height * dim._length * examples) * sparseDensity), dim,
                false, isWeighted, name)
{
    _attributes = NNDataSetEnums::Sparse;
    if (isWeighted) _attributes |= NNDataSetEnums::Weighted;
}

template <typename T>
NNDataSet<T>::NNDataSet(uint32_t examples, uint32_t uniqueExamples, size_t sparseDataSize, const NNDataSetDimensions& dim,
                        bool isIndexed, bool isWeighted, const string& name)
    : NNDataSetBase(name, NNDataSetEnums::getDataType<T>(), examples, uniqueExamples, dim)
{
    _attributes = NNDataSetEnums::Sparse;
    _sparseDataSize = sparseDataSize;
    _vSparseStart.assign(_uniqueExamples, 0);
    _vSparseEnd.assign(_uniqueExamples, 0);
    _vSparseData.resize(_sparseDataSize);
    _vSparseIndex.assign(_sparseDataSize, 0);
    // even placeholder rows so Shard() keeps a well-formed CSR before data is loaded
    const size_t per = _uniqueExamples ? (_sparseDataSize + _uniqueExamples - 1) / _uniqueExamples : 0;
    for (uint32_t i = 0; i < _uniqueExamples; i++) {
        _vSparseStart[i] = i ? _vSparseEnd[i - 1] : 0;
        _vSparseEnd[i] = _vSparseStart[i] + per;
    }
    _pbSparseStart.reset(new GpuBuffer<uint64_t>(_vSparseStart.size()));
    _pbSparseEnd.reset(new GpuBuffer<uint64_t>(_vSparseEnd.size()));
    _pbSparseData.reset(new GpuBuffer<T>(_vSparseData.size()));
    _pbSparseIndex.reset(new GpuBuffer<uint32_t>(_vSparseIndex.size()));
    if (isIndexed) {
        _attributes |= NNDataSetEnums::Indexed;
        _bIndexed = true;
        _vIndex.assign(_examples, 0);
        _pbIndex.reset(new GpuBuffer<uint32_t>(_vIndex.size()));
    }
    if (isWeighted) {
        _attributes |= NNDataSetEnums::Weighted;
        _vDataWeight.resize(_examples);
        _pbDataWeight.reset(new GpuBuffer<NNFloat>(_vDataWeight.size()));
    }
}

template <typename T>
void NNDataSet<T>::UploadSparse()
{
    _pbSparseStart->Upload(_vSparseStart.data());
    _pbSparseEnd->Upload(_vSparseEnd.data());
    if (!_vSparseIndex.empty()) _pbSparseIndex->Upload(_vSparseIndex.data());
    if (!(_attributes & NNDataSetEnums::Boolean) && !_vSparseData.empty()) _pbSparseData->Upload(_vSparseData.data());
}

template <typename T>
void NNDataSet<T>::CopySparseData(const uint64_t* srcSparseStart, const uint64_t* srcSparseEnd, const void* srcSparseData,
                                  const uint32_t* srcSparseIndex)
{
    if (!(_attributes & NNDataSetEnums::Sparse)) throw std::runtime_error("Cannot set sparse data on a non sparse NNDataSet");
    if (srcSparseStart[0] != 0) throw std::runtime_error("Sparse data should be zero indexed; srcSparseStart[0] != 0");
    const uint64_t dataLength = srcSparseEnd[_uniqueExamples - 1];
    if (dataLength > _vSparseData.size() || dataLength > _vSparseIndex.size()) {
        stringstream msg;
        msg << "Not enough space to store sparse data. Allocated: " << _vSparseData.size() << " Required: " << dataLength;
        throw std::length_error(msg.str());
    }
    copy(srcSparseStart, srcSparseStart + _uniqueExamples, _vSparseStart.data());
    copy(srcSparseEnd, srcSparseEnd + _uniqueExamples, _vSparseEnd.data());
    copy(srcSparseIndex, srcSparseIndex + dataLength, _vSparseIndex.data());
    if (srcSparseData) {
        const T* typed = static_cast<const T*>(srcSparseData);
        copy(typed, typed + dataLength, _vSparseData.data());
    } else {
        _attributes |= NNDataSetEnums::Boolean;          // no values => all datapoints are 1
    }
    _bDirty = true;
}

template <typename T>
void NNDataSet<T>::LoadSparseData(const uint64_t* srcSparseStart, const uint64_t* srcSparseEnd, const void* srcSparseData,
                                  const uint32_t* srcSparseIndex)
{
    if (_sharding == NNDataSetEnums::Model && getGpu()._numprocs > 1) {
        // model parallel: every rank receives the full batch, keeps it as the un-sharded host copy and re-slices its columns
        if (srcSparseStart[0] != 0) throw std::runtime_error("Sparse data should be zero indexed; srcSparseStart[0] != 0");
        const uint64_t dataLength = srcSparseEnd[_uniqueExamples - 1];
        if (dataLength > _sparseDataSize) {
            stringstream msg; msg << "Not enough space to store sparse data. Allocated: " << _sparseDataSize << " Required: " << dataLength;
            throw std::length_error(msg.str());
        }
        _vFullSparseStart.assign(srcSparseStart, srcSparseStart + _uniqueExamples);
        _vFullSparseEnd.assign(srcSparseEnd, srcSparseEnd + _uniqueExamples);
        _vFullSparseIndex.assign(srcSparseIndex, srcSparseIndex + dataLength);
        if (srcSparseData) { const T* typed = static_cast<const T*>(srcSparseData); _vFullSparseData.assign(typed, typed + dataLength); }
        else _attributes |= NNDataSetEnums::Boolean;
        SliceFromFull();
        return;
    }
    if (getGpu()._bPinnedMirror && !_mirror.unavailable) {
        // the previous batch's asynchronous copies read the mirror: they must have finished before it is overwritten (a whole
        // training step has passed in a streaming loop, so this wait is normally free)
        if (_mirror.pending) { RTERROR(cudaEventSynchronize(_mirror.done), "NNDataSet mirror wait"); _mirror.pending = false; }
        CopySparseData(srcSparseStart, srcSparseEnd, srcSparseData, srcSparseIndex);
        if (UploadMirrorAsync(srcSparseEnd[_uniqueExamples - 1])) return;
        UploadSparseAsync(srcSparseStart, srcSparseEnd, srcSparseData, srcSparseIndex, srcSparseEnd[_uniqueExamples - 1]);
        return;
    }
    CopySparseData(srcSparseStart, srcSparseEnd, srcSparseData, srcSparseIndex);
    UploadSparseAsync(srcSparseStart, srcSparseEnd, srcSparseData, srcSparseIndex, srcSparseEnd[_uniqueExamples - 1]);
}

// Uploads of a streamed batch go to the copy stream when the network has recorded where the last step stopped reading the CSR
// buffers (GpuContext::_dataConsumedEvent): they then overlap that step's backward pass instead of queueing behind it.
template <typename T>
cudaStream_t NNDataSet<T>::BeginUpload()
{
    if (!getGpu()._bDataConsumedValid) return getGpu().GetStream();
    cudaStream_t cs = getGpu().CopyStream();
    RTERROR(cudaStreamWaitEvent(cs, getGpu()._dataConsumedEvent, 0), "NNDataSet upload wait");
    return cs;
}

template <typename T>
void NNDataSet<T>::EndUpload(cudaStream_t stream, cudaEvent_t done)
{
    _uploadEvent = (stream != getGpu().GetStream()) ? done : nullptr;
}

template <typename T>
void NNDataSet<T>::WaitForUpload(cudaStream_t stream)
{
    if (_uploadEvent) { RTERROR(cudaStreamWaitEvent(stream, _uploadEvent, 0), "NNDataSet upload join"); _uploadEvent = nullptr; }
}

// Engine option "pinned_mirror" (default on; 141 -> 84 us of host time for the two data sets of a config-2 batch).  The staging
// path copies every batch twice on the host (mirror, then pinned staging); here the mirror vectors are page-locked in
// place (cudaHostRegister, once -- their storage never moves after construction) and are themselves the source of the
// asynchronous copies, so a batch is copied once.
template <typename T>
bool NNDataSet<T>::UploadMirrorAsync(uint64_t dataLength)
{
    void* want[4] = {_vSparseStart.data(), _vSparseEnd.data(), _vSparseIndex.data(), _vSparseData.data()};
    const size_t bytes[4] = {_vSparseStart.size() * sizeof(uint64_t), _vSparseEnd.size() * sizeof(uint64_t), _vSparseIndex.size() * sizeof(uint32_t),
                             _vSparseData.size() * sizeof(T)};
    for (int i = 0; i < 4; i++) {
        if (_mirror.ptr[i] == want[i] || !bytes[i]) continue;
        if (_mirror.ptr[i]) { cudaHostUnregister(_mirror.ptr[i]); _mirror.ptr[i] = nullptr; }
        if (cudaHostRegister(want[i], bytes[i], cudaHostRegisterDefault) != cudaSuccess) {
            cudaGetLastError();
            for (int j = 0; j < 4; j++) if (_mirror.ptr[j]) { cudaHostUnregister(_mirror.ptr[j]); _mirror.ptr[j] = nullptr; }
            _mirror.unavailable = true;
            return false;
        }
        _mirror.ptr[i] = want[i];
    }
    if (!_mirror.done) RTERROR(cudaEventCreateWithFlags(&_mirror.done, cudaEventDisableTiming), "NNDataSet mirror event");
    cudaStream_t stream = BeginUpload();
    RTERROR(cudaMemcpyAsync(_pbSparseStart->_pDevData, _vSparseStart.data(), _uniqueExamples * sizeof(uint64_t), cudaMemcpyHostToDevice, stream), "NNDataSet upload");
    RTERROR(cudaMemcpyAsync(_pbSparseEnd->_pDevData, _vSparseEnd.data(), _uniqueExamples * sizeof(uint64_t), cudaMemcpyHostToDevice, stream), "NNDataSet upload");
    if (dataLength) RTERROR(cudaMemcpyAsync(_pbSparseIndex->_pDevData, _vSparseIndex.data(), dataLength * sizeof(uint32_t), cudaMemcpyHostToDevice, stream), "NNDataSet upload");
    if (!(_attributes & NNDataSetEnums::Boolean) && dataLength)
        RTERROR(cudaMemcpyAsync(_pbSparseData->_pDevData, _vSparseData.data(), dataLength * sizeof(T), cudaMemcpyHostToDevice, stream), "NNDataSet upload");
    RTERROR(cudaEventRecord(_mirror.done, stream), "NNDataSet mirror record");
    _mirror.pending = true;
    EndUpload(stream, _mirror.done);
    return true;
}

// The call a serving / streaming caller makes once per batch (the reference's JNI binding, java/.../dsstne.cpp): the source
// arrays go through pinned staging (two sets, so the caller never waits for the previous batch's copy) and only the used
// part of the index / value arrays travels, as asynchronous copies on the engine's stream -- no stream synchronisation.
template <typename T>
void NNDataSet<T>::UploadSparseAsync(const uint64_t* srcStart, const uint64_t* srcEnd, const void* srcData, const uint32_t* srcIndex, uint64_t dataLength)
{
    Staging& st = _staging[_stagingCur];
    _stagingCur ^= 1;
    if (!st.start) {
        RTERROR(cudaHostAlloc((void**)&st.start, _vSparseStart.size() * sizeof(uint64_t), cudaHostAllocDefault), "NNDataSet staging");
        RTERROR(cudaHostAlloc((void**)&st.end, _vSparseEnd.size() * sizeof(uint64_t), cudaHostAllocDefault), "NNDataSet staging");
        RTERROR(cudaHostAlloc((void**)&st.index, max<size_t>(_vSparseIndex.size(), 1) * sizeof(uint32_t), cudaHostAllocDefault), "NNDataSet staging");
        RTERROR(cudaHostAlloc((void**)&st.data, max<size_t>(_vSparseData.size(), 1) * sizeof(T), cudaHostAllocDefault), "NNDataSet staging");
        RTERROR(cudaEventCreateWithFlags(&st.done, cudaEventDisableTiming), "NNDataSet staging event");
        st.capacity = max<size_t>(_vSparseIndex.size(), 1);
    }
    if (st.pending) { RTERROR(cudaEventSynchronize(st.done), "NNDataSet staging wait"); st.pending = false; }
    cudaStream_t stream = BeginUpload();
    memcpy(st.start, srcStart, _uniqueExamples * sizeof(uint64_t));
    memcpy(st.end, srcEnd, _uniqueExamples * sizeof(uint64_t));
    memcpy(st.index, srcIndex, dataLength * sizeof(uint32_t));
    RTERROR(cudaMemcpyAsync(_pbSparseStart->_pDevData, st.start, _uniqueExamples * sizeof(uint64_t), cudaMemcpyHostToDevice, stream), "NNDataSet upload");
    RTERROR(cudaMemcpyAsync(_pbSparseEnd->_pDevData, st.end, _uniqueExamples * sizeof(uint64_t), cudaMemcpyHostToDevice, stream), "NNDataSet upload");
    if (dataLength) RTERROR(cudaMemcpyAsync(_pbSparseIndex->_pDevData, st.index, dataLength * sizeof(uint32_t), cudaMemcpyHostToDevice, stream), "NNDataSet upload");
    if (srcData && dataLength) {
        memcpy(st.data, srcData, dataLength * sizeof(T));
        RTERROR(cudaMemcpyAsync(_pbSparseData->_pDevData, st.data, dataLength * sizeof(T), cudaMemcpyHostToDevice, stream), "NNDataSet upload");
    }
    RTERROR(cudaEventRecord(st.done, stream), "NNDataSet staging record");
    st.pending = true;
    EndUpload(stream, st.done);
}

template <typename T>
NNDataSet<T>::~NNDataSet()
{
    if (_mirror.pending) cudaEventSynchronize(_mirror.done);
    for (void* p : _mirror.ptr) if (p) cudaHostUnregister(p);
    if (_mirror.done) cudaEventDestroy(_mirror.done);
    for (Staging& st : _staging) {
        if (st.pending) cudaEventSynchronize(st.done);
        if (st.start) cudaFreeHost(st.start);
        if (st.end) cudaFreeHost(st.end);
        if (st.index) cudaFreeHost(st.index);
        if (st.data) cudaFreeHost(st.data);
        if (st.done) cudaEventDestroy(st.done);
    }
}

template <typename T>
void NNDataSet<T>::CopySparseData(const long* srcSparseStart, const long* srcSparseEnd, const void* srcSparseData,
                                  const long* srcSparseIndex)
{
    if (!(_attributes & NNDataSetEnums::Sparse)) throw std::runtime_error("Cannot set sparse data on a non sparse NNDataSet");
    vector<uint64_t> s(srcSparseStart, srcSparseStart + _uniqueExamples), e(srcSparseEnd, srcSparseEnd + _uniqueExamples);
    const uint64_t n = e[_uniqueExamples - 1];
    vector<uint32_t> idx(n);
    for (uint64_t i = 0; i < n; i++) idx[i] = (uint32_t)srcSparseIndex[i];
    CopySparseData(s.data(), e.data(), srcSparseData, idx.data());
}

template <typename T>
void NNDataSet<T>::LoadSparseData(const long* srcSparseStart, const long* srcSparseEnd, const void* srcSparseData,
                                  const long* srcSparseIndex)
{
    CopySparseData(srcSparseStart, srcSparseEnd, srcSparseData, srcSparseIndex);
    UploadSparse();
}

template <typename T>
void NNDataSet<T>::LoadIndexedData(const uint32_t* srcIndexedData)
{
    if (!(_attributes & NNDataSetEnums::Indexed)) throw std::runtime_error("Cannot set indexed data on a non indexed NNDataSet");
    copy(srcIndexedData, srcIndexedData + _vIndex.size(), _vIndex.data());
    _pbIndex->Upload(_vIndex.data());
    _bDirty = true;
}

template <typename T>
void NNDataSet<T>::LoadDataWeight(const NNFloat* srcWeightData)
{
    if (!(_attributes & NNDataSetEnums::Weighted)) throw std::runtime_error("Cannot set weight data on a non weighted NNDataSet");
    copy(srcWeightData, srcWeightData + _vDataWeight.size(), _vDataWeight.data());
    _pbDataWeight->Upload(_vDataWeight.data());
}

// E/NNTypes.cpp:1845-1955.  Every rank holds the full host copy (single node), so the shard is a
// local column slice [width*r/P, width*(r+1)/P) with indices rebased to the shard.
template <typename T>
bool NNDataSet<T>::Shard(NNDataSetEnums::Sharding sharding)
{
    if (sharding == _sharding) return true;
    if (_sharding != NNDataSetEnums::None) UnShard();
    if (sharding != NNDataSetEnums::Model) {
        if (sharding == NNDataSetEnums::Data) throw DsbEngineError("NNDataSet::Shard: data-parallel sharding is not implemented (nor in the reference, E/NNTypes.cpp:2091-2095)");
        return true;
    }
    const size_t P = (size_t)getGpu()._numprocs, r = (size_t)getGpu()._id;
    _sharding = NNDataSetEnums::Model;
    _minX = (uint32_t)(((size_t)_width * r) / P);
    _maxX = (uint32_t)(((size_t)_width * (r + 1)) / P);
    if (P == 1) { UploadSparse(); return true; }
    _vFullSparseStart = _vSparseStart; _vFullSparseEnd = _vSparseEnd; _vFullSparseIndex = _vSparseIndex; _vFullSparseData = _vSparseData;
    SliceFromFull();
    return true;
}

// local column slice [_minX, _maxX) of the full host copy, indices rebased to the shard, uploaded.
// Called once per step by a streaming caller (LoadSparseData while model parallel), so it follows UploadSparseAsync: the slice is
// written straight into page-locked staging (two sets), the copies are asynchronous on the upload stream and nothing synchronises
// -- the first version built fresh vectors, copied from pageable memory and ended in cudaStreamSynchronize, which put the whole
// host pass in series with the training step (round-2 bench, 2 GPUs: 1.07 ms per step end to end against 0.37 ms device-resident).
template <typename T>
void NNDataSet<T>::SliceFromFull()
{
    const bool analog = !(_attributes & NNDataSetEnums::Boolean);
    const size_t cap = max<size_t>(max<size_t>(_sparseDataSize, _vFullSparseIndex.size()), 1);
    Staging& st = _staging[_stagingCur];
    _stagingCur ^= 1;
    if (!st.start || st.capacity < cap) {
        if (st.pending) { RTERROR(cudaEventSynchronize(st.done), "NNDataSet staging wait"); st.pending = false; }
        if (st.start) { cudaFreeHost(st.start); cudaFreeHost(st.end); cudaFreeHost(st.index); cudaFreeHost(st.data); }
        RTERROR(cudaHostAlloc((void**)&st.start, max<size_t>(_uniqueExamples, 1) * sizeof(uint64_t), cudaHostAllocDefault), "NNDataSet staging");
        RTERROR(cudaHostAlloc((void**)&st.end, max<size_t>(_uniqueExamples, 1) * sizeof(uint64_t), cudaHostAllocDefault), "NNDataSet staging");
        RTERROR(cudaHostAlloc((void**)&st.index, cap * sizeof(uint32_t), cudaHostAllocDefault), "NNDataSet staging");
        RTERROR(cudaHostAlloc((void**)&st.data, cap * sizeof(T), cudaHostAllocDefault), "NNDataSet staging");
        if (!st.done) RTERROR(cudaEventCreateWithFlags(&st.done, cudaEventDisableTiming), "NNDataSet staging event");
        st.capacity = cap;
    }
    if (st.pending) { RTERROR(cudaEventSynchronize(st.done), "NNDataSet staging wait"); st.pending = false; }
    const uint32_t* fullIndex = _vFullSparseIndex.data();
    const T* fullData = analog ? _vFullSparseData.data() : nullptr;
    const uint32_t lo = _minX, span = _maxX - _minX;
    uint64_t n = 0;
    for (uint32_t j = 0; j < _uniqueExamples; j++) {
        const uint64_t s = _vFullSparseStart[j], e = _vFullSparseEnd[j];
        st.start[j] = n;
        if (analog) {
            for (uint64_t k = s; k < e; k++) {
                const uint32_t c = fullIndex[k] - lo;                          // one unsigned compare covers both bounds
                if (c < span) { st.index[n] = c; st.data[n] = fullData[k]; n++; }
            }
        } else {
            for (uint64_t k = s; k < e; k++) {
                const uint32_t c = fullIndex[k] - lo;
                st.index[n] = c;                                               // branch-free: the slot is overwritten unless the entry stays
                n += c < span;
            }
        }
        st.end[j] = n;
    }
    // the host mirror of the shard (what Shard / SaveNetCDF / the capacity pass read)
    _vSparseStart.assign(st.start, st.start + _uniqueExamples);
    _vSparseEnd.assign(st.end, st.end + _uniqueExamples);
    _vSparseIndex.assign(st.index, st.index + n);
    if (analog) _vSparseData.assign(st.data, st.data + n);
    if (!_pbSparseIndex || _pbSparseIndex->_length < max<size_t>(n, 1)) {
        RTERROR(cudaStreamSynchronize(getGpu().GetStream()), "NNDataSet shard buffer");   // readers of the old buffer
        _pbSparseIndex.reset(new GpuBuffer<uint32_t>(cap));
    }
    if (analog && (!_pbSparseData || _pbSparseData->_length < max<size_t>(n, 1))) {
        RTERROR(cudaStreamSynchronize(getGpu().GetStream()), "NNDataSet shard buffer");
        _pbSparseData.reset(new GpuBuffer<T>(cap));
    }
    cudaStream_t stream = BeginUpload();
    RTERROR(cudaMemcpyAsync(_pbSparseStart->_pDevData, st.start, _uniqueExamples * sizeof(uint64_t), cudaMemcpyHostToDevice, stream), "NNDataSet shard upload");
    RTERROR(cudaMemcpyAsync(_pbSparseEnd->_pDevData, st.end, _uniqueExamples * sizeof(uint64_t), cudaMemcpyHostToDevice, stream), "NNDataSet shard upload");
    if (n) RTERROR(cudaMemcpyAsync(_pbSparseIndex->_pDevData, st.index, n * sizeof(uint32_t), cudaMemcpyHostToDevice, stream), "NNDataSet shard upload");
    if (analog && n) RTERROR(cudaMemcpyAsync(_pbSparseData->_pDevData, st.data, n * sizeof(T), cudaMemcpyHostToDevice, stream), "NNDataSet shard upload");
    RTERROR(cudaEventRecord(st.done, stream), "NNDataSet staging record");
    st.pending = true;
    EndUpload(stream, st.done);
    _bDirty = true;
}

template <typename T>
bool NNDataSet<T>::UnShard()
{
    if (_sharding == NNDataSetEnums::Model && getGpu()._numprocs > 1 && !_vFullSparseStart.empty()) {
        _vSparseStart.swap(_vFullSparseStart); _vSparseEnd.swap(_vFullSparseEnd);
        _vSparseIndex.swap(_vFullSparseIndex); _vSparseData.swap(_vFullSparseData);
        _vFullSparseStart.clear(); _vFullSparseEnd.clear(); _vFullSparseIndex.clear(); _vFullSparseData.clear();
        _pbSparseIndex.reset(new GpuBuffer<uint32_t>(_vSparseIndex.size()));
        if (!(_attributes & NNDataSetEnums::Boolean)) _pbSparseData.reset(new GpuBuffer<T>(_vSparseData.size()));
        UploadSparse();
        _bDirty = true;
    }
    _sharding = NNDataSetEnums::None;
    return true;
}

// E/NNTypes.cpp:1427-1492
template <typename T>
bool NNDataSet<T>::CalculateSparseDatapointCounts()
{
    if (!(_attributes & NNDataSetEnums::Sparse)) return false;
    const uint64_t N = (_sharding == NNDataSetEnums::Model && getGpu()._numprocs > 1) ? (_maxX - _minX)
                                                                                     : (uint64_t)_width * _height * _length;
    _vSparseDatapointCount.assign(N, 0);
    _vSparseMaxDatapointCount.assign(N, 0);
    _vSparseMultiDatapointCount.assign(N, 0);
    vector<uint32_t> vCount(N, 0);
    vector<uint32_t> vExampleCount(_uniqueExamples, 0);
    if (_attributes & NNDataSetEnums::Indexed) { for (size_t i = 0; i < _examples; i++) vExampleCount[_vIndex[i]]++; }
    else fill(vExampleCount.begin(), vExampleCount.end(), 1);
    for (size_t i = 0; i < _uniqueExamples; i++) {
        for (uint64_t j = _vSparseStart[i]; j < _vSparseEnd[i]; j++) {
            if (_vSparseIndex[j] >= N) {
                stringstream msg; msg << "NNDataSet::CalculateSparseDatapointCounts: vCount address = " << _vSparseIndex[j] << " >= vCount size = " << N;
                throw std::out_of_range(msg.str());
            }
            vCount[_vSparseIndex[j]]++;
        }
        for (uint64_t j = _vSparseStart[i]; j < _vSparseEnd[i]; j++) {
            const uint32_t x = _vSparseIndex[j];
            if (vCount[x] > 0) {
                _vSparseMaxDatapointCount[x] = max(_vSparseMaxDatapointCount[x], vCount[x]);
                if (vCount[x] > 1) _vSparseMultiDatapointCount[x] += vExampleCount[i];
                _vSparseDatapointCount[x] += (uint64_t)vExampleCount[i] * vCount[x];
                vCount[x] = 0;
            }
        }
    }
    const uint64_t denom = (uint64_t)_uniqueExamples * (uint64_t)_width * _height * _length;
    _sparseDensity = denom ? (NNFloat)((double)_vSparseIndex.size() / (double)denom) : 0.0f;
    return true;
}

// E/NNTypes.cpp:1520-1570
template <typename T>
bool NNDataSet<T>::GenerateSparseTransposedMatrix(uint32_t batch, NNLayer* pLayer)
{
    // A dataset no larger than one batch that is re-loaded every step (streaming / serving callers): the capacity table is
    // built on the device from exact per-column counts (dsb200_transposed_capacity) -- no host pass over the entries,
    // no read-back; the index buffer is sized by its upper bound nnz + 31 * N.
    if (_bDirty && !(_attributes & NNDataSetEnums::Indexed) && _uniqueExamples <= batch && !(_sharding == NNDataSetEnums::Model && getGpu()._numprocs > 1) &&
        getGpu()._pNetwork && getGpu()._pNetwork->FusionEnabled()) {
        uint64_t N = (uint64_t)_width * _height * _length;
        if (pLayer) { uint32_t Nx, Ny, Nz, Nw; tie(Nx, Ny, Nz, Nw) = pLayer->GetLocalDimensions(); N = max<uint64_t>(N, (uint64_t)Nx * Ny * Nz * Nw); }
        if (_vSparseTransposedStart.size() != N) _vSparseTransposedStart.assign(N, 0);        // size only: the table itself lives on the device
        if (!_pbSparseTransposedStart || _pbSparseTransposedStart->_length < N) _pbSparseTransposedStart.reset(new GpuBuffer<uint32_t>(N));
        if (!_pbSparseTransposedEnd || _pbSparseTransposedEnd->_length < N) _pbSparseTransposedEnd.reset(new GpuBuffer<uint32_t>(N));
        if (!_pbColumnCount || _pbColumnCount->_length < N) _pbColumnCount.reset(new GpuBuffer<uint32_t>(N));
        const uint64_t bound = (uint64_t)_vSparseIndex.size() + 31ull * N + 32;
        if (bound > _sparseTransposedIndices || !_pbSparseTransposedIndex) {
            _sparseTransposedIndices = bound;
            _pbSparseTransposedIndex.reset(new GpuBuffer<uint32_t>(_sparseTransposedIndices));
            if (!(_attributes & NNDataSetEnums::Boolean) || (_attributes & NNDataSetEnums::Weighted))
                _pbSparseTransposedData.reset(new GpuBuffer<NNFloat>(_sparseTransposedIndices));
        }
        dsb200_sparse v = View();
        getGpu().Check(dsb200_transposed_capacity(getGpu()._ctx, &v, _uniqueExamples, (uint32_t)N, _pbColumnCount->_pDevData,
                                                  _pbSparseTransposedStart->_pDevData, NULL), "dsb200_transposed_capacity");
        _batch = batch;
        _maxBatchNnz = (uint32_t)min<uint64_t>(_vSparseEnd[_uniqueExamples - 1] - _vSparseStart[0], 0xffffffffu);
        _bDirty = false;
        return true;
    }
    if (_bDirty) { CalculateSparseDatapointCounts(); _bDirty = false; }
    const uint64_t NData = _vSparseDatapointCount.size();
    uint64_t NLayer = NData;
    if (pLayer) { uint32_t Nx, Ny, Nz, Nw; tie(Nx, Ny, Nz, Nw) = pLayer->GetLocalDimensions(); NLayer = (uint64_t)Nx * Ny * Nz * Nw; }
    const uint64_t N = max(NData, NLayer);
    _vSparseTransposedStart.assign(N, 0);
    if (!_pbSparseTransposedStart || _pbSparseTransposedStart->_length < N) _pbSparseTransposedStart.reset(new GpuBuffer<uint32_t>(N));
    if (!_pbSparseTransposedEnd || _pbSparseTransposedEnd->_length < N) _pbSparseTransposedEnd.reset(new GpuBuffer<uint32_t>(N));
    _batch = batch;
    uint32_t offset = 0;
    for (size_t i = 0; i < NData; i++) {
        _vSparseTransposedStart[i] = offset;
        size_t size1 = min((size_t)batch, (size_t)_vSparseDatapointCount[i]);
        if (_vSparseMaxDatapointCount[i] > 1) {
            const size_t size2 = min((size_t)_vSparseMaxDatapointCount[i] * batch,
                                     (size_t)batch + (size_t)(_vSparseMaxDatapointCount[i] - 1) * _vSparseMultiDatapointCount[i]);
            size1 = max(size1, size2);
        }
        offset += (uint32_t)size1;
        offset = ((offset + 31) >> 5) << 5;
    }
    _pbSparseTransposedStart->Upload(_vSparseTransposedStart.data());
    if (offset > _sparseTransposedIndices || !_pbSparseTransposedIndex) {
        _sparseTransposedIndices = offset;
        _pbSparseTransposedIndex.reset(new GpuBuffer<uint32_t>(_sparseTransposedIndices));
        if (!(_attributes & NNDataSetEnums::Boolean) || (_attributes & NNDataSetEnums::Weighted))
            _pbSparseTransposedData.reset(new GpuBuffer<NNFloat>(_sparseTransposedIndices));
    }
    // workspace sizing for the split-row path of the sparse-Z kernel: worst nnz of a batch window
    uint64_t worst = 0;
    if (!(_attributes & NNDataSetEnums::Indexed)) {
        for (uint32_t p = 0; p < _uniqueExamples; p += batch) {
            const uint32_t last = min(_uniqueExamples, p + batch) - 1;
            worst = max<uint64_t>(worst, _vSparseEnd[last] - _vSparseStart[p]);
        }
    } else worst = _vSparseIndex.size();
    _maxBatchNnz = (uint32_t)min<uint64_t>(worst, 0xffffffffu);
    return true;
}

template <typename T>
dsb200_sparse NNDataSet<T>::View()
{
    dsb200_sparse v;
    v.sparseStart = _pbSparseStart->_pDevData;
    v.sparseEnd = _pbSparseEnd->_pDevData;
    v.sparseIndex = _pbSparseIndex->_pDevData;
    v.sparseData = (_attributes & NNDataSetEnums::Boolean) ? nullptr : (const void*)_pbSparseData->_pDevData;
    v.dataType = (int32_t)_dataType;
    v.dataWeight = (_attributes & NNDataSetEnums::Weighted) ? _pbDataWeight->_pDevData : nullptr;
    v.index = (_attributes & NNDataSetEnums::Indexed) ? _pbIndex->_pDevData : nullptr;
    v.denoisingRandom = _pbDenoisingRandom ? _pbDenoisingRandom->_pDevData : nullptr;
    return v;
}

template <typename T>
bool NNDataSet<T>::CalculateSparseTransposedMatrix(uint32_t position, uint32_t batch, NNLayer* pLayer)
{
    if (_bDirty || batch != _batch) GenerateSparseTransposedMatrix(batch, pLayer);     // E/NNTypes.h:570-573
    dsb200_sparse v = View();
    getGpu().Check(dsb200_sparse_transpose(getGpu()._ctx, &v, position, batch, 0, (uint32_t)_vSparseTransposedStart.size(),
                                           _pbSparseTransposedStart->_pDevData, _pbSparseTransposedEnd->_pDevData,
                                           _pbSparseTransposedIndex->_pDevData,
                                           _pbSparseTransposedData ? _pbSparseTransposedData->_pDevData : nullptr),
                   "dsb200_sparse_transpose");
    return true;
}

template <typename T>
bool NNDataSet<T>::CalculateSparseTransposedDenoisedMatrix(uint32_t position, uint32_t batch, NNLayer* pLayer)
{
    if (_bDirty || batch != _batch) GenerateSparseTransposedMatrix(batch, pLayer);     // E/NNTypes.h:602-605
    dsb200_sparse v = View();
    getGpu().Check(dsb200_sparse_transpose(getGpu()._ctx, &v, position, batch, 1, (uint32_t)_vSparseTransposedStart.size(),
                                           _pbSparseTransposedStart->_pDevData, _pbSparseTransposedEnd->_pDevData,
                                           _pbSparseTransposedIndex->_pDevData,
                                           _pbSparseTransposedData ? _pbSparseTransposedData->_pDevData : nullptr),
                   "dsb200_sparse_transpose(denoised)");
    return true;
}

template <typename T>
bool NNDataSet<T>::CalculateSparseTransposedWeightGradient(NNFloat alpha, NNFloat beta, uint32_t m, uint32_t n, NNFloat* pDelta,
                                                           NNFloat* pWeightGradient)
{
    // Boolean unweighted -> index-only kernel, else the analog one (E/NNTypes.h:642-649)
    const bool plain = (_attributes & NNDataSetEnums::Boolean) && !(_attributes & NNDataSetEnums::Weighted);
    getGpu().Check(dsb200_sparse_wgrad(getGpu()._ctx, alpha, beta, m, n, _pbSparseTransposedStart->_pDevData,
                                       _pbSparseTransposedEnd->_pDevData, _pbSparseTransposedIndex->_pDevData,
                                       plain ? nullptr : _pbSparseTransposedData->_pDevData, pDelta, pWeightGradient),
                   "dsb200_sparse_wgrad");
    return true;
}

template <typename T>
bool NNDataSet<T>::CalculateSparseTransposedWeightGradientUpdate(TrainingMode mode, NNFloat galpha, uint32_t m, uint32_t n, NNFloat* pDelta,
                                                                 NNFloat alpha, NNFloat lambda, NNFloat lambda1, NNFloat mu, NNFloat mu1,
                                                                 NNFloat t, NNFloat* pVelocity, NNFloat* pGradientVelocity, NNFloat* pWeight)
{
    const bool plain = (_attributes & NNDataSetEnums::Boolean) && !(_attributes & NNDataSetEnums::Weighted);
    getGpu().Check(dsb200_sparse_wgrad_update(getGpu()._ctx, (int)mode, galpha, m, n, _pbSparseTransposedStart->_pDevData,
                                              _pbSparseTransposedEnd->_pDevData, _pbSparseTransposedIndex->_pDevData,
                                              plain ? nullptr : _pbSparseTransposedData->_pDevData, pDelta, alpha, lambda, lambda1,
                                              mu, mu1, t, pVelocity, pGradientVelocity, pWeight),
                   "dsb200_sparse_wgrad_update");
    return true;
}

template <typename T>
bool NNDataSet<T>::SetDenoising(bool flag)
{
    if (!(_attributes & NNDataSetEnums::Sparse)) return false;                        // E/NNTypes.cpp:1572-1600
    if (!flag) { _pbDenoisingRandom.reset(); _bDenoising = false; }
    else if (!_bDenoising) { _pbDenoisingRandom.reset(new GpuBuffer<NNFloat>(_vSparseIndex.size())); _bDenoising = true; }
    return true;
}

template <typename T>
bool NNDataSet<T>::GenerateDenoisingData()
{
    if (!(_attributes & NNDataSetEnums::Sparse) || !_pbDenoisingRandom) return false;
    // the index array can have been replaced since SetDenoising (LoadSparseData while model parallel swaps in this rank's column
    // shard, which can hold more non-zeros than the one before): the random buffer follows its length
    if (_pbDenoisingRandom->_length < _vSparseIndex.size()) _pbDenoisingRandom.reset(new GpuBuffer<NNFloat>(_vSparseIndex.size()));
    // the reference refills the whole buffer with cuRAND XORWOW once per epoch (E/NNTypes.cpp:1617-1629);
    // here: counter-based generator keyed by (seed, rank, epoch), uniform in (0, 1]
    const uint64_t key = (uint64_t)getGpu()._seed + (uint64_t)getGpu()._id * 76801ull;
    getGpu().Check(dsb200_fill_uniform(getGpu()._ctx, _pbDenoisingRandom->_pDevData, _vSparseIndex.size(), key, _denoisingEpoch++),
                   "dsb200_fill_uniform");
    return true;
}

template <typename T>
bool NNDataSet<T>::CalculateSparseZ(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pWeight, NNFloat* pUnit, NNFloat beta)
{
    dsb200_sparse v = View();
    getGpu().Check(dsb200_sparse_z(getGpu()._ctx, &v, position, batch, stride, pWeight, pUnit, beta, 0), "dsb200_sparse_z");
    return true;
}

template <typename T>
bool NNDataSet<T>::CalculateSparseDenoisedZ(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pWeight, NNFloat* pUnit, NNFloat beta)
{
    dsb200_sparse v = View();
    getGpu().Check(dsb200_sparse_z(getGpu()._ctx, &v, position, batch, stride, pWeight, pUnit, beta, 1), "dsb200_sparse_z(denoised)");
    return true;
}

template <typename T>
bool NNDataSet<T>::CalculateSparseZBiasActivation(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pWeight, NNFloat* pBias,
                                                  Activation activation, NNFloat* pUnit, bool bDenoised)
{
    dsb200_sparse v = View();
    getGpu().Check(dsb200_sparse_z_bias_act(getGpu()._ctx, &v, position, batch, stride, pWeight, pBias, (int)activation, pUnit, bDenoised ? 1 : 0),
                   "dsb200_sparse_z_bias_act");
    return true;
}

template <typename T>
float NNDataSet<T>::SyncError(ErrorFunction ef, Activation activation, uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pUnit)
{
    if (!(_attributes & NNDataSetEnums::Sparse)) throw DsbEngineError("dense targets are outside the hot path");
    dsb200_sparse v = View();
    float loss = 0.0f;
    getGpu().Check(dsb200_sparse_loss(getGpu()._ctx, &v, (int)ef, (int)activation, position, batch, stride, pUnit,
                                      (_attributes & NNDataSetEnums::SparseIgnoreZero) ? 1 : 0, &loss), "dsb200_sparse_loss");
    return loss;
}

template <typename T> float NNDataSet<T>::CalculateL2Error(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pUnit)
{ return SyncError(L2, Sigmoid, position, batch, stride, pUnit); }
template <typename T> float NNDataSet<T>::CalculateCrossEntropyError(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pUnit)
{ return SyncError(CrossEntropy, Sigmoid, position, batch, stride, pUnit); }
template <typename T> float NNDataSet<T>::CalculateScaledMarginalCrossEntropyError(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pUnit)
{ return SyncError(ScaledMarginalCrossEntropy, Sigmoid, position, batch, stride, pUnit); }
template <typename T> float NNDataSet<T>::CalculateMultinomialCrossEntropyError(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pUnit)
{ return SyncError(CrossEntropy, SoftMax, position, batch, stride, pUnit); }
template <typename T> float NNDataSet<T>::CalculateMultinomialScaledMarginalCrossEntropyError(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pUnit)
{ return SyncError(ScaledMarginalCrossEntropy, SoftMax, position, batch, stride, pUnit); }

template <typename T>
bool NNDataSet<T>::CalculateErrorAsync(ErrorFunction ef, Activation activation, uint32_t position, uint32_t batch, uint32_t stride,
                                       NNFloat* pUnit, unsigned long long* pDevAccumulator)
{
    dsb200_sparse v = View();
    getGpu().Check(dsb200_sparse_loss_async(getGpu()._ctx, &v, (int)ef, (int)activation, position, batch, stride, pUnit,
                                            (_attributes & NNDataSetEnums::SparseIgnoreZero) ? 1 : 0, pDevAccumulator),
                   "dsb200_sparse_loss_async");
    return true;
}

template <typename T>
bool NNDataSet<T>::CalculateFusedOutput(ErrorFunction ef, Activation activation, uint32_t position, uint32_t batch, uint32_t stride,
                                        NNFloat* pUnit, NNFloat* pDelta, unsigned long long* pDevAccumulator, bool writeUnits)
{
    dsb200_sparse v = View();
    getGpu().Check(dsb200_output_pass(getGpu()._ctx, &v, (int)ef, (int)activation, position, batch, stride, pUnit, writeUnits ? pUnit : NULL, pDelta,
                                      pDevAccumulator),
                   "dsb200_output_pass");
    return true;
}

template <typename T>
bool NNDataSet<T>::CalculateCrossEntropyOutputDelta(Activation activation, uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pUnit, NNFloat* pDelta)
{
    dsb200_sparse v = View();
    getGpu().Check(dsb200_sparse_output_delta(getGpu()._ctx, &v, DSB200_ERR_CROSS_ENTROPY, (int)activation, position, batch, stride, pUnit, pDelta,
                                              (_attributes & NNDataSetEnums::SparseIgnoreZero) ? 1 : 0, 0.0f, 0.0f, 0.0f), "dsb200_sparse_output_delta(CE)");
    return true;
}

template <typename T>
bool NNDataSet<T>::CalculateScaledMarginalCrossEntropyOutputDelta(Activation activation, uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pUnit, NNFloat* pDelta)
{
    dsb200_sparse v = View();
    getGpu().Check(dsb200_sparse_output_delta(getGpu()._ctx, &v, DSB200_ERR_SMCE, (int)activation, position, batch, stride, pUnit, pDelta,
                                              (_attributes & NNDataSetEnums::SparseIgnoreZero) ? 1 : 0, 0.0f, 0.0f, 0.0f), "dsb200_sparse_output_delta(SMCE)");
    return true;
}

template <typename T>
bool NNDataSet<T>::CalculateOutputDelta(Activation activation, uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pUnit, NNFloat* pDelta,
                                        NNFloat slope, NNFloat alpha, NNFloat lambda)
{
    dsb200_sparse v = View();
    getGpu().Check(dsb200_sparse_output_delta(getGpu()._ctx, &v, DSB200_ERR_L2, (int)activation, position, batch, stride, pUnit, pDelta,
                                              (_attributes & NNDataSetEnums::SparseIgnoreZero) ? 1 : 0, slope, alpha, lambda), "dsb200_sparse_output_delta(L2)");
    return true;
}

template <typename T>
bool NNDataSet<T>::SaveNetCDF(const string& fname)
{
    vector<NNDataSetBase*> v(1, this);
    return ::SaveNetCDF(fname, v);
}

NNDataSetBase* createNNDataSet(const NNDataSetDescriptor& d)
{
    using namespace NNDataSetEnums;
    if (!NNDataSetDescriptor::isSupported(d._attributes)) {
        stringstream msg; msg << "Unsupported attributes " << d._attributes << " for dataset " << d._name << " (sparse datasets only on this path)";
        throw std::runtime_error(msg.str());
    }
    const bool weighted = d._attributes & Weighted;
    NNDataSetBase* p = nullptr;
    switch (d._dataType) {
    case UInt:   p = new NNDataSet<uint32_t>(d._examples, d._sparseDensity, d._dim, weighted, d._name); break;
    case Int:    p = new NNDataSet<int32_t>(d._examples, d._sparseDensity, d._dim, weighted, d._name); break;
    case Float:  p = new NNDataSet<float>(d._examples, d._sparseDensity, d._dim, weighted, d._name); break;
    case Double: p = new NNDataSet<double>(d._examples, d._sparseDensity, d._dim, weighted, d._name); break;
    case Char:   p = new NNDataSet<char>(d._examples, d._sparseDensity, d._dim, weighted, d._name); break;
    case UChar:  p = new NNDataSet<unsigned char>(d._examples, d._sparseDensity, d._dim, weighted, d._name); break;
    default: { stringstream msg; msg << "Unsupported data type: " << d._dataType; throw std::runtime_error(msg.str()); }
    }
    if (d._attributes & Boolean) p->_attributes |= Boolean;
    if (d._attributes & SparseIgnoreZero) p->_attributes |= SparseIgnoreZero;
    return p;
}

template class NNDataSet<float>;
template class NNDataSet<double>;
template class NNDataSet<unsigned char>;
template class NNDataSet<char>;
template class NNDataSet<uint32_t>;
template class NNDataSet<uint64_t>;
template class NNDataSet<int32_t>;
template class NNDataSet<int64_t>;
