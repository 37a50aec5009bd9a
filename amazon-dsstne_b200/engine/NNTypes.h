// NNTypes.h -- enums and the NNDataSet<T> API of the reference (E/NNTypes.h, E/NNEnum.h),
// re-hosted on the dsstne_b200 C ABI for the sparse fully-connected path.
//
// Kept: names, argument meaning and error behaviour of NNDataSetBase / NNDataSet<T>
// (E/NNTypes.h:197-476), the attribute / dataType enums (E/NNEnum.h), the sparse layout
// (E/NNTypes.h:213-236) and the dispatch from dataset attributes to kernel variants
// (E/NNTypes.h:478-1251).  Out of scope: dense/image datasets, L1 / Hinge /
// L2Hinge / DataScaled errors -- those entry points throw DsbEngineError instead of computing.
#pragma once

#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "GpuTypes.h"

using std::string;
using std::vector;
using std::tuple;
using std::unique_ptr;

class NNDataSetBase;
class NNLayer;
class NNNetwork;
class NNWeight;

static const float NN_VERSION     = 0.9f;
static const float MIN_ERROR      = 1.0e-12f;
static const float MIN_ACTIVATION = 0.000001f;
static const float MAX_ACTIVATION = 0.999999f;
static const float MAX_VALUE      = 999999999999999.0f;

enum { DefaultBatch = 512 };

enum Mode { Prediction = 0, Training = 1, Validation = 2, Unspecified = 3 };

enum TrainingMode { SGD = 0, Momentum = 1, AdaGrad = 2, Nesterov = 3, RMSProp = 4, AdaDelta = 5, Adam = 6 };

enum ErrorFunction { L1, L2, CrossEntropy, ScaledMarginalCrossEntropy, DataScaledMarginalCrossEntropy, Hinge, L2Hinge };

enum Activation {
    Sigmoid, Tanh, RectifiedLinear, Linear, ParametricRectifiedLinear, SoftPlus, SoftSign, SoftMax, RELUMax, LinearMax,
    ExponentialLinear, LeakyRectifiedLinear, ScaledExponentialLinear
};

enum WeightInitialization { Xavier, CaffeXavier, Gaussian, Uniform, UnitBall, Constant, SELU };

enum PoolingFunction { None, Max, Average, LRN, Maxout, DotProduct, Cosine, Stochastic, LCN, GlobalTemporal };

namespace NNDataSetEnums {
enum Attributes {
    Sparse = 1, Boolean = 2, Compressed = 4, Recurrent = 8, Mutable = 16, SparseIgnoreZero = 32, Indexed = 64, Weighted = 128
};
enum Kind { Numeric = 0, Image = 1, Audio = 2 };
enum Sharding { None = 0, Model = 1, Data = 2 };
enum DataType { UInt = 0, Int = 1, LLInt = 2, ULLInt = 3, Float = 4, Double = 5, RGB8 = 6, RGB16 = 7, UChar = 8, Char = 9 };

template <typename T> inline DataType getDataType() { throw std::runtime_error("Default data type not defined"); }
template <> inline DataType getDataType<uint32_t>() { return UInt; }
template <> inline DataType getDataType<int32_t>() { return Int; }
template <> inline DataType getDataType<int64_t>() { return LLInt; }
template <> inline DataType getDataType<uint64_t>() { return ULLInt; }
template <> inline DataType getDataType<float>() { return Float; }
template <> inline DataType getDataType<double>() { return Double; }
template <> inline DataType getDataType<char>() { return Char; }
template <> inline DataType getDataType<unsigned char>() { return UChar; }
}  // namespace NNDataSetEnums

struct NNDataSetDimensions {
    uint32_t _dimensions, _width, _height, _length;
    NNDataSetDimensions() : _dimensions(1), _width(1), _height(1), _length(1) {}
    NNDataSetDimensions(uint32_t width, uint32_t height = 1, uint32_t length = 1)
        : _dimensions((width > 1) + (height > 1) + (length > 1)), _width(width), _height(height), _length(length) {}
};

struct NNDataSetDescriptor {
    string _name;
    NNDataSetEnums::DataType _dataType;
    uint32_t _attributes;
    NNDataSetDimensions _dim;
    uint32_t _examples;
    float _sparseDensity;
    static bool isSupported(uint32_t attributes)
    {
        const uint32_t ok = NNDataSetEnums::Sparse | NNDataSetEnums::Boolean | NNDataSetEnums::Indexed |
                            NNDataSetEnums::Weighted | NNDataSetEnums::SparseIgnoreZero;
        return (attributes & NNDataSetEnums::Sparse) && !(attributes & ~ok);
    }
};

NNDataSetBase* createNNDataSet(const NNDataSetDescriptor& descriptor);

struct NNDataSetBase {
    string                          _name;
    NNDataSetEnums::DataType        _dataType;
    uint32_t                        _attributes;
    uint32_t                        _examples;
    uint32_t                        _uniqueExamples;
    uint32_t                        _localExamples;
    uint32_t                        _dimensions, _width, _height, _length, _stride;
    NNDataSetEnums::Sharding        _sharding;
    uint32_t                        _minX, _maxX;          // local column range when model sharded
    uint64_t                        _sparseDataSize;
    NNFloat                         _sparseDensity;
    vector<uint64_t>                _vSparseStart;
    unique_ptr<GpuBuffer<uint64_t>> _pbSparseStart;
    vector<uint64_t>                _vSparseEnd;
    unique_ptr<GpuBuffer<uint64_t>> _pbSparseEnd;
    vector<uint32_t>                _vSparseIndex;
    unique_ptr<GpuBuffer<uint32_t>> _pbSparseIndex;
    vector<NNFloat>                 _vDataWeight;
    unique_ptr<GpuBuffer<NNFloat>>  _pbDataWeight;
    vector<uint32_t>                _vIndex;
    unique_ptr<GpuBuffer<uint32_t>> _pbIndex;
    unique_ptr<GpuBuffer<NNFloat>>  _pbDenoisingRandom;
    vector<uint64_t>                _vSparseDatapointCount;
    vector<uint32_t>                _vSparseMaxDatapointCount;
    vector<uint32_t>                _vSparseMultiDatapointCount;
    vector<uint32_t>                _vSparseTransposedStart;
    uint64_t                        _sparseTransposedIndices;
    unique_ptr<GpuBuffer<uint32_t>> _pbSparseTransposedStart;
    unique_ptr<GpuBuffer<uint32_t>> _pbSparseTransposedEnd;
    unique_ptr<GpuBuffer<uint32_t>> _pbSparseTransposedIndex;
    unique_ptr<GpuBuffer<NNFloat>>  _pbSparseTransposedData;
    bool                            _bDenoising;
    bool                            _bDirty;
    bool                            _bStreaming;
    bool                            _bIndexed;
    uint32_t                        _batch;
    uint64_t                        _denoisingEpoch;       // counter-based RNG stream id (see GenerateDenoisingData)
    uint32_t                        _maxBatchNnz;          // largest nnz of any contiguous `batch` rows (workspace sizing)
    vector<uint64_t>                _vFullSparseStart, _vFullSparseEnd;    // un-sharded host copy while model sharded
    vector<uint32_t>                _vFullSparseIndex;

    NNDataSetBase();
    NNDataSetBase(const string& name, NNDataSetEnums::DataType dataType, uint32_t examples, uint32_t uniqueExamples,
                  const NNDataSetDimensions& datasetDim);
    virtual ~NNDataSetBase() {}
    NNDataSetDimensions GetDimensions() { return NNDataSetDimensions(_width, _height, _length); }
    uint32_t GetExamples() { return _examples; }
    uint32_t GetUniqueExamples() { return _uniqueExamples; }

    virtual bool SaveNetCDF(const string& fname) = 0;
    virtual void WaitForUpload(cudaStream_t stream) = 0;      // a step that reads this data set waits for a pending copy-stream upload
    virtual void RefreshState(uint32_t batch) = 0;
    virtual bool Shard(NNDataSetEnums::Sharding sharding) = 0;
    virtual bool UnShard() = 0;
    virtual bool CalculateSparseDatapointCounts() = 0;
    virtual bool GenerateSparseTransposedMatrix(uint32_t batch, NNLayer* pLayer) = 0;
    virtual bool CalculateSparseTransposedMatrix(uint32_t position, uint32_t batch, NNLayer* pLayer) = 0;
    virtual bool CalculateSparseTransposedDenoisedMatrix(uint32_t position, uint32_t batch, NNLayer* pLayer) = 0;
    virtual bool CalculateSparseTransposedWeightGradient(NNFloat alpha, NNFloat beta, uint32_t m, uint32_t n, NNFloat* pDelta, NNFloat* pWeightGradient) = 0;
    virtual bool SetDenoising(bool flag) = 0;
    virtual bool GenerateDenoisingData() = 0;
    virtual bool CalculateSparseZ(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pWeight, NNFloat* pUnit, NNFloat beta = (NNFloat)0.0) = 0;
    virtual bool CalculateSparseDenoisedZ(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pWeight, NNFloat* pUnit, NNFloat beta = (NNFloat)0.0) = 0;
    virtual float CalculateL2Error(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pUnit) = 0;
    virtual float CalculateCrossEntropyError(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pUnit) = 0;
    virtual float CalculateScaledMarginalCrossEntropyError(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pUnit) = 0;
    virtual float CalculateMultinomialCrossEntropyError(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pUnit) = 0;
    virtual float CalculateMultinomialScaledMarginalCrossEntropyError(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pUnit) = 0;
    virtual bool CalculateCrossEntropyOutputDelta(Activation activation, uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pUnit, NNFloat* pDelta) = 0;
    virtual bool CalculateScaledMarginalCrossEntropyOutputDelta(Activation activation, uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pUnit, NNFloat* pDelta) = 0;
    virtual bool CalculateOutputDelta(Activation activation, uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pUnit, NNFloat* pDelta, NNFloat slope, NNFloat alpha, NNFloat lambda) = 0;
    virtual void LoadSparseData(const uint64_t* srcSparseStart, const uint64_t* srcSparseEnd, const void* srcSparseData, const uint32_t* srcSparseIndex) = 0;
    virtual void CopySparseData(const uint64_t* srcSparseStart, const uint64_t* srcSparseEnd, const void* srcSparseData, const uint32_t* srcSparseIndex) = 0;
    virtual void LoadSparseData(const long* srcSparseStart, const long* srcSparseEnd, const void* srcSparseData, const long* srcSparseIndex) = 0;
    virtual void CopySparseData(const long* srcSparseStart, const long* srcSparseEnd, const void* srcSparseData, const long* srcSparseIndex) = 0;
    virtual void LoadIndexedData(const uint32_t* srcIndexedData) = 0;
    virtual void LoadDataWeight(const NNFloat* srcWeightData) = 0;

    // ---- B200 additions (not in the reference API) ----
    // device view of this dataset for the C ABI
    virtual dsb200_sparse View() = 0;
    // fused activation + loss + delta for a sparse-target output layer (dsb200_output_pass)
    // writeUnits == false: the activations are not stored (training needs only the delta); pUnit keeps Z
    virtual bool CalculateFusedOutput(ErrorFunction ef, Activation activation, uint32_t position, uint32_t batch, uint32_t stride,
                                      NNFloat* pUnit, NNFloat* pDelta, unsigned long long* pDevAccumulator, bool writeUnits = true) = 0;
    // asynchronous loss (fixed-point accumulate on the device, no host sync)
    virtual bool CalculateErrorAsync(ErrorFunction ef, Activation activation, uint32_t position, uint32_t batch, uint32_t stride,
                                     NNFloat* pUnit, unsigned long long* pDevAccumulator) = 0;
    // fused forward of the sparse input layer: bias + sparse Z + activation
    virtual bool CalculateSparseZBiasActivation(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pWeight, NNFloat* pBias,
                                                Activation activation, NNFloat* pUnit, bool bDenoised) = 0;
    // fused sparse gradient + optimizer step (dsb200_sparse_wgrad_update)
    virtual bool CalculateSparseTransposedWeightGradientUpdate(TrainingMode mode, NNFloat galpha, uint32_t m, uint32_t n, NNFloat* pDelta,
                                                               NNFloat alpha, NNFloat lambda, NNFloat lambda1, NNFloat mu, NNFloat mu1, NNFloat t,
                                                               NNFloat* pVelocity, NNFloat* pGradientVelocity, NNFloat* pWeight) = 0;
};

template <typename T>
class NNDataSet : public NNDataSetBase {
public:
    friend class NNNetwork;
    friend class NNLayer;
    vector<T>                _vSparseData;
    vector<T>                _vFullSparseData;
    unique_ptr<GpuBuffer<T>> _pbSparseData;

    // sparse dataset with room for examples * sparseDensity * stride data points (E/NNTypes.cpp:667-701)
    NNDataSet(uint32_t examples, NNFloat sparseDensity, const NNDataSetDimensions& dim, bool isWeighted = false, const string& name = "");
    // sparse indexed dataset (E/NNTypes.cpp:703-741)
    NNDataSet(uint32_t examples, uint32_t uniqueExamples, size_t sparseDataSize, const NNDataSetDimensions& dim,
              bool isIndexed = false, bool isWeighted = false, const string& name = "");

    bool SaveNetCDF(const string& fname);
    void RefreshState(uint32_t batch) { (void)batch; }
    bool Shard(NNDataSetEnums::Sharding sharding);
    bool UnShard();
    bool CalculateSparseDatapointCounts();
    bool GenerateSparseTransposedMatrix(uint32_t batch, NNLayer* pLayer);
    bool CalculateSparseTransposedMatrix(uint32_t position, uint32_t batch, NNLayer* pLayer);
    bool CalculateSparseTransposedDenoisedMatrix(uint32_t position, uint32_t batch, NNLayer* pLayer);
    bool CalculateSparseTransposedWeightGradient(NNFloat alpha, NNFloat beta, uint32_t m, uint32_t n, NNFloat* pDelta, NNFloat* pWeightGradient);
    bool SetDenoising(bool flag);
    bool GenerateDenoisingData();
    bool CalculateSparseZ(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pWeight, NNFloat* pUnit, NNFloat beta);
    bool CalculateSparseDenoisedZ(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pWeight, NNFloat* pUnit, NNFloat beta);
    float CalculateL2Error(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pUnit);
    float CalculateCrossEntropyError(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pUnit);
    float CalculateScaledMarginalCrossEntropyError(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pUnit);
    float CalculateMultinomialCrossEntropyError(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pUnit);
    float CalculateMultinomialScaledMarginalCrossEntropyError(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pUnit);
    bool CalculateCrossEntropyOutputDelta(Activation activation, uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pUnit, NNFloat* pDelta);
    bool CalculateScaledMarginalCrossEntropyOutputDelta(Activation activation, uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pUnit, NNFloat* pDelta);
    bool CalculateOutputDelta(Activation activation, uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pUnit, NNFloat* pDelta, NNFloat slope, NNFloat alpha, NNFloat lambda);
    void LoadSparseData(const uint64_t* srcSparseStart, const uint64_t* srcSparseEnd, const void* srcSparseData, const uint32_t* srcSparseIndex);
    void CopySparseData(const uint64_t* srcSparseStart, const uint64_t* srcSparseEnd, const void* srcSparseData, const uint32_t* srcSparseIndex);
    void LoadSparseData(const long* srcSparseStart, const long* srcSparseEnd, const void* srcSparseData, const long* srcSparseIndex);
    void CopySparseData(const long* srcSparseStart, const long* srcSparseEnd, const void* srcSparseData, const long* srcSparseIndex);
    void LoadIndexedData(const uint32_t* srcIndexedData);
    void LoadDataWeight(const NNFloat* srcWeightData);

    dsb200_sparse View();
    bool CalculateFusedOutput(ErrorFunction ef, Activation activation, uint32_t position, uint32_t batch, uint32_t stride,
                              NNFloat* pUnit, NNFloat* pDelta, unsigned long long* pDevAccumulator, bool writeUnits = true);
    bool CalculateErrorAsync(ErrorFunction ef, Activation activation, uint32_t position, uint32_t batch, uint32_t stride,
                             NNFloat* pUnit, unsigned long long* pDevAccumulator);
    bool CalculateSparseZBiasActivation(uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pWeight, NNFloat* pBias,
                                        Activation activation, NNFloat* pUnit, bool bDenoised);
    bool CalculateSparseTransposedWeightGradientUpdate(TrainingMode mode, NNFloat galpha, uint32_t m, uint32_t n, NNFloat* pDelta,
                                                       NNFloat alpha, NNFloat lambda, NNFloat lambda1, NNFloat mu, NNFloat mu1, NNFloat t,
                                                       NNFloat* pVelocity, NNFloat* pGradientVelocity, NNFloat* pWeight);

private:
    void UploadSparse();
    void SliceFromFull();
    // streaming path of LoadSparseData: pinned double-buffered staging + asynchronous copies of the used part only
    void UploadSparseAsync(const uint64_t* srcStart, const uint64_t* srcEnd, const void* srcData, const uint32_t* srcIndex, uint64_t dataLength);
    struct Staging {
        uint64_t* start = nullptr; uint64_t* end = nullptr; uint32_t* index = nullptr; T* data = nullptr;
        cudaEvent_t done = nullptr; bool pending = false; size_t capacity = 0;   // capacity: entries of index / data
    } _staging[2];
    int _stagingCur = 0;
    cudaEvent_t _uploadEvent = nullptr;                       // recorded on the copy stream: readers of this step wait for it (WaitForUpload)
    // single-copy path (engine option "pinned_mirror", default on): the host mirror itself is page-locked and is the copy source
    struct Mirror {
        void* ptr[4] = {nullptr, nullptr, nullptr, nullptr};      // registered storage of _vSparseStart / End / Index / Data
        cudaEvent_t done = nullptr; bool pending = false; bool unavailable = false;
    } _mirror;
    bool UploadMirrorAsync(uint64_t dataLength);              // false: page-locking failed, the caller takes the staging path
    cudaStream_t BeginUpload();                               // the stream this upload goes to (copy stream when the last step's readers are known)
    void EndUpload(cudaStream_t stream, cudaEvent_t done);
    unique_ptr<GpuBuffer<uint32_t>> _pbColumnCount;        // scratch of the device-side capacity table
public:
    ~NNDataSet();
    void WaitForUpload(cudaStream_t stream);
private:
    float SyncError(ErrorFunction ef, Activation activation, uint32_t position, uint32_t batch, uint32_t stride, NNFloat* pUnit);
};

// NetCDF dataset files (classic CDF-1 / CDF-2 / CDF-5), E/NNTypes.cpp:2456-2584 -- see NetCDF.h, NetCDFIO.cpp
vector<NNDataSetBase*> LoadNetCDF(const string& fname);
bool SaveNetCDF(const string& fname, vector<NNDataSetBase*> vDataSet);
