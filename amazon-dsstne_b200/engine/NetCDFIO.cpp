// NetCDFIO.cpp -- DSSTNE dataset and network files ("next" row 1 of SURVEY 8f) on the classic-format reader / writer
// of NetCDF.h.
//
// Dataset schema (read: E/NNTypes.cpp:1081-1418; written by generateNetCDF: U/NetCDFhelper.cpp:332-416):
//   global attributes  datasets, name<n>, attributes<n>, kind<n>, dataType<n>, dimensions<n>, width<n>[, height<n>, length<n>]
//   dimensions         examplesDim<n>[, uniqueExamplesDim<n>], sparseDataDim<n>
//   variables          sparseStart<n>, sparseEnd<n> (uint or uint64), sparseIndex<n> (uint), [sparseData<n>], [dataWeight<n>], [index<n>]
// The reference's own NNDataSet::SaveNetCDF spells the type attribute `datatype<n>` (E/NNTypes.cpp:2247) while its reader
// wants `dataType<n>` (:1107); this reader accepts both and this writer emits the spelling the reader needs.
// Network schema: E/NNNetwork.cpp:1936-1970 (network), E/NNLayer.cpp:3447-3525 (layer<i>_*), E/NNWeight.cpp:853-900
// (weight<i>_*, variables weight<i>_bias / weight<i>_weights as full row-major [input][output]).
#include "NNNetwork.h"
#include "NetCDF.h"

using namespace NNDataSetEnums;

namespace {

const nc::Att& need_att(const nc::File& f, const string& name, const string& fname)
{
    const nc::Att* a = f.att(name);
    if (!a) throw DsbEngineError("NetCDF: no attribute " + name + " in " + fname);
    return *a;
}
const nc::Var& need_var(const nc::File& f, const string& name, const string& fname)
{
    const nc::Var* v = f.var(name);
    if (!v) throw DsbEngineError("NetCDF: no variable " + name + " in " + fname);
    return *v;
}
uint64_t need_dim(const nc::File& f, const string& name, const string& fname)
{
    const nc::Dim* d = f.dim(name);
    if (!d) throw DsbEngineError("NetCDF: no dimension " + name + " in " + fname);
    return d->size;
}
float att_f(const nc::File& f, const string& name, float dflt)
{
    const nc::Att* a = f.att(name);
    return a ? (float)a->as_double() : dflt;
}
int64_t att_i(const nc::File& f, const string& name, int64_t dflt)
{
    const nc::Att* a = f.att(name);
    return a ? a->as_int() : dflt;
}
string att_s(const nc::File& f, const string& name, const string& dflt)
{
    const nc::Att* a = f.att(name);
    return a ? a->as_string() : dflt;
}

template <typename T>
NNDataSetBase* load_one(const nc::File& f, const string& fname, uint32_t n, uint32_t attributes)
{
    const string ns = std::to_string(n);
    const string name = need_att(f, "name" + ns, fname).as_string();
    const uint32_t examples = (uint32_t)need_dim(f, "examplesDim" + ns, fname);
    const nc::Dim* ud = f.dim("uniqueExamplesDim" + ns);
    const uint32_t uniqueExamples = ud ? (uint32_t)ud->size : examples;
    const uint32_t dimensions = (uint32_t)need_att(f, "dimensions" + ns, fname).as_int();
    if (dimensions < 1 || dimensions > 3) throw DsbEngineError("NetCDF: invalid dimension count in " + fname);
    const uint32_t width = (uint32_t)need_att(f, "width" + ns, fname).as_int();
    const uint32_t height = dimensions > 1 ? (uint32_t)need_att(f, "height" + ns, fname).as_int() : 1;
    const uint32_t length = dimensions > 2 ? (uint32_t)need_att(f, "length" + ns, fname).as_int() : 1;
    if (!width || !height || !length) throw DsbEngineError("NetCDF: zero-sized dimension in " + fname);
    if (!NNDataSetDescriptor::isSupported(attributes))
        throw DsbEngineError("NetCDF: dataset " + name + " in " + fname + " is not a sparse dataset (dense / image data is outside the dsstne_b200 hot path)");
    const uint64_t nnz = need_dim(f, "sparseDataDim" + ns, fname);
    vector<uint64_t> vStart, vEnd;
    vector<uint32_t> vIndex;
    f.read(need_var(f, "sparseStart" + ns, fname), vStart);                     // uint or uint64 (E/NNTypes.cpp:1258-1277), any integer type here
    f.read(need_var(f, "sparseEnd" + ns, fname), vEnd);
    f.read(need_var(f, "sparseIndex" + ns, fname), vIndex);
    if (vStart.size() != uniqueExamples || vEnd.size() != uniqueExamples || vIndex.size() != nnz)
        throw DsbEngineError("NetCDF: sparse variable sizes do not match the dimensions in " + fname);
    for (uint32_t i = 0; i < uniqueExamples; i++)
        if (vStart[i] > vEnd[i] || vEnd[i] > nnz) throw DsbEngineError("NetCDF: sparseStart/sparseEnd out of range in " + fname);
    for (uint64_t i = 0; i < nnz; i++)
        if (vIndex[i] >= (uint64_t)width * height * length) throw DsbEngineError("NetCDF: sparseIndex beyond the dataset width in " + fname);
    vector<T> vData;
    if (!(attributes & Boolean)) {
        f.read(need_var(f, "sparseData" + ns, fname), vData);
        if (vData.size() != nnz) throw DsbEngineError("NetCDF: sparseData size does not match sparseDataDim in " + fname);
    }
    vector<NNFloat> vWeight;
    if (attributes & Weighted) {
        f.read(need_var(f, "dataWeight" + ns, fname), vWeight);
        if (vWeight.size() < uniqueExamples) throw DsbEngineError("NetCDF: dataWeight is shorter than the examples in " + fname);
        vWeight.resize(examples);
    }
    vector<uint32_t> vEx;
    if (attributes & Indexed) {
        f.read(need_var(f, "index" + ns, fname), vEx);
        if (vEx.size() != examples) throw DsbEngineError("NetCDF: index size does not match examplesDim in " + fname);
        for (uint32_t e : vEx)
            if (e >= uniqueExamples) throw DsbEngineError("NetCDF: index refers to a missing unique example in " + fname);
    }
    NNDataSet<T>* p = new NNDataSet<T>(examples, uniqueExamples, (size_t)nnz, NNDataSetDimensions(width, height, length), (attributes & Indexed) != 0,
                                       (attributes & Weighted) != 0, name);
    p->LoadSparseData(vStart.data(), vEnd.data(), (attributes & Boolean) ? NULL : vData.data(), vIndex.data());
    if (attributes & Indexed) p->LoadIndexedData(vEx.data());
    if (attributes & Weighted) p->LoadDataWeight(vWeight.data());
    p->_attributes = attributes;
    return p;
}

template <typename T> nc::Type nc_type_of();
template <> nc::Type nc_type_of<uint32_t>() { return nc::NC_UINT; }
template <> nc::Type nc_type_of<int32_t>() { return nc::NC_INT; }
template <> nc::Type nc_type_of<float>() { return nc::NC_FLOAT; }
template <> nc::Type nc_type_of<double>() { return nc::NC_DOUBLE; }
template <> nc::Type nc_type_of<char>() { return nc::NC_BYTE; }
template <> nc::Type nc_type_of<unsigned char>() { return nc::NC_UBYTE; }

// keeps converted copies alive until Writer::write
struct Keep {
    std::vector<std::vector<uint32_t>> u32;
    std::vector<std::vector<uint64_t>> u64;
};

template <typename T>
void save_one(nc::Writer& w, Keep& keep, NNDataSetBase* base, uint32_t n)
{
    NNDataSet<T>* p = static_cast<NNDataSet<T>*>(base);
    const bool sharded = p->_sharding == Model && !p->_vFullSparseStart.empty();    // column-sharded over several ranks: write the full copy
    const vector<uint64_t>& vStart = sharded ? p->_vFullSparseStart : p->_vSparseStart;
    const vector<uint64_t>& vEnd = sharded ? p->_vFullSparseEnd : p->_vSparseEnd;
    const vector<uint32_t>& vIndex = sharded ? p->_vFullSparseIndex : p->_vSparseIndex;
    const vector<T>& vData = sharded ? p->_vFullSparseData : p->_vSparseData;
    const string ns = std::to_string(n);
    w.put_att("name" + ns, p->_name);
    w.put_att("attributes" + ns, nc::NC_UINT, p->_attributes);
    w.put_att("kind" + ns, nc::NC_UINT, Numeric);
    w.put_att("dataType" + ns, nc::NC_UINT, p->_dataType);
    w.put_att("dimensions" + ns, nc::NC_UINT, p->_dimensions);
    w.put_att("width" + ns, nc::NC_UINT, p->_width);
    if (p->_dimensions > 1) w.put_att("height" + ns, nc::NC_UINT, p->_height);
    if (p->_dimensions > 2) w.put_att("length" + ns, nc::NC_UINT, p->_length);
    if (p->_uniqueExamples != p->_examples || (p->_attributes & Indexed)) w.add_dim("uniqueExamplesDim" + ns, p->_uniqueExamples);
    w.add_dim("examplesDim" + ns, p->_examples);
    const string rowDim = (p->_uniqueExamples != p->_examples || (p->_attributes & Indexed)) ? "uniqueExamplesDim" + ns : "examplesDim" + ns;
    const uint64_t nnz = p->_uniqueExamples ? vEnd[p->_uniqueExamples - 1] : 0;
    if (nnz == 0) throw DsbEngineError("SaveNetCDF: dataset " + p->_name + " has no data points");
    w.add_dim("sparseDataDim" + ns, nnz);
    if (nnz <= 0xffffffffull) {                                                  // what generateNetCDF writes: uint offsets
        keep.u32.emplace_back(vStart.begin(), vStart.begin() + p->_uniqueExamples);
        w.add_var("sparseStart" + ns, nc::NC_UINT, rowDim, keep.u32.back().data());
        keep.u32.emplace_back(vEnd.begin(), vEnd.begin() + p->_uniqueExamples);
        w.add_var("sparseEnd" + ns, nc::NC_UINT, rowDim, keep.u32.back().data());
    } else {
        w.add_var("sparseStart" + ns, nc::NC_UINT64, rowDim, vStart.data());
        w.add_var("sparseEnd" + ns, nc::NC_UINT64, rowDim, vEnd.data());
    }
    w.add_var("sparseIndex" + ns, nc::NC_UINT, "sparseDataDim" + ns, vIndex.data());
    if (!(p->_attributes & Boolean)) w.add_var("sparseData" + ns, nc_type_of<T>(), "sparseDataDim" + ns, vData.data());
    if (p->_attributes & Weighted) w.add_var("dataWeight" + ns, nc::NC_FLOAT, rowDim, p->_vDataWeight.data());
    if (p->_attributes & Indexed) w.add_var("index" + ns, nc::NC_UINT, "examplesDim" + ns, p->_vIndex.data());
}

}  // namespace

// LoadNetCDF (E/NNTypes.cpp:2456-2584): every dataset of the file, typed by its dataType<n>
vector<NNDataSetBase*> LoadNetCDF(const string& fname)
{
    vector<NNDataSetBase*> v;
    try {
        nc::File f(fname);
        const uint32_t datasets = (uint32_t)need_att(f, "datasets", fname).as_int();
        for (uint32_t i = 0; i < datasets; i++) {
            const string ns = std::to_string(i);
            const nc::Att* dt = f.att("dataType" + ns);
            if (!dt) dt = f.att("datatype" + ns);
            if (!dt) throw DsbEngineError("LoadNetCDF: No datatype supplied in NetCDF input file " + fname);
            const uint32_t attributes = (uint32_t)need_att(f, "attributes" + ns, fname).as_int();
            switch ((DataType)dt->as_int()) {
            case UInt:   v.push_back(load_one<uint32_t>(f, fname, i, attributes)); break;
            case Int:    v.push_back(load_one<int32_t>(f, fname, i, attributes)); break;
            case Float:  v.push_back(load_one<float>(f, fname, i, attributes)); break;
            case Double: v.push_back(load_one<double>(f, fname, i, attributes)); break;
            case Char:   v.push_back(load_one<char>(f, fname, i, attributes)); break;
            case UChar:  v.push_back(load_one<unsigned char>(f, fname, i, attributes)); break;
            default: throw DsbEngineError("LoadNetCDF: unsupported data type " + std::to_string(dt->as_int()) + " in " + fname);
            }
            if (getGpu()._id == 0)
                printf("LoadNetCDF: dataset %s: %u examples, width %u, %llu data points\n", v.back()->_name.c_str(), v.back()->_examples, v.back()->_width,
                       (unsigned long long)v.back()->_sparseDataSize);
        }
    } catch (const nc::Error& e) {
        for (auto p : v) delete p;
        throw DsbEngineError(string("LoadNetCDF: ") + e.what());
    } catch (...) {
        for (auto p : v) delete p;
        throw;
    }
    return v;
}

bool SaveNetCDF(const string& fname, vector<NNDataSetBase*> vDataSet)
{
    if (getGpu()._id != 0) return true;                                          // rank 0 writes (E/NNTypes.cpp:2386-2453)
    try {
        nc::Writer w(5);
        Keep keep;
        w.put_att("datasets", nc::NC_UINT, (double)vDataSet.size());
        for (uint32_t i = 0; i < vDataSet.size(); i++) {
            switch (vDataSet[i]->_dataType) {
            case UInt:   save_one<uint32_t>(w, keep, vDataSet[i], i); break;
            case Int:    save_one<int32_t>(w, keep, vDataSet[i], i); break;
            case Float:  save_one<float>(w, keep, vDataSet[i], i); break;
            case Double: save_one<double>(w, keep, vDataSet[i], i); break;
            case Char:   save_one<char>(w, keep, vDataSet[i], i); break;
            case UChar:  save_one<unsigned char>(w, keep, vDataSet[i], i); break;
            default: throw DsbEngineError("SaveNetCDF: unsupported data type");
            }
        }
        w.write(fname);
    } catch (const nc::Error& e) {
        throw DsbEngineError(string("SaveNetCDF: ") + e.what());
    }
    return true;
}

// ---------------------------------------------------------------- network files

// full (un-sharded) row-major [input][output] weights and biases on every rank
static void gather_full(NNWeight* w, uint64_t in, uint64_t out, bool outgoingLarger, vector<NNFloat>& vW, vector<NNFloat>& vB)
{
    vector<NNFloat> lw, lb;
    w->GetWeights(lw);
    w->GetBiases(lb);
    if (getGpu()._numprocs == 1) { vW = lw; vB = lb; return; }
    // place the local shard into a zero matrix and sum over ranks (checkpoint path: simplicity over speed)
    const int r = getGpu()._id, P = getGpu()._numprocs;
    const uint64_t o0 = out * r / P, o1 = out * (r + 1) / P, i0 = in * r / P, i1 = in * (r + 1) / P;
    vW.assign(in * out, 0.0f);
    vB.assign(out, 0.0f);
    if (outgoingLarger) {
        for (uint64_t i = 0; i < in; i++)
            for (uint64_t o = o0; o < o1; o++) vW[i * out + o] = lw[i * (o1 - o0) + (o - o0)];
    } else {
        for (uint64_t i = i0; i < i1; i++)
            for (uint64_t o = 0; o < out; o++) vW[i * out + o] = lw[(i - i0) * out + o];
    }
    for (uint64_t o = o0; o < o1; o++) vB[o] = lb[o - o0];
    GpuBuffer<NNFloat> dW(vW.size()), dB(vB.size());
    dW.Upload(vW.data());
    dB.Upload(vB.data());
    getGpu().Check(dsb200_all_reduce(getGpu()._ctx, dW._pDevData, vW.size()), "dsb200_all_reduce");
    getGpu().Check(dsb200_all_reduce(getGpu()._ctx, dB._pDevData, vB.size()), "dsb200_all_reduce");
    dW.Download(vW.data());
    dB.Download(vB.data());
}

bool NNNetwork::SaveNetCDF(const string& fname)
{
    // un-shard on every rank (collective), write on rank 0 (E/NNNetwork.cpp:1829-1926)
    vector<vector<NNFloat>> vvWeight(_vWeight.size()), vvBias(_vWeight.size());
    for (size_t i = 0; i < _vWeight.size(); i++)
        gather_full(_vWeight[i], _vWeight[i]->_inputLayer._stride, _vWeight[i]->_outputLayer._stride, _vWeight[i]->_bOutgoingLarger, vvWeight[i], vvBias[i]);
    if (getGpu()._id != 0) return true;
    try {
        nc::Writer w(5);
        w.put_att("version", nc::NC_FLOAT, NN_VERSION);
        w.put_att("name", _name);
        w.put_att("kind", nc::NC_UINT, _kind);
        w.put_att("errorFunction", nc::NC_UINT, _errorFunction);
        w.put_att("maxout_k", nc::NC_INT, 2);
        w.put_att("decay", nc::NC_FLOAT, _decay);
        w.put_att("LRN_k", nc::NC_FLOAT, 2.0);
        w.put_att("LRN_n", nc::NC_INT, 5);
        w.put_att("LRN_alpha", nc::NC_FLOAT, 0.0001);
        w.put_att("LRN_beta", nc::NC_FLOAT, 0.75);
        w.put_att("RELUSlope", nc::NC_FLOAT, _RELUSlope);
        w.put_att("ELUAlpha", nc::NC_FLOAT, _ELUAlpha);
        w.put_att("SELULambda", nc::NC_FLOAT, _SELULambda);
        w.put_att("bSparsenessPenalty", nc::NC_UINT, (uint32_t)_bSparsenessPenalty);
        w.put_att("sparsenessPenalty_p", nc::NC_FLOAT, _sparsenessPenalty_p);
        w.put_att("sparsenessPenalty_beta", nc::NC_FLOAT, _sparsenessPenalty_beta);
        w.put_att("bDenoising", nc::NC_UINT, (uint32_t)_bDenoising);
        w.put_att("denoising_p", nc::NC_FLOAT, _denoising_p);
        w.put_att("deltaBoost_one", nc::NC_FLOAT, _deltaBoost_one);
        w.put_att("deltaBoost_zero", nc::NC_FLOAT, _deltaBoost_zero);
        w.put_att("SMCE_oneScale", nc::NC_FLOAT, _SMCE_oneScale);
        w.put_att("SMCE_zeroScale", nc::NC_FLOAT, _SMCE_zeroScale);
        w.put_att("SMCE_oneTarget", nc::NC_FLOAT, _SMCE_oneTarget);
        w.put_att("SMCE_zeroTarget", nc::NC_FLOAT, _SMCE_zeroTarget);
        w.put_att("ShuffleIndices", nc::NC_UINT, (uint32_t)_bShuffleIndices);
        w.put_att("checkpoint_name", _checkpoint_name);
        w.put_att("checkpoint_interval", nc::NC_INT, _checkpoint_interval);
        w.put_att("checkpoint_epochs", nc::NC_INT, _checkpoint_epochs);
        w.put_att("layers", nc::NC_UINT, (double)_vLayer.size());
        for (size_t i = 0; i < _vLayer.size(); i++) {
            const NNLayer* l = _vLayer[i];
            const string ls = "layer" + std::to_string(i) + "_";
            w.put_att(ls + "name", l->_name);
            w.put_att(ls + "kind", nc::NC_UINT, l->_kind);
            w.put_att(ls + "type", nc::NC_UINT, l->_type);
            w.put_att(ls + "poolingfunction", nc::NC_UINT, (uint32_t)PoolingFunction::None);
            w.put_att(ls + "dataSet", l->_dataSet);
            w.put_att(ls + "Nx", nc::NC_UINT, l->_Nx);
            w.put_att(ls + "Ny", nc::NC_UINT, l->_Ny);
            w.put_att(ls + "Nz", nc::NC_UINT, l->_Nz);
            w.put_att(ls + "Nw", nc::NC_UINT, l->_Nw);
            w.put_att(ls + "dimensions", nc::NC_UINT, l->_dimensions);
            for (const char* k : {"kernelX", "kernelY", "kernelZ", "kernelStrideX", "kernelStrideY", "kernelStrideZ"}) w.put_att(ls + k, nc::NC_UINT, 1);
            w.put_att(ls + "kernelDimensions", nc::NC_UINT, 1);
            for (const char* k : {"kernelPaddingX", "kernelPaddingY", "kernelPaddingZ"}) w.put_att(ls + k, nc::NC_UINT, 0);
            w.put_att(ls + "pDropout", nc::NC_FLOAT, l->_pDropout);
            w.put_att(ls + "weightInit", nc::NC_UINT, l->_weightInit);
            w.put_att(ls + "weightInitScale", nc::NC_FLOAT, l->_weightInitScale);
            w.put_att(ls + "biasInit", nc::NC_FLOAT, l->_biasInit);
            w.put_att(ls + "weightNorm", nc::NC_FLOAT, l->_weightNorm);
            w.put_att(ls + "deltaNorm", nc::NC_FLOAT, l->_deltaNorm);
            w.put_att(ls + "activation", nc::NC_UINT, l->_activation);
            w.put_att(ls + "sparsenessPenalty_p", nc::NC_FLOAT, l->_sparsenessPenalty_p);
            w.put_att(ls + "sparsenessPenalty_beta", nc::NC_FLOAT, l->_sparsenessPenalty_beta);
            w.put_att(ls + "RELUSlope", nc::NC_FLOAT, l->_RELUSlope);
            w.put_att(ls + "ELUAlpha", nc::NC_FLOAT, l->_ELUAlpha);
            w.put_att(ls + "SELULambda", nc::NC_FLOAT, l->_SELULambda);
            w.put_att(ls + "attributes", nc::NC_UINT, l->_attributes);
            w.put_att(ls + "sources", nc::NC_UINT, (double)l->_vSource.size());
            for (size_t s = 0; s < l->_vSource.size(); s++) w.put_att(ls + "source" + std::to_string(s), l->_vSource[s]);
            w.put_att(ls + "skips", nc::NC_UINT, 0);
        }
        w.put_att("weights", nc::NC_UINT, (double)_vWeight.size());
        for (size_t i = 0; i < _vWeight.size(); i++) {
            const NNWeight* wt = _vWeight[i];
            const string ws = "weight" + std::to_string(i) + "_";
            w.put_att(ws + "inputLayer", wt->_inputLayer._name);
            w.put_att(ws + "outputLayer", wt->_outputLayer._name);
            w.put_att_u64(ws + "width", nc::NC_UINT64, wt->_outputLayer._stride);
            w.put_att_u64(ws + "height", nc::NC_UINT64, wt->_inputLayer._stride);
            w.put_att_u64(ws + "length", nc::NC_UINT64, 1);
            w.put_att_u64(ws + "depth", nc::NC_UINT64, 1);
            w.put_att_u64(ws + "breadth", nc::NC_UINT64, 1);
            w.put_att(ws + "bShared", nc::NC_UINT, 0);
            w.put_att(ws + "bLocked", nc::NC_UINT, (uint32_t)wt->_bLocked);
            w.put_att(ws + "norm", nc::NC_FLOAT, wt->_norm);
            w.add_dim(ws + "biasDim", vvBias[i].size());
            w.add_var(ws + "bias", nc::NC_FLOAT, ws + "biasDim", vvBias[i].data());
            w.add_dim(ws + "weightDim", vvWeight[i].size());
            w.add_var(ws + "weights", nc::NC_FLOAT, ws + "weightDim", vvWeight[i].data());
        }
        w.write(fname);
    } catch (const nc::Error& e) {
        throw DsbEngineError(string("NNNetwork::SaveNetCDF: ") + e.what());
    }
    return true;
}

// LoadNeuralNetworkNetCDF (E/NNNetwork.cpp:3761-4188)
NNNetwork* LoadNeuralNetworkNetCDF(const string& fname, const uint32_t batch)
{
    NNNetworkDescriptor nd;
    try {
        nc::File f(fname);
        const float version = att_f(f, "version", 0.0f);
        if (version <= 0.0f) throw DsbEngineError("LoadNeuralNetworkNetCDF: No version supplied in NetCDF input file " + fname);
        nd._name = need_att(f, "name", fname).as_string();
        nd._kind = (NNNetwork::Kind)need_att(f, "kind", fname).as_int();
        nd._errorFunction = (ErrorFunction)need_att(f, "errorFunction", fname).as_int();
        nd._decay = att_f(f, "decay", 0.0f);
        nd._RELUSlope = att_f(f, "RELUSlope", nd._RELUSlope);
        nd._ELUAlpha = att_f(f, "ELUAlpha", nd._ELUAlpha);
        nd._SELULambda = att_f(f, "SELULambda", nd._SELULambda);
        nd._bSparsenessPenalty = att_i(f, "bSparsenessPenalty", 0) != 0;
        nd._sparsenessPenalty_p = att_f(f, "sparsenessPenalty_p", 0.0f);
        nd._sparsenessPenalty_beta = att_f(f, "sparsenessPenalty_beta", 0.0f);
        nd._bDenoising = att_i(f, "bDenoising", 0) != 0;
        nd._denoising_p = att_f(f, "denoising_p", 0.0f);
        nd._deltaBoost_one = att_f(f, "deltaBoost_one", 1.0f);
        nd._deltaBoost_zero = att_f(f, "deltaBoost_zero", 1.0f);
        nd._SMCE_oneScale = att_f(f, "SMCE_oneScale", 1.0f);
        nd._SMCE_zeroScale = att_f(f, "SMCE_zeroScale", 1.0f);
        nd._SMCE_oneTarget = att_f(f, "SMCE_oneTarget", 0.9f);
        nd._SMCE_zeroTarget = att_f(f, "SMCE_zeroTarget", 0.1f);
        nd._bShuffleIndices = att_i(f, "ShuffleIndices", 1) != 0;
        nd._checkpoint_name = att_s(f, "checkpoint_name", "checkpoint");
        nd._checkpoint_interval = (int32_t)att_i(f, "checkpoint_interval", 0);
        nd._checkpoint_epochs = (int32_t)att_i(f, "checkpoint_epochs", 0);
        const uint32_t layers = (uint32_t)need_att(f, "layers", fname).as_int();
        for (uint32_t i = 0; i < layers; i++) {
            const string ls = "layer" + std::to_string(i) + "_";
            NNLayerDescriptor ld;
            ld._name = need_att(f, ls + "name", fname).as_string();
            ld._kind = (NNLayer::Kind)need_att(f, ls + "kind", fname).as_int();
            ld._type = (NNLayer::Type)need_att(f, ls + "type", fname).as_int();
            ld._poolingFunction = (PoolingFunction)att_i(f, ls + "poolingfunction", (int64_t)PoolingFunction::None);
            ld._dataSet = att_s(f, ls + "dataSet", "");
            ld._Nx = (uint32_t)need_att(f, ls + "Nx", fname).as_int();
            ld._Ny = (uint32_t)att_i(f, ls + "Ny", 1);
            ld._Nz = (uint32_t)att_i(f, ls + "Nz", 1);
            ld._Nw = (uint32_t)att_i(f, ls + "Nw", 1);
            ld._dimensions = (uint32_t)att_i(f, ls + "dimensions", 1);
            ld._pDropout = att_f(f, ls + "pDropout", 0.0f);
            ld._weightInit = (WeightInitialization)att_i(f, ls + "weightInit", Xavier);
            ld._weightInitScale = att_f(f, ls + "weightInitScale", 1.0f);
            ld._biasInit = att_f(f, ls + "biasInit", 0.0f);
            ld._weightNorm = att_f(f, ls + "weightNorm", 0.0f);
            ld._deltaNorm = att_f(f, ls + "deltaNorm", 0.0f);
            ld._activation = (Activation)need_att(f, ls + "activation", fname).as_int();
            ld._sparsenessPenalty_p = att_f(f, ls + "sparsenessPenalty_p", 0.0f);
            ld._sparsenessPenalty_beta = att_f(f, ls + "sparsenessPenalty_beta", 0.0f);
            ld._RELUSlope = att_f(f, ls + "RELUSlope", NAN);
            ld._ELUAlpha = att_f(f, ls + "ELUAlpha", NAN);
            ld._SELULambda = att_f(f, ls + "SELULambda", NAN);
            ld._attributes = (uint32_t)att_i(f, ls + "attributes", 0);
            const uint32_t sources = (uint32_t)att_i(f, ls + "sources", 0);
            for (uint32_t s = 0; s < sources; s++) ld._vSource.push_back(need_att(f, ls + "source" + std::to_string(s), fname).as_string());
            const uint32_t skips = (uint32_t)att_i(f, ls + "skips", 0);
            for (uint32_t s = 0; s < skips; s++) ld._vSkip.push_back(need_att(f, ls + "skip" + std::to_string(s), fname).as_string());
            nd._vLayerDescriptor.push_back(ld);
        }
        const uint32_t weights = (uint32_t)need_att(f, "weights", fname).as_int();
        for (uint32_t i = 0; i < weights; i++) {
            const string ws = "weight" + std::to_string(i) + "_";
            NNWeightDescriptor wd;
            wd._inputLayer = need_att(f, ws + "inputLayer", fname).as_string();
            wd._outputLayer = need_att(f, ws + "outputLayer", fname).as_string();
            wd._width = (uint64_t)att_i(f, ws + "width", 1);
            wd._height = (uint64_t)att_i(f, ws + "height", 1);
            wd._bShared = att_i(f, ws + "bShared", 0) != 0;
            wd._bLocked = att_i(f, ws + "bLocked", 0) != 0;
            wd._norm = att_f(f, ws + "norm", 0.0f);
            if (wd._bShared) throw DsbEngineError("LoadNeuralNetworkNetCDF: shared weights are outside the dsstne_b200 hot path (" + fname + ")");
            f.read(need_var(f, ws + "bias", fname), wd._vBias);
            f.read(need_var(f, ws + "weights", fname), wd._vWeight);
            nd._vWeightDescriptor.push_back(wd);
        }
    } catch (const nc::Error& e) {
        throw DsbEngineError(string("LoadNeuralNetworkNetCDF: ") + e.what());
    }
    return new NNNetwork(nd, batch);
}
