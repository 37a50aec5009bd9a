// capi.cpp -- network-level C ABI (include/dsstne_b200_engine.h) over the C++ engine classes.
// Exceptions thrown by the engine (its replacement for the reference's print + exit(-1)) become
// error codes here; the message is kept for dsb200_engine_last_error().
#include <cstring>
#include <string>

#include "../../include/dsstne_b200_engine.h"
#include "NNNetwork.h"
#include "NetCDF.h"

using namespace std;

static thread_local string g_lastError;

#define DSB_ENGINE_TRY try {
#define DSB_ENGINE_CATCH                                                             \
    } catch (const std::exception& e) { g_lastError = e.what(); return DSB200_ESTATE; } \
    catch (...) { g_lastError = "unknown exception"; return DSB200_ESTATE; }         \
    return 0;

static NNDataSetBase* DS(dsb200_dataset* d) { return reinterpret_cast<NNDataSetBase*>(d); }
static NNNetwork* NET(dsb200_network* n) { return reinterpret_cast<NNNetwork*>(n); }

template <typename T>
static NNDataSetBase* make_sparse(const char* name, uint32_t examples, uint32_t uniqueExamples, uint32_t w, uint32_t h, uint32_t l,
                                  const uint64_t* s, const uint64_t* e, const uint32_t* idx, const void* data, const float* weight,
                                  const uint32_t* index)
{
    const uint64_t nnz = uniqueExamples ? e[uniqueExamples - 1] : 0;
    NNDataSet<T>* p = new NNDataSet<T>(examples, uniqueExamples, (size_t)nnz, NNDataSetDimensions(w, h, l), index != NULL, weight != NULL, name ? name : "");
    p->LoadSparseData(s, e, data, idx);
    if (index) p->LoadIndexedData(index);
    if (weight) p->LoadDataWeight(weight);
    return p;
}

extern "C" {

const char* dsb200_engine_last_error(void) { return g_lastError.c_str(); }

int dsb200_engine_startup(int rank, int nranks, int device, const void* ncclUniqueId128, uint64_t seed)
{
    DSB_ENGINE_TRY
    getGpu().Startup(rank, nranks, device, ncclUniqueId128);
    getGpu().SetRandomSeed((unsigned long)seed);
    DSB_ENGINE_CATCH
}

int dsb200_engine_shutdown(void)
{
    DSB_ENGINE_TRY
    getGpu().Shutdown();
    DSB_ENGINE_CATCH
}

int dsb200_engine_set_stream(void* cudaStream)
{
    DSB_ENGINE_TRY
    getGpu().SetStream((cudaStream_t)cudaStream);
    DSB_ENGINE_CATCH
}

int dsb200_engine_sync(void)
{
    DSB_ENGINE_TRY
    getGpu().Synchronize();
    DSB_ENGINE_CATCH
}

int dsb200_engine_set_option(const char* name, int value)
{
    DSB_ENGINE_TRY
    if (name && !strcmp(name, "pinned_mirror")) { getGpu()._bPinnedMirror = value != 0; return 0; }      // engine-level option
    if (name && !strcmp(name, "fuse_output_gemm")) { getGpu()._bFuseOutputGemm = value != 0; return 0; }
    if (name && !strcmp(name, "step_trace")) { if (!getGpu()._pNetwork) throw DsbEngineError("step_trace: no network"); getGpu()._pNetwork->SetStepTrace(value != 0); return 0; }
    if (name && !strcmp(name, "p2p_exchange")) { getGpu()._bP2PExchange = value != 0; if (getGpu()._pNetwork) getGpu()._pNetwork->MarkDirty(); return 0; }
    getGpu().Check(dsb200_ctx_set_option(getGpu()._ctx, name, value), "dsb200_ctx_set_option");
    DSB_ENGINE_CATCH
}

int dsb200_engine_profile_report(char* buf, size_t cap)
{
    DSB_ENGINE_TRY
    getGpu().Check(dsb200_profile_report(getGpu()._ctx, buf, cap), "dsb200_profile_report");
    DSB_ENGINE_CATCH
}

int dsb200_engine_step_trace(double* out, int cap)
{
    try { return getGpu()._pNetwork ? getGpu()._pNetwork->StepTraceReport(out, cap) : 0; } catch (...) { return -1; }
}

int dsb200_engine_rank(void) { return getGpu()._id; }
int dsb200_engine_nranks(void) { return getGpu()._numprocs; }

int dsb200_dataset_create_sparse(dsb200_dataset** out, const char* name, int dataType, uint32_t examples, uint32_t uniqueExamples,
                                 uint32_t width, uint32_t height, uint32_t length, const uint64_t* s, const uint64_t* e,
                                 const uint32_t* idx, const void* data, const float* weight, const uint32_t* index, int sparseIgnoreZero)
{
    DSB_ENGINE_TRY
    if (!out || !s || !e || !idx) throw DsbEngineError("dsb200_dataset_create_sparse: null argument");
    NNDataSetBase* p = NULL;
    switch (dataType) {
    case NNDataSetEnums::UInt:   p = make_sparse<uint32_t>(name, examples, uniqueExamples, width, height, length, s, e, idx, data, weight, index); break;
    case NNDataSetEnums::Int:    p = make_sparse<int32_t>(name, examples, uniqueExamples, width, height, length, s, e, idx, data, weight, index); break;
    case NNDataSetEnums::Float:  p = make_sparse<float>(name, examples, uniqueExamples, width, height, length, s, e, idx, data, weight, index); break;
    case NNDataSetEnums::Double: p = make_sparse<double>(name, examples, uniqueExamples, width, height, length, s, e, idx, data, weight, index); break;
    case NNDataSetEnums::UChar:  p = make_sparse<unsigned char>(name, examples, uniqueExamples, width, height, length, s, e, idx, data, weight, index); break;
    case NNDataSetEnums::Char:   p = make_sparse<char>(name, examples, uniqueExamples, width, height, length, s, e, idx, data, weight, index); break;
    default: throw DsbEngineError("dsb200_dataset_create_sparse: unsupported data type");
    }
    if (sparseIgnoreZero) p->_attributes |= NNDataSetEnums::SparseIgnoreZero;
    *out = reinterpret_cast<dsb200_dataset*>(p);
    DSB_ENGINE_CATCH
}

int dsb200_dataset_load_sparse(dsb200_dataset* d, const uint64_t* s, const uint64_t* e, const uint32_t* idx, const void* data)
{
    DSB_ENGINE_TRY
    DS(d)->LoadSparseData(s, e, data, idx);          // NNDataSet<T>::LoadSparseData, E/NNTypes.cpp:610-650
    DSB_ENGINE_CATCH
}

int dsb200_dataset_destroy(dsb200_dataset* d)
{
    DSB_ENGINE_TRY
    delete DS(d);
    DSB_ENGINE_CATCH
}

int dsb200_datasets_load_netcdf(const char* fname, dsb200_dataset** out, int maxOut, int* nOut)
{
    DSB_ENGINE_TRY
    vector<NNDataSetBase*> v = LoadNetCDF(fname);
    if ((int)v.size() > maxOut) { for (auto p : v) delete p; throw DsbEngineError("dsb200_datasets_load_netcdf: output array too small"); }
    for (size_t i = 0; i < v.size(); i++) out[i] = reinterpret_cast<dsb200_dataset*>(v[i]);
    if (nOut) *nOut = (int)v.size();
    DSB_ENGINE_CATCH
}

int dsb200_datasets_save_netcdf(const char* fname, dsb200_dataset** sets, int n)
{
    DSB_ENGINE_TRY
    vector<NNDataSetBase*> v;
    for (int i = 0; i < n; i++) v.push_back(DS(sets[i]));
    if (!SaveNetCDF(fname, v)) throw DsbEngineError(string("SaveNetCDF failed for ") + fname);
    DSB_ENGINE_CATCH
}

int dsb200_netcdf_describe(const char* fname, char* buf, size_t cap)
{
    DSB_ENGINE_TRY
    if (!fname || !buf || !cap) throw DsbEngineError("dsb200_netcdf_describe: null argument");
    const string text = nc::File(fname).describe();
    strncpy(buf, text.c_str(), cap - 1);
    buf[cap - 1] = 0;
    DSB_ENGINE_CATCH
}

int dsb200_netcdf_read_var(const char* fname, const char* var, double* out, uint64_t cap, uint64_t* n)
{
    DSB_ENGINE_TRY
    if (!fname || !var) throw DsbEngineError("dsb200_netcdf_read_var: null argument");
    nc::File f(fname);
    const nc::Var* v = f.var(var);
    if (!v) throw DsbEngineError(string("dsb200_netcdf_read_var: no variable ") + var + " in " + fname);
    if (n) *n = v->nelems;
    if (out) {
        if (cap < v->nelems) throw DsbEngineError("dsb200_netcdf_read_var: output buffer too small");
        vector<double> tmp;
        f.read(*v, tmp);
        memcpy(out, tmp.data(), tmp.size() * sizeof(double));
    }
    DSB_ENGINE_CATCH
}

int dsb200_netcdf_write_sparse(const char* fname, int version, const char* name, uint32_t attributes, int dataType, uint32_t width, uint32_t examples,
                               uint32_t uniqueExamples, const uint64_t* s, const uint64_t* e, const uint32_t* idx, const void* data, const float* weight,
                               const uint32_t* index)
{
    DSB_ENGINE_TRY
    if (!fname || !name || !s || !e || !idx || !uniqueExamples) throw DsbEngineError("dsb200_netcdf_write_sparse: null argument");
    const uint64_t nnz = e[uniqueExamples - 1];
    attributes |= NNDataSetEnums::Sparse;
    if (!data) attributes |= NNDataSetEnums::Boolean;
    if (weight) attributes |= NNDataSetEnums::Weighted;
    if (index) attributes |= NNDataSetEnums::Indexed;
    nc::Writer w(version);
    const bool classic = version != 5;
    const nc::Type U = classic ? nc::NC_INT : nc::NC_UINT;
    w.put_att("datasets", nc::NC_UINT, 1);
    w.put_att("name0", string(name));
    w.put_att("attributes0", nc::NC_UINT, attributes);
    w.put_att("kind0", nc::NC_UINT, NNDataSetEnums::Numeric);
    w.put_att("dataType0", nc::NC_UINT, dataType);
    w.put_att("dimensions0", nc::NC_UINT, 1);
    w.put_att("width0", nc::NC_UINT, width);
    const bool uniq = index != NULL || uniqueExamples != examples;
    if (uniq) w.add_dim("uniqueExamplesDim0", uniqueExamples);
    w.add_dim("examplesDim0", examples);
    w.add_dim("sparseDataDim0", nnz);
    const string rowDim = uniq ? "uniqueExamplesDim0" : "examplesDim0";
    if (nnz > 0x7fffffffull && classic) throw DsbEngineError("dsb200_netcdf_write_sparse: more than 2^31 data points need CDF-5");
    vector<uint32_t> s32(s, s + uniqueExamples), e32(e, e + uniqueExamples);
    if (nnz <= 0xffffffffull) {
        w.add_var("sparseStart0", U, rowDim, s32.data());
        w.add_var("sparseEnd0", U, rowDim, e32.data());
    } else {
        w.add_var("sparseStart0", nc::NC_UINT64, rowDim, s);
        w.add_var("sparseEnd0", nc::NC_UINT64, rowDim, e);
    }
    w.add_var("sparseIndex0", U, "sparseDataDim0", idx);
    if (data) {
        nc::Type t;
        switch (dataType) {
        case NNDataSetEnums::UInt: t = U; break;
        case NNDataSetEnums::Int: t = nc::NC_INT; break;
        case NNDataSetEnums::Float: t = nc::NC_FLOAT; break;
        case NNDataSetEnums::Double: t = nc::NC_DOUBLE; break;
        case NNDataSetEnums::UChar: t = classic ? nc::NC_BYTE : nc::NC_UBYTE; break;
        case NNDataSetEnums::Char: t = nc::NC_BYTE; break;
        default: throw DsbEngineError("dsb200_netcdf_write_sparse: unsupported data type");
        }
        w.add_var("sparseData0", t, "sparseDataDim0", data);
    }
    if (weight) w.add_var("dataWeight0", nc::NC_FLOAT, rowDim, weight);
    if (index) w.add_var("index0", U, "examplesDim0", index);
    w.write(fname);
    DSB_ENGINE_CATCH
}

int dsb200_dataset_info(dsb200_dataset* d, char* name, int nameCap, uint32_t* attributes, uint32_t* examples, uint32_t* width, uint64_t* nnz)
{
    DSB_ENGINE_TRY
    NNDataSetBase* p = DS(d);
    if (name && nameCap > 0) { strncpy(name, p->_name.c_str(), nameCap - 1); name[nameCap - 1] = 0; }
    if (attributes) *attributes = p->_attributes;
    if (examples) *examples = p->_examples;
    if (width) *width = p->_width;
    if (nnz) *nnz = p->_vSparseIndex.size();
    DSB_ENGINE_CATCH
}

static vector<NNDataSetBase*> to_vec(dsb200_dataset** sets, int n)
{
    vector<NNDataSetBase*> v;
    for (int i = 0; i < n; i++) v.push_back(DS(sets[i]));
    return v;
}

int dsb200_network_load_json(dsb200_network** out, const char* jsonText, uint32_t batch, dsb200_dataset** sets, int nSets)
{
    DSB_ENGINE_TRY
    *out = reinterpret_cast<dsb200_network*>(LoadNeuralNetworkJSONString(jsonText, batch, to_vec(sets, nSets)));
    DSB_ENGINE_CATCH
}

int dsb200_network_load_json_file(dsb200_network** out, const char* fname, uint32_t batch, dsb200_dataset** sets, int nSets)
{
    DSB_ENGINE_TRY
    *out = reinterpret_cast<dsb200_network*>(LoadNeuralNetworkJSON(fname, batch, to_vec(sets, nSets)));
    DSB_ENGINE_CATCH
}

int dsb200_network_load_netcdf(dsb200_network** out, const char* fname, uint32_t batch)
{
    DSB_ENGINE_TRY
    *out = reinterpret_cast<dsb200_network*>(LoadNeuralNetworkNetCDF(fname, batch));
    DSB_ENGINE_CATCH
}

int dsb200_network_save_netcdf(dsb200_network* n, const char* fname)
{
    DSB_ENGINE_TRY
    if (!NET(n)->SaveNetCDF(fname)) throw DsbEngineError(string("NNNetwork::SaveNetCDF failed for ") + fname);
    DSB_ENGINE_CATCH
}

int dsb200_network_destroy(dsb200_network* n)
{
    DSB_ENGINE_TRY
    delete NET(n);
    DSB_ENGINE_CATCH
}

int dsb200_network_load_datasets(dsb200_network* n, dsb200_dataset** sets, int nSets)
{
    DSB_ENGINE_TRY
    vector<NNDataSetBase*> v = to_vec(sets, nSets);
    NET(n)->LoadDataSets(v);
    DSB_ENGINE_CATCH
}

int dsb200_network_set_training_mode(dsb200_network* n, int mode) { DSB_ENGINE_TRY NET(n)->SetTrainingMode((TrainingMode)mode); DSB_ENGINE_CATCH }
int dsb200_network_set_batch(dsb200_network* n, uint32_t batch) { DSB_ENGINE_TRY NET(n)->SetBatch(batch); DSB_ENGINE_CATCH }
int dsb200_network_set_position(dsb200_network* n, uint32_t position) { DSB_ENGINE_TRY NET(n)->SetPosition(position); DSB_ENGINE_CATCH }
int dsb200_network_set_shuffle_indices(dsb200_network* n, int flag) { DSB_ENGINE_TRY NET(n)->SetShuffleIndices(flag != 0); DSB_ENGINE_CATCH }
int dsb200_network_get_shuffle_indices(dsb200_network* n, uint32_t* out, uint32_t cap, uint32_t* pCount)
{
    DSB_ENGINE_TRY
    const std::vector<uint32_t>& v = NET(n)->ShuffleIndexVector();
    if (pCount) *pCount = (uint32_t)v.size();
    if (out) for (size_t i = 0; i < v.size() && i < cap; i++) out[i] = v[i];
    DSB_ENGINE_CATCH
}
int dsb200_network_set_decay(dsb200_network* n, float decay) { DSB_ENGINE_TRY NET(n)->SetDecay(decay); DSB_ENGINE_CATCH }
int dsb200_network_set_fusion(dsb200_network* n, int flag) { DSB_ENGINE_TRY NET(n)->SetFusion(flag != 0); DSB_ENGINE_CATCH }
int dsb200_network_set_gemm_mode(dsb200_network* n, int gemmMode)
{
    DSB_ENGINE_TRY
    (void)n;
    getGpu().Check(dsb200_ctx_set_option(getGpu()._ctx, "gemm_mode", gemmMode), "dsb200_ctx_set_option(gemm_mode)");
    DSB_ENGINE_CATCH
}
int dsb200_network_examples(dsb200_network* n, uint32_t* out) { DSB_ENGINE_TRY *out = NET(n)->GetExamples(); DSB_ENGINE_CATCH }

int dsb200_network_train(dsb200_network* n, uint32_t epochs, float alpha, float lambda, float lambda1, float mu, float mu1, float* pError)
{
    DSB_ENGINE_TRY
    const float e = NET(n)->Train(epochs, alpha, lambda, lambda1, mu, mu1);
    if (pError) *pError = e;
    DSB_ENGINE_CATCH
}

int dsb200_network_train_step(dsb200_network* n, uint32_t position, float alpha, float lambda, float lambda1, float mu, float mu1, float* pError)
{
    DSB_ENGINE_TRY
    const float e = NET(n)->TrainStep(position, alpha, lambda, lambda1, mu, mu1);
    if (pError) *pError = e;
    DSB_ENGINE_CATCH
}

int dsb200_network_validate(dsb200_network* n, uint32_t samplesPerMatrix, int* pOk)
{
    DSB_ENGINE_TRY
    if (samplesPerMatrix) NET(n)->SetValidateSamples(samplesPerMatrix);
    const bool ok = NET(n)->Validate();
    if (pOk) *pOk = ok ? 1 : 0;
    DSB_ENGINE_CATCH
}

int dsb200_describe_network_json(const char* jsonText, const char* const* dataSetNames, const uint32_t* dataSetWidths, int nSets, char* buf, size_t cap)
{
    DSB_ENGINE_TRY
    if (!jsonText || !buf || !cap) throw DsbEngineError("describe_network_json: null argument");
    vector<NNDataSetShape> v;
    for (int i = 0; i < nSets; i++) v.push_back(NNDataSetShape{dataSetNames[i], dataSetWidths[i], 1, 1, 1});
    const string d = DescribeNeuralNetworkJSON(jsonText, v);
    if (d.size() + 1 > cap) throw DsbEngineError("describe_network_json: buffer too small");
    memcpy(buf, d.c_str(), d.size() + 1);
    DSB_ENGINE_CATCH
}

int dsb200_network_predict_batch(dsb200_network* n) { DSB_ENGINE_TRY NET(n)->PredictBatch(); DSB_ENGINE_CATCH }

int dsb200_network_topk(dsb200_network* n, const char* layer, uint32_t k, dsb200_dataset* filter, float* outKey, uint32_t* outValue)
{
    DSB_ENGINE_TRY
    NNNetwork* net = NET(n);
    uint32_t batch = net->GetBatch();
    if (net->GetPosition() + batch > net->GetExamples()) batch = net->GetExamples() - net->GetPosition();
    GpuBuffer<NNFloat> key((size_t)batch * k);
    GpuBuffer<uint32_t> val((size_t)batch * k);
    net->CalculateTopKFiltered(layer, k, DS(filter), &key, &val);
    key.Download(outKey);
    val.Download(outValue);
    DSB_ENGINE_CATCH
}

int dsb200_network_topk_global(dsb200_network* n, const char* layer, uint32_t k, dsb200_dataset* filter, float* outKey, uint32_t* outValue)
{
    DSB_ENGINE_TRY
    NNNetwork* net = NET(n);
    uint32_t batch = net->GetBatch();
    if (net->GetPosition() + batch > net->GetExamples()) batch = net->GetExamples() - net->GetPosition();
    GpuBuffer<NNFloat> key((size_t)batch * k);
    GpuBuffer<uint32_t> val((size_t)batch * k);
    net->CalculateTopKGlobal(layer, k, DS(filter), &key, &val);
    key.Download(outKey);
    val.Download(outValue);
    DSB_ENGINE_CATCH
}

int dsb200_network_set_weights(dsb200_network* n, const char* inputLayer, const char* outputLayer, const float* w, uint64_t nW, const float* b, uint64_t nB)
{
    DSB_ENGINE_TRY
    NNWeight* p = NET(n)->GetWeight(inputLayer, outputLayer);
    if (!p) throw DsbEngineError(string("no weights between ") + inputLayer + " and " + outputLayer);
    if (w && !p->SetWeights(vector<NNFloat>(w, w + nW))) throw DsbEngineError("NNWeight::SetWeights: Input vector smaller than weight vector.");
    if (b && !p->SetBiases(vector<NNFloat>(b, b + nB))) throw DsbEngineError("NNWeight::SetBiases: Input vector smaller than bias vector.");
    DSB_ENGINE_CATCH
}

int dsb200_network_get_weights(dsb200_network* n, const char* inputLayer, const char* outputLayer, float* w, uint64_t capW, float* b, uint64_t capB,
                               uint64_t* nW, uint64_t* nB)
{
    DSB_ENGINE_TRY
    NNWeight* p = NET(n)->GetWeight(inputLayer, outputLayer);
    if (!p) throw DsbEngineError(string("no weights between ") + inputLayer + " and " + outputLayer);
    vector<NNFloat> vw, vb;
    p->GetWeights(vw); p->GetBiases(vb);
    if (nW) *nW = vw.size();
    if (nB) *nB = vb.size();
    if (w) { if (capW < vw.size()) throw DsbEngineError("weight buffer too small"); memcpy(w, vw.data(), vw.size() * sizeof(float)); }
    if (b) { if (capB < vb.size()) throw DsbEngineError("bias buffer too small"); memcpy(b, vb.data(), vb.size() * sizeof(float)); }
    DSB_ENGINE_CATCH
}

int dsb200_network_get_gradients(dsb200_network* n, const char* inputLayer, const char* outputLayer, float* g, uint64_t capG, uint64_t* nG)
{
    DSB_ENGINE_TRY
    NNWeight* p = NET(n)->GetWeight(inputLayer, outputLayer);
    if (!p) throw DsbEngineError(string("no weights between ") + inputLayer + " and " + outputLayer);
    vector<NNFloat> vg;
    p->GetGradients(vg);
    if (nG) *nG = vg.size();
    if (g) { if (capG < vg.size()) throw DsbEngineError("gradient buffer too small"); memcpy(g, vg.data(), vg.size() * sizeof(float)); }
    DSB_ENGINE_CATCH
}

static int get_layer_buffer(dsb200_network* n, const char* layer, float* out, uint64_t cap, uint64_t* nOut, bool deltas)
{
    DSB_ENGINE_TRY
    NNLayer* l = NET(n)->GetLayer(layer);
    if (!l) throw DsbEngineError(string("unknown layer ") + layer);
    vector<NNFloat> v;
    if (!(deltas ? l->GetDeltas(v) : l->GetUnits(v))) throw DsbEngineError(string("layer ") + layer + " has no such buffer");
    if (nOut) *nOut = v.size();
    if (out) { if (cap < v.size()) throw DsbEngineError("output buffer too small"); memcpy(out, v.data(), v.size() * sizeof(float)); }
    DSB_ENGINE_CATCH
}

int dsb200_network_get_units(dsb200_network* n, const char* layer, float* out, uint64_t cap, uint64_t* nOut) { return get_layer_buffer(n, layer, out, cap, nOut, false); }
int dsb200_network_get_deltas(dsb200_network* n, const char* layer, float* out, uint64_t cap, uint64_t* nOut) { return get_layer_buffer(n, layer, out, cap, nOut, true); }

int dsb200_network_layer_info(dsb200_network* n, const char* layer, uint32_t* stride, uint32_t* localStride, uint32_t* minX, uint32_t* maxX)
{
    DSB_ENGINE_TRY
    NNLayer* l = NET(n)->GetLayer(layer);
    if (!l) throw DsbEngineError(string("unknown layer ") + layer);
    uint32_t Nx, Ny, Nz, Nw, lx, ly, lz, lw;
    tie(Nx, Ny, Nz, Nw) = l->GetDimensions();
    tie(lx, ly, lz, lw) = l->GetLocalDimensions();
    if (stride) *stride = Nx * Ny * Nz * Nw;
    if (localStride) *localStride = l->GetLocalStride();
    uint32_t a, b;
    dsb200_shard_range(Nx, (uint32_t)getGpu()._id, (uint32_t)getGpu()._numprocs, &a, &b);
    if (minX) *minX = a;
    if (maxX) *maxX = b;
    DSB_ENGINE_CATCH
}

}  // extern "C"
