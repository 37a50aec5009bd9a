/*
 * dsstne_b200_engine.h -- network-level C ABI of libdsstne_b200.so.
 *
 * This is the surface the reference's language bindings sit on: what
 * java/src/main/native/com_amazon_dsstne_Dsstne.cpp (load / load_datasets / predict / shutdown via
 * src/amazon/dsstne/runtime/DsstneContext.cpp) and python/dsstnemodule.cc (Startup, LoadNetCDF,
 * LoadNeuralNetworkJSON, Train, PredictBatch, CalculateTopK, SetTrainingMode ...) call on the C++
 * classes NNNetwork / NNDataSet.  Behind it is the C++ host mirror in amazon-dsstne_b200/engine/
 * (same class and method names as E/NNNetwork.h, E/NNTypes.h); this header only flattens those
 * calls to plain C so ctypes / cgo / JNI can bind them without a C++ toolchain.
 *
 * HOST pointers in, HOST pointers out (the engine owns the device copies, as NNDataSet does).
 * Every function returns 0 on success; dsb200_engine_last_error() explains a failure.
 */
#ifndef DSSTNE_B200_ENGINE_H
#define DSSTNE_B200_ENGINE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dsb200_dataset dsb200_dataset;     /* NNDataSetBase*  */
typedef struct dsb200_network dsb200_network;     /* NNNetwork*      */

const char* dsb200_engine_last_error(void);

/* getGpu().Startup / SetRandomSeed / Shutdown (U/Train.cpp:118-119,166).  ncclUniqueId128 may be NULL
 * when nranks == 1; otherwise it is the id from dsb200_comm_unique_id broadcast by the launcher. */
int dsb200_engine_startup(int rank, int nranks, int device, const void* ncclUniqueId128, uint64_t seed);
int dsb200_engine_shutdown(void);
int dsb200_engine_set_stream(void* cudaStream);
int dsb200_engine_sync(void);
int dsb200_engine_set_option(const char* name, int value);      /* forwards to dsb200_ctx_set_option */
int dsb200_engine_profile_report(char* buf, size_t cap);        /* forwards to dsb200_profile_report */
/* diagnostics, engine option "step_trace" = 1 (set after the network exists): means over the traced NNNetwork::TrainStep calls, in
 * microseconds -- out[0..5] host time of: side-stream preparation launches, forward launches, loss-pass launches, backward launches,
 * the wait for the loss, update launches; out[6..8] device time between: start of the step and the loss, the loss and the end of
 * the updates, the end of a step and the start of the next (the device waiting for the host).  Returns the number of values.      */
int dsb200_engine_step_trace(double* out, int cap);
int dsb200_engine_rank(void);
int dsb200_engine_nranks(void);

/* new NNDataSet<T>(examples, uniqueExamples, sparseDataSize, dim, isIndexed, isWeighted, name) + LoadSparseData
 * + LoadIndexedData + LoadDataWeight (E/NNTypes.cpp:526-770).  sparseData == NULL => Boolean. */
int dsb200_dataset_create_sparse(dsb200_dataset** out, const char* name, int dataType, uint32_t examples, uint32_t uniqueExamples,
                                 uint32_t width, uint32_t height, uint32_t length,
                                 const uint64_t* sparseStart, const uint64_t* sparseEnd, const uint32_t* sparseIndex,
                                 const void* sparseData, const float* dataWeight, const uint32_t* index, int sparseIgnoreZero);
/* NNDataSet<T>::LoadSparseData on an existing dataset (what the JNI binding does per request,
 * java/src/main/native/com_amazon_dsstne_Dsstne.cpp): host CSR in, copied and uploaded */
int dsb200_dataset_load_sparse(dsb200_dataset* d, const uint64_t* sparseStart, const uint64_t* sparseEnd,
                               const uint32_t* sparseIndex, const void* sparseData);
int dsb200_dataset_destroy(dsb200_dataset* d);
/* LoadNetCDF / SaveNetCDF (E/NNTypes.cpp:2456-2584) */
int dsb200_datasets_load_netcdf(const char* fname, dsb200_dataset** out, int maxOut, int* nOut);
int dsb200_datasets_save_netcdf(const char* fname, dsb200_dataset** sets, int n);
int dsb200_dataset_info(dsb200_dataset* d, char* name, int nameCap, uint32_t* attributes, uint32_t* examples, uint32_t* width, uint64_t* nnz);
/* Host-only NetCDF helpers (no GPU, no engine start-up): the classic-format (CDF-1/2/5) reader / writer the engine uses.
 *   describe     `ncdump -h`-like text of any classic file into buf
 *   read_var     a whole variable converted to double (cap elements available in out; *n = elements in the file)
 *   write_sparse one sparse dataset in the schema generateNetCDF emits (U/NetCDFhelper.cpp:332-416); version 5 = CDF-5
 *                (uint / uint64 variables, what the engine itself writes), 2 = CDF-2 with int variables for tools that
 *                only know the classic types; data / weight / exIndex may be NULL                                       */
int dsb200_netcdf_describe(const char* fname, char* buf, size_t cap);
int dsb200_netcdf_read_var(const char* fname, const char* var, double* out, uint64_t cap, uint64_t* n);
int dsb200_netcdf_write_sparse(const char* fname, int version, const char* name, uint32_t attributes, int dataType, uint32_t width,
                               uint32_t examples, uint32_t uniqueExamples, const uint64_t* sparseStart, const uint64_t* sparseEnd,
                               const uint32_t* sparseIndex, const void* sparseData, const float* dataWeight, const uint32_t* index);

/* LoadNeuralNetworkJSON / LoadNeuralNetworkNetCDF / SaveNetCDF / delete (E/NNNetwork.h:287-291) */
int dsb200_network_load_json(dsb200_network** out, const char* jsonText, uint32_t batch, dsb200_dataset** sets, int nSets);
/* HOST ONLY (no GPU needed): parses a network description exactly as dsb200_network_load_json does and writes what it understood,
 * one line per network / layer / weight, into buf.  dataSetNames / dataSetWidths give the dimensions auto-sized layers take from
 * their data sets.  Unknown keys and features outside the hot path fail as in the loader (message: dsb200_engine_last_error).  */
int dsb200_describe_network_json(const char* jsonText, const char* const* dataSetNames, const uint32_t* dataSetWidths, int nSets,
                                 char* buf, size_t cap);
int dsb200_network_load_json_file(dsb200_network** out, const char* fname, uint32_t batch, dsb200_dataset** sets, int nSets);
int dsb200_network_load_netcdf(dsb200_network** out, const char* fname, uint32_t batch);
int dsb200_network_save_netcdf(dsb200_network* n, const char* fname);
int dsb200_network_destroy(dsb200_network* n);

/* NNNetwork::LoadDataSets / SetTrainingMode / SetBatch / SetPosition / SetShuffleIndices / SetDecay ... */
int dsb200_network_load_datasets(dsb200_network* n, dsb200_dataset** sets, int nSets);
int dsb200_network_set_training_mode(dsb200_network* n, int trainingMode);
int dsb200_network_set_batch(dsb200_network* n, uint32_t batch);
int dsb200_network_set_position(dsb200_network* n, uint32_t position);
int dsb200_network_set_shuffle_indices(dsb200_network* n, int flag);
/* the permutation the last NNNetwork::ShuffleIndices() (E/NNNetwork.cpp:826-907) left: position p of an epoch reads example out[p];
 * *pCount = its length (0 before the first shuffled epoch); copies min(cap, *pCount) entries                                  */
int dsb200_network_get_shuffle_indices(dsb200_network* n, uint32_t* out, uint32_t cap, uint32_t* pCount);
int dsb200_network_set_decay(dsb200_network* n, float decay);
int dsb200_network_set_fusion(dsb200_network* n, int flag);          /* B200 fusions on (default) / off */
int dsb200_network_set_gemm_mode(dsb200_network* n, int gemmMode);   /* DSB200_GEMM_* */
int dsb200_network_examples(dsb200_network* n, uint32_t* out);

/* NNNetwork::Train (E/NNNetwork.cpp:1536): returns average error through *pError */
int dsb200_network_train(dsb200_network* n, uint32_t epochs, float alpha, float lambda, float lambda1, float mu, float mu1, float* pError);
/* one minibatch of Train's loop body at `position` (loss of that minibatch through *pError) */
int dsb200_network_train_step(dsb200_network* n, uint32_t position, float alpha, float lambda, float lambda1, float mu, float mu1, float* pError);
/* NNNetwork::Validate (E/NNNetwork.cpp:2459-2633): finite-difference check of the weight and bias gradients on the first batch, through
 * the training kernels; samplesPerMatrix elements per matrix (0 = default 256).  *pOk = 1 when every probe is within 20 * 1e-3. */
int dsb200_network_validate(dsb200_network* n, uint32_t samplesPerMatrix, int* pOk);
/* NNNetwork::PredictBatch at the current position */
int dsb200_network_predict_batch(dsb200_network* n);
/* NNNetwork::CalculateTopK (+ device-side exclusion filter when filter != NULL); HOST outputs [batch][k] */
int dsb200_network_topk(dsb200_network* n, const char* layer, uint32_t k, dsb200_dataset* filter, float* outKey, uint32_t* outValue);
/* model parallel: top-K of the whole layer with GLOBAL unit ids on every rank (what U/NNRecsGenerator.cpp:150-244 assembles on
 * the host from the per-GPU lists); identical to dsb200_network_topk on one GPU */
int dsb200_network_topk_global(dsb200_network* n, const char* layer, uint32_t k, dsb200_dataset* filter, float* outKey, uint32_t* outValue);

/* NNWeight::SetWeights / SetBiases / GetWeights / GetBiases; NNLayer::GetUnits / GetDeltas (HOST buffers).
 * set_* take the FULL [inputStride][outputStride] matrix; get_* return this rank's shard (see NNWeight.h). */
int dsb200_network_set_weights(dsb200_network* n, const char* inputLayer, const char* outputLayer, const float* w, uint64_t nW, const float* b, uint64_t nB);
int dsb200_network_get_weights(dsb200_network* n, const char* inputLayer, const char* outputLayer, float* w, uint64_t capW, float* b, uint64_t capB,
                               uint64_t* nW, uint64_t* nB);
int dsb200_network_get_gradients(dsb200_network* n, const char* inputLayer, const char* outputLayer, float* g, uint64_t capG, uint64_t* nG);
int dsb200_network_get_units(dsb200_network* n, const char* layer, float* out, uint64_t cap, uint64_t* nOut);
int dsb200_network_get_deltas(dsb200_network* n, const char* layer, float* out, uint64_t cap, uint64_t* nOut);
int dsb200_network_layer_info(dsb200_network* n, const char* layer, uint32_t* stride, uint32_t* localStride, uint32_t* minX, uint32_t* maxX);

#ifdef __cplusplus
}
#endif
#endif /* DSSTNE_B200_ENGINE_H */
