// GpuTypes.h -- runtime context of the host-side engine (mirror of the reference's
// E/GpuTypes.h:265-454 API for the fully-connected sparse path).
//
// Same names as the reference (getGpu(), GpuContext, GpuBuffer<T>) so NNDataSet / NNLayer /
// NNWeight / NNNetwork read like the reference's, but the implementation is new:
//  * one process per GPU; rank / world size come from the launcher (torchrun-style RANK /
//    WORLD_SIZE / LOCAL_RANK, or dsb200_engine_startup) instead of MPI_Init, and the exchange
//    steps are NCCL collectives inside the C ABI (no CUDA-IPC ring buffers, no MPI barriers);
//  * the `__constant__ GpuData cData` block of the reference becomes the explicit dsb200_params
//    of the kernel library's context (GpuContext::CopyConstants pushes it);
//  * every kernel goes to one stream (the caller's, e.g. torch's current stream).
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/dsstne_b200.h"

typedef float NNFloat;

class NNNetwork;

// error convention of the reference: print + Shutdown + exit(-1) (LAUNCHERROR / RTERROR,
// E/GpuTypes.h:215-254).  The engine throws instead; the C API maps exceptions to error codes.
struct DsbEngineError : public std::runtime_error {
    explicit DsbEngineError(const std::string& s) : std::runtime_error(s) {}
};

#define RTERROR(status, s)                                                                       \
    do { cudaError_t _st = (status); if (_st != cudaSuccess)                                      \
        throw DsbEngineError(std::string(s) + " " + cudaGetErrorString(_st)); } while (0)

struct GpuContext {
    dsb200_ctx*     _ctx;                 // kernel-library context (C ABI)
    dsb200_params   _data;                // what the reference keeps in GpuData (E/GpuTypes.h:265-311)
    int             _numprocs;            // number of model-parallel processes (one GPU each)
    int             _id;                  // this process' rank
    int             _device;              // CUDA device ordinal
    unsigned int    _warpSize;
    NNNetwork*      _pNetwork;
    unsigned long   _seed;
    bool            _bStarted;
    bool            _bFuseOutputGemm = true;  // engine option "fuse_output_gemm" (default on): the forward GEMM of a sigmoid output layer over Boolean
                                              // sparse targets is deferred into the loss / delta pass and runs as dsb200_gemm_fwd_output_pass (Z never written)
    bool            _bP2PExchange = true;     // engine option "p2p_exchange" (default on): model-parallel exchange steps as one kernel over peer
                                              // memory (dsb200_p2p_*); falls back to NCCL when the peers cannot be mapped
    bool            _bPinnedMirror = true;    // engine option "pinned_mirror" (default on): NNDataSet::LoadSparseData uploads straight from the
                                              // page-locked host mirror instead of copying every batch a second time into staging (0, or a
                                              // failed cudaHostRegister: the two-copy staging path)
    long long       _totalGPUMemory, _totalCPUMemory;

    GpuContext();
    ~GpuContext();
    void Startup(int argc, char** argv);                          // E/GpuTypes.cpp:62
    void Startup(int rank, int nranks, int device, const void* ncclUniqueId128);
    void Shutdown();
    void SetRandomSeed(unsigned long seed);                       // E/GpuTypes.cpp:501
    void SetNeuralNetwork(NNNetwork* pNetwork);                   // E/GpuTypes.cpp:475
    void CopyConstants();                                         // E/GpuTypes.cpp:408
    void SetStream(cudaStream_t stream);
    cudaStream_t GetStream() const { return _stream; }
    // Streamed datasets (NNDataSet::LoadSparseData once per step) upload on their own stream so that the copies of batch k + 1 run
    // beside the backward pass of batch k: the network records _dataConsumedEvent once the last reader of the CSR buffers of a step
    // has been launched (NNNetwork::LaunchError); an upload waits for it on the copy stream, and the next step waits for the upload.
    cudaStream_t CopyStream();
    cudaEvent_t     _dataConsumedEvent = nullptr;
    bool            _bDataConsumedValid = false;
    void GetMemoryUsage(int* gpuMemory, int* cpuMemory);
    void Check(int rc, const char* what);                         // throws on a non-zero C-ABI code
    void Synchronize();

private:
    cudaStream_t    _stream;
    cudaStream_t    _copyStream = nullptr;
};

GpuContext& getGpu();

template <typename T>
struct GpuBuffer {
    size_t  _length;
    bool    _bSysMem;          // keep a pinned host shadow
    bool    _bManaged;         // kept for API compatibility; managed memory is not used on B200
    T*      _pSysData;
    T*      _pDevData;

    GpuBuffer(size_t length, bool bSysMem = false, bool bManaged = false)
        : _length(length), _bSysMem(bSysMem), _bManaged(bManaged), _pSysData(nullptr), _pDevData(nullptr) { Allocate(); }
    virtual ~GpuBuffer() { Deallocate(); }

    void Allocate()
    {
        size_t n = _length ? _length : 1;
        RTERROR(cudaMalloc((void**)&_pDevData, n * sizeof(T)), "GpuBuffer::Allocate failed");
        RTERROR(cudaMemsetAsync(_pDevData, 0, n * sizeof(T), getGpu().GetStream()), "GpuBuffer::Allocate memset failed");
        getGpu()._totalGPUMemory += (long long)(n * sizeof(T));
        if (_bSysMem) {
            RTERROR(cudaHostAlloc((void**)&_pSysData, n * sizeof(T), cudaHostAllocDefault), "GpuBuffer::Allocate pinned failed");
            memset(_pSysData, 0, n * sizeof(T));
            getGpu()._totalCPUMemory += (long long)(n * sizeof(T));
        }
    }
    void Deallocate()
    {
        if (_pDevData) { cudaFree(_pDevData); getGpu()._totalGPUMemory -= (long long)((_length ? _length : 1) * sizeof(T)); }
        if (_pSysData) { cudaFreeHost(_pSysData); getGpu()._totalCPUMemory -= (long long)((_length ? _length : 1) * sizeof(T)); }
        _pDevData = nullptr; _pSysData = nullptr;
    }
    // grows only; contents are lost (same contract as the reference, E/GpuTypes.h:412-424)
    void Resize(size_t length) { if (length > _length) { Deallocate(); _length = length; Allocate(); } }
    void Upload(const T* pBuff = nullptr) const
    {
        const T* src = pBuff ? pBuff : _pSysData;
        if (!src) throw DsbEngineError("GpuBuffer::Upload: no source");
        RTERROR(cudaMemcpyAsync(_pDevData, src, _length * sizeof(T), cudaMemcpyHostToDevice, getGpu().GetStream()), "GpuBuffer::Upload failed");
        RTERROR(cudaStreamSynchronize(getGpu().GetStream()), "GpuBuffer::Upload sync failed");
    }
    void Download(T* pBuff = nullptr)
    {
        T* dst = pBuff ? pBuff : _pSysData;
        if (!dst) throw DsbEngineError("GpuBuffer::Download: no destination");
        RTERROR(cudaMemcpyAsync(dst, _pDevData, _length * sizeof(T), cudaMemcpyDeviceToHost, getGpu().GetStream()), "GpuBuffer::Download failed");
        RTERROR(cudaStreamSynchronize(getGpu().GetStream()), "GpuBuffer::Download sync failed");
    }
    void Copy(T* pBuff)
    {
        RTERROR(cudaMemcpyAsync(_pDevData, pBuff, _length * sizeof(T), cudaMemcpyDeviceToDevice, getGpu().GetStream()), "GpuBuffer::Copy failed");
    }
    size_t GetLength() { return _length; }
    size_t GetSize() { return _length * sizeof(T); }
};
