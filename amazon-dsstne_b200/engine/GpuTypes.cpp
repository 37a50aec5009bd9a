// GpuTypes.cpp -- see GpuTypes.h.  Replaces E/GpuTypes.cpp (MPI rank -> GPU selection, P2P
// probing, cuBLAS/cuRAND/cuDNN handles) with: rank from the launcher, one dsb200_ctx, NCCL.
#include "GpuTypes.h"

#include <fcntl.h>

#include <sys/stat.h>
#include <time.h>
#include <unistd.h>

static GpuContext g_gpu;
GpuContext& getGpu() { return g_gpu; }

GpuContext::GpuContext()
    : _ctx(nullptr), _numprocs(1), _id(0), _device(0), _warpSize(32), _pNetwork(nullptr), _seed(0), _bStarted(false),
      _totalGPUMemory(0), _totalCPUMemory(0), _stream(nullptr)
{
    dsb200_params_default(&_data);
}

GpuContext::~GpuContext() {}

void GpuContext::Check(int rc, const char* what)
{
    if (rc) throw DsbEngineError(std::string(what) + ": " + (_ctx ? dsb200_last_error(_ctx) : "no context") + " (" + std::to_string(rc) + ")");
}

static int env_int(const char* name, int dflt)
{
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
}

// torchrun-style environment instead of MPI_Init (E/GpuTypes.cpp:62-140): RANK / WORLD_SIZE /
// LOCAL_RANK.  The NCCL unique id has to be handed in by the launcher for nranks > 1.
// For the command line tools (train / predict started by torchrun or any launcher that sets those variables) the id
// is exchanged through a small file: rank 0 writes it, the others wait for it.  DSB200_RENDEZVOUS_FILE names the file;
// the default is /tmp/dsb200_nccl_<MASTER_PORT>_<run id>.
void GpuContext::Startup(int argc, char** argv)
{
    (void)argc; (void)argv;
    const int rank = env_int("RANK", 0), nranks = env_int("WORLD_SIZE", 1), local = env_int("LOCAL_RANK", 0);
    if (nranks <= 1) { Startup(rank, nranks, local, nullptr); return; }
    std::string path;
    if (const char* f = getenv("DSB200_RENDEZVOUS_FILE")) path = f;
    else {
        const char* port = getenv("MASTER_PORT");
        const char* run = getenv("TORCHELASTIC_RUN_ID");
        // every rank of one launch shares the launcher as parent process: its pid keeps the leftovers of a crashed earlier run
        // (same port, same run id) from being picked up
        path = std::string("/tmp/dsb200_nccl_") + (port ? port : "0") + "_" + (run ? run : "default") + "_" + std::to_string((long)getppid());
    }
    unsigned char id[128];
    if (rank == 0) {
        if (dsb200_comm_unique_id(id)) throw DsbEngineError("GpuContext::Startup: cannot create the NCCL unique id");
        const std::string tmp = path + ".tmp";
        remove(path.c_str());                                                     // a leftover of a run with the same launcher pid cannot be ours
        remove(tmp.c_str());
        const int fd = open(tmp.c_str(), O_WRONLY | O_CREAT | O_EXCL, 0600);
        FILE* f = fd >= 0 ? fdopen(fd, "wb") : nullptr;
        if (!f || fwrite(id, 1, sizeof(id), f) != sizeof(id)) throw DsbEngineError("GpuContext::Startup: cannot write " + tmp);
        fclose(f);
        if (rename(tmp.c_str(), path.c_str()) != 0) throw DsbEngineError("GpuContext::Startup: cannot publish " + path);
    } else {
        const time_t started = time(nullptr);
        bool ok = false;
        for (int tries = 0; tries < 1200 && !ok; tries++) {                      // up to two minutes
            struct stat st;
            if (stat(path.c_str(), &st) == 0 && st.st_size == (off_t)sizeof(id) && st.st_mtime + 30 >= started) {   // ignore leftovers of older runs
                FILE* f = fopen(path.c_str(), "rb");
                if (f) { ok = fread(id, 1, sizeof(id), f) == sizeof(id); fclose(f); }
            }
            if (!ok) usleep(100000);
        }
        if (!ok) throw DsbEngineError("GpuContext::Startup: rank " + std::to_string(rank) + " never saw the NCCL id file " + path);
    }
    Startup(rank, nranks, local, id);
    if (rank == 0) { usleep(2000000); remove(path.c_str()); }                    // everyone has joined the communicator by now
}

void GpuContext::Startup(int rank, int nranks, int device, const void* ncclUniqueId128)
{
    if (_bStarted) Shutdown();
    _id = rank; _numprocs = nranks; _device = device;
    int rc = dsb200_ctx_create(&_ctx, device);
    if (rc) throw DsbEngineError("GpuContext::Startup: no sm_100 GPU " + std::to_string(device) + " (there is no CPU fallback)");
    RTERROR(cudaSetDevice(device), "GpuContext::Startup cudaSetDevice");
    dsb200_ctx_set_stream(_ctx, _stream);
    if (nranks > 1) {
        if (!ncclUniqueId128) throw DsbEngineError("GpuContext::Startup: nranks > 1 needs the NCCL unique id from the launcher");
        Check(dsb200_comm_init(_ctx, ncclUniqueId128, rank, nranks), "dsb200_comm_init");
    }
    _bStarted = true;
    CopyConstants();
}

void GpuContext::Shutdown()
{
    if (_ctx) { dsb200_ctx_destroy(_ctx); _ctx = nullptr; }
    if (_copyStream) { cudaStreamDestroy(_copyStream); _copyStream = nullptr; }
    if (_dataConsumedEvent) { cudaEventDestroy(_dataConsumedEvent); _dataConsumedEvent = nullptr; }
    _bDataConsumedValid = false;
    _bStarted = false;
    _pNetwork = nullptr;
}

cudaStream_t GpuContext::CopyStream()
{
    if (!_copyStream) {
        RTERROR(cudaStreamCreateWithFlags(&_copyStream, cudaStreamNonBlocking), "GpuContext: copy stream");
        RTERROR(cudaEventCreateWithFlags(&_dataConsumedEvent, cudaEventDisableTiming), "GpuContext: event");
    }
    return _copyStream;
}

void GpuContext::SetStream(cudaStream_t stream)
{
    _stream = stream;
    if (_ctx) dsb200_ctx_set_stream(_ctx, stream);
}

void GpuContext::SetRandomSeed(unsigned long seed) { _seed = seed; }

void GpuContext::CopyConstants()
{
    if (_ctx) Check(dsb200_ctx_set_params(_ctx, &_data), "dsb200_ctx_set_params");
}

void GpuContext::GetMemoryUsage(int* gpuMemory, int* cpuMemory)
{
    *gpuMemory = (int)(_totalGPUMemory / 1024ll);
    *cpuMemory = (int)(_totalCPUMemory / 1024ll);
}

void GpuContext::Synchronize()
{
    Check(dsb200_ctx_sync(_ctx), "dsb200_ctx_sync");
}
