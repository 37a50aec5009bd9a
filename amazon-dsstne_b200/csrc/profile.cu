// profile.cu -- per-family device timing for the roofline numbers bench.py reports.
// When option "profile" is on, every C-ABI kernel entry records a CUDA event pair on the context's
// stream; dsb200_profile_report() synchronises and returns, per family, the number of calls and the
// summed device milliseconds.  Off by default (zero overhead beyond one branch per call).
#include "common.cuh"
#include "launch.h"

#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace dsb {

struct ProfileRecord { const char* name; unsigned long long tag; cudaEvent_t start, stop; };
struct ProfileState { std::vector<ProfileRecord> recs; std::vector<cudaEvent_t> pool; };
static std::map<dsb200_ctx*, ProfileState> g_prof;

static cudaEvent_t get_event(ProfileState& st)
{
    if (!st.pool.empty()) { cudaEvent_t e = st.pool.back(); st.pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}

ProfileScope::ProfileScope(dsb200_ctx* c, const char* name, unsigned long long tag) : ctx(c), slot(-1)
{
    if (!c || !c->profile) return;
    ProfileState& st = g_prof[c];
    ProfileRecord r; r.name = name; r.tag = tag; r.start = get_event(st); r.stop = get_event(st);
    cudaEventRecord(r.start, c->stream);
    slot = (int)st.recs.size();
    st.recs.push_back(r);
}

ProfileScope::~ProfileScope()
{
    if (slot < 0) return;
    ProfileState& st = g_prof[ctx];
    cudaEventRecord(st.recs[slot].stop, ctx->stream);
}

}  // namespace dsb

extern "C" {

// writes lines "name calls total_ms\n" into buf (NUL terminated); clears the records
int dsb200_profile_report(dsb200_ctx* ctx, char* buf, size_t cap)
{
    using namespace dsb;
    if (!ctx || !buf || !cap) return DSB200_EINVAL;
    buf[0] = 0;
    DSB_CUDA_OK(cudaStreamSynchronize(ctx->stream));
    ProfileState& st = g_prof[ctx];
    std::map<std::string, std::pair<int, double>> tot;
    for (auto& r : st.recs) {
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, r.start, r.stop);
        auto& t = tot[r.tag ? std::string(r.name) + "@" + std::to_string(r.tag) : std::string(r.name)];
        t.first++; t.second += ms;
        st.pool.push_back(r.start); st.pool.push_back(r.stop);
    }
    st.recs.clear();
    size_t off = 0;
    for (auto& kv : tot) {
        int n = snprintf(buf + off, cap - off, "%s %d %.6f\n", kv.first.c_str(), kv.second.first, kv.second.second);
        if (n < 0 || (size_t)n >= cap - off) break;
        off += (size_t)n;
    }
    return 0;
}

}  // extern "C"
