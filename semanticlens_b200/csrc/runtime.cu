// Library-wide plumbing: version, thread-local error string, device query.
#include "slb_common.cuh"

#include <string.h>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace {
thread_local char g_err[512] = "";
int g_sm_count = 0;
std::atomic<long long> g_launches{0};
}  // namespace

void slb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void slb_count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int slb_sm_count() {
    if (g_sm_count > 0) return g_sm_count;
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    g_sm_count = n;
    return n;
}

extern "C" int slb_version(void) { return SLB_VERSION; }

extern "C" const char* slb_last_error(void) { return g_err; }

extern "C" int64_t slb_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }

extern "C" int slb_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    SLB_CUDA_OK(cudaGetDevice(&dev));
    int n = 0, maj = 0, min = 0;
    SLB_CUDA_OK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    SLB_CUDA_OK(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
    SLB_CUDA_OK(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
    if (sm_count) *sm_count = n;
    if (cc_major) *cc_major = maj;
    if (cc_minor) *cc_minor = min;
    return SLB_OK;
}


// ---------------------------------------------------------------------------------------------
// live profiling
// ---------------------------------------------------------------------------------------------
namespace {
struct ProfRec {
    const char* name;
    cudaEvent_t a, b;
    double flops, bytes;
};
std::atomic<bool> g_prof_on{false};
std::mutex g_prof_mu;
std::vector<ProfRec*> g_prof;
}  // namespace

bool slb_profile_enabled() { return g_prof_on.load(std::memory_order_relaxed); }

void slb_profile_open(const char* name, void* stream, double flops, double bytes, void** token) {
    ProfRec* r = new ProfRec{name, nullptr, nullptr, flops, bytes};
    if (cudaEventCreate(&r->a) != cudaSuccess || cudaEventCreate(&r->b) != cudaSuccess ||
        cudaEventRecord(r->a, static_cast<cudaStream_t>(stream)) != cudaSuccess) {
        delete r;
        *token = nullptr;
        return;
    }
    *token = r;
}

void slb_profile_close(void* token, void* stream) {
    ProfRec* r = static_cast<ProfRec*>(token);
    cudaEventRecord(r->b, static_cast<cudaStream_t>(stream));
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_prof.push_back(r);
}

static void prof_clear() {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (ProfRec* r : g_prof) {
        cudaEventDestroy(r->a);
        cudaEventDestroy(r->b);
        delete r;
    }
    g_prof.clear();
}

extern "C" int slb_profile_begin(void) {
    prof_clear();
    g_prof_on.store(true);
    return SLB_OK;
}

extern "C" int slb_profile_end(void) {
    g_prof_on.store(false);
    return SLB_OK;
}

extern "C" int slb_profile_summary(SlbKernelTime* out, int max_entries, int* n_entries) {
    SLB_REQUIRE(out && n_entries && max_entries > 0, SLB_EINVAL, "slb_profile_summary: bad arguments");
    std::map<std::string, SlbKernelTime> agg;
    {
        std::lock_guard<std::mutex> lk(g_prof_mu);
        for (ProfRec* r : g_prof) {
            SLB_CUDA_OK(cudaEventSynchronize(r->b));
            float ms = 0.f;
            SLB_CUDA_OK(cudaEventElapsedTime(&ms, r->a, r->b));
            SlbKernelTime& k = agg[r->name];
            if (k.launches == 0) {
                memset(&k, 0, sizeof(k));
                strncpy(k.name, r->name, sizeof(k.name) - 1);
            }
            k.launches += 1;
            k.ms += ms;
            k.flops += r->flops;
            k.bytes += r->bytes;
        }
    }
    int n = 0;
    for (auto& kv : agg) {
        if (n >= max_entries) break;
        out[n++] = kv.second;
    }
    *n_entries = n;
    return SLB_OK;
}
