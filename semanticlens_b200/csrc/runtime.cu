// Library-wide plumbing: version, thread-local error string, device query.
#include "slb_common.cuh"

#include <string.h>
#include <atomic>

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
