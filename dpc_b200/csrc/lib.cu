// Library-level plumbing: error string, launch counter, SM count.
#include "common.cuh"
#include <stdarg.h>
#include <atomic>

static thread_local char g_err[1024] = "";
std::atomic<long long> g_launches{0};

void dpc_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void dpc_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int dpc_num_sms() {
    static thread_local int cached_dev = -1;
    static thread_local int cached = 0;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev) {
        cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
        cached_dev = dev;
    }
    return cached > 0 ? cached : 148;
}

extern "C" int dpc_abi_version(void) { return DPC_B200_ABI_VERSION; }
extern "C" const char* dpc_last_error(void) { return g_err; }
extern "C" int64_t dpc_launch_count(void) { return (int64_t)g_launches.load(); }
