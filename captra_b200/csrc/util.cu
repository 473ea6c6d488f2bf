// util.cu -- error string, launch counter, device queries for libcaptra_ops.so
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "common.cuh"

namespace captra {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached[dev] = v;
    }
    return cached[dev];
}

}  // namespace captra

extern "C" const char *captra_last_error(void) { return captra::g_err; }
extern "C" int captra_abi_version(void) { return 1; }
extern "C" int64_t captra_launch_count(void) { return (int64_t)captra::g_launches.load(); }
