// util.cu -- error string, launch counter, device queries for libcaptra_ops.so
#include <atomic>
#include <cstdarg>
#include <cstdio>

#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "det_accum.cuh"

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

// Stream-ordered scratch from a pool OWNED by this library (one per device, created on first use, release threshold
// raised so freed scratch is reused by the next call instead of going back to the driver at every synchronisation).
// The process's default pool -- torch's, or anybody else's -- is not touched.
int lib_pool_alloc(void **ptr, size_t bytes, cudaStream_t stream) {
    static std::mutex mu;
    static cudaMemPool_t pools[64] = {};
    int dev = 0;
    CAPTRA_CUDA(cudaGetDevice(&dev));
    CAPTRA_REQUIRE(dev >= 0 && dev < 64, "device index %d out of range", dev);
    cudaMemPool_t pool;
    {
        std::lock_guard<std::mutex> lock(mu);
        if (!pools[dev]) {
            cudaMemPoolProps props = {};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            CAPTRA_CUDA(cudaMemPoolCreate(&pools[dev], &props));
            unsigned long long thr = ~0ull;
            CAPTRA_CUDA(cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &thr));
        }
        pool = pools[dev];
    }
    CAPTRA_CUDA(cudaMallocFromPoolAsync(ptr, bytes, pool, stream));
    return CAPTRA_OK;
}

// ---- deterministic fixed-point accumulation (det_accum.cuh) ----------------------------------------------------
__global__ void det_maxabs_kernel(const float *__restrict__ g, int64_t total, unsigned *__restrict__ maxbits) {
    float m = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const float v = fabsf(__ldg(g + i));
        if (isfinite(v)) m = fmaxf(m, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(maxbits, __float_as_uint(m));     // non-negative floats order like their bits
}

__global__ void det_finalize_kernel(const long long *__restrict__ acc, int64_t total, DetScale sc, float *__restrict__ out) {
    const int e = sc.exponent();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const long long a = acc[i];
        if (a != 0) out[i] += (float)ldexp((double)a, -e);       // the reference adds into the caller's (zeroed) buffer
    }
}

bool det_enabled() {
    static const bool on = [] { const char *e = getenv("CAPTRA_GRAD_ATOMICS"); return !(e && atoi(e) != 0); }();
    return on;
}

int det_begin(const float *grad_out, int64_t n_in, int64_t n_out, long long **acc, unsigned **maxbits, cudaStream_t stream) {
    void *p = nullptr;
    const size_t bytes = (size_t)n_out * sizeof(long long) + 256;
    int rc = lib_pool_alloc(&p, bytes, stream);
    if (rc) return rc;
    CAPTRA_CUDA(cudaMemsetAsync(p, 0, bytes, stream));
    *acc = reinterpret_cast<long long *>(p);
    *maxbits = reinterpret_cast<unsigned *>(reinterpret_cast<char *>(p) + (size_t)n_out * sizeof(long long));
    const int64_t blocks = ceil_div<int64_t>(n_in, 256 * 8);
    det_maxabs_kernel<<<(unsigned)(blocks < 148 * 8 ? (blocks > 0 ? blocks : 1) : 148 * 8), 256, 0, stream>>>(grad_out, n_in, *maxbits);
    CAPTRA_CHECK_LAUNCH("grad(max)");
    return CAPTRA_OK;
}

int det_finish(long long *acc, unsigned *maxbits, int headroom, int64_t n_out, float *grad_points, cudaStream_t stream) {
    DetScale sc{maxbits, headroom};
    const int64_t blocks = ceil_div<int64_t>(n_out, 256 * 4);
    det_finalize_kernel<<<(unsigned)(blocks < 148 * 8 ? (blocks > 0 ? blocks : 1) : 148 * 8), 256, 0, stream>>>(acc, n_out, sc, grad_points);
    cudaError_t e = cudaGetLastError();
    cudaFreeAsync(acc, stream);                                   // stream-ordered: after the finalize kernel
    if (e != cudaSuccess) {
        set_error("grad(finalize): CUDA launch failed: %s", cudaGetErrorString(e));
        return CAPTRA_ERR_CUDA;
    }
    count_launch();
    return CAPTRA_OK;
}

}  // namespace captra

extern "C" const char *captra_last_error(void) { return captra::g_err; }
extern "C" int captra_abi_version(void) { return 1; }
extern "C" int64_t captra_launch_count(void) { return (int64_t)captra::g_launches.load(); }
