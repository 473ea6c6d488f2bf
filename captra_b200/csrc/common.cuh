// common.cuh -- shared helpers for libcaptra_ops.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "captra_ops.h"

namespace captra {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// error plumbing ----------------------------------------------------------------------------
void set_error(const char *fmt, ...);
void count_launch(int n = 1);

#define CAPTRA_REQUIRE(cond, ...)                 \
    do {                                          \
        if (!(cond)) {                            \
            captra::set_error(__VA_ARGS__);       \
            return CAPTRA_ERR_INVALID_ARG;        \
        }                                         \
    } while (0)

// Checks the launch (not the execution -- we never synchronise inside the library).
#define CAPTRA_CHECK_LAUNCH(name)                                                      \
    do {                                                                               \
        cudaError_t e__ = cudaGetLastError();                                          \
        if (e__ != cudaSuccess) {                                                      \
            captra::set_error("%s: CUDA launch failed: %s", name, cudaGetErrorString(e__)); \
            return CAPTRA_ERR_CUDA;                                                    \
        }                                                                              \
        captra::count_launch();                                                        \
    } while (0)

#define CAPTRA_CUDA(call)                                                              \
    do {                                                                               \
        cudaError_t e__ = (call);                                                      \
        if (e__ != cudaSuccess) {                                                      \
            captra::set_error("%s failed: %s", #call, cudaGetErrorString(e__));        \
            return CAPTRA_ERR_CUDA;                                                    \
        }                                                                              \
    } while (0)

inline cudaStream_t as_stream(captra_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();

// arithmetic contract -----------------------------------------------------------------------
// Squared distance in the exact operation order nvcc gives the reference kernels
// (FMUL dy*dy; FFMA dx*dx+.; FFMA dz*dz+.  -- verified in the SASS of oracle/_ref, SURVEY App.
// A.1).  Intrinsics are never re-associated or re-contracted by the compiler.
__device__ __forceinline__ float sqdist_ref(float ax, float ay, float az, float bx, float by,
                                            float bz) {
    const float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    float t = __fmul_rn(dy, dy);
    t = __fmaf_rn(dx, dx, t);
    return __fmaf_rn(dz, dz, t);
}

__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// streaming (read-once / write-once) accesses that should not displace reusable L1 lines
__device__ __forceinline__ void st_stream(float *p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream4(float4 *p, float4 v) { __stcs(p, v); }

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) {
    return (a + b - 1) / b;
}

}  // namespace captra
