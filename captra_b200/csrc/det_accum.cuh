// det_accum.cuh -- deterministic scatter-add for the training-only adjoints (group_points_grad, gather_points_grad,
// three_interpolate_grad; reference: group_points_gpu.cu:8-25, sampling_gpu.cu:46-63, interpolate_gpu.cu:192-214).
//
// The reference accumulates with fp32 atomicAdd, so the sum order -- and the low bits of every gradient -- change from
// run to run.  Here every contribution is converted to 64-bit FIXED POINT with one scale for the whole call
// (2^e, e chosen from max|grad_out| so that 2^22 contributions of that size cannot overflow) and accumulated with
// integer atomics: integer addition is associative, so the result is bit-identical from run to run whatever the
// interleaving, and equals the exact sum rounded once to fp32 (up to 2^-41 max|grad_out| per contribution).
//   pass 1  max|grad_out| (atomicMax on the bit pattern: order-independent)
//   pass 2  the op's scatter with det_add()
//   pass 3  grad_points += float(acc * 2^-e)
// Scratch (8 bytes per output element + one word) comes from the library's own stream-ordered pool.
#pragma once
#include "common.cuh"

namespace captra {

int lib_pool_alloc(void **ptr, size_t bytes, cudaStream_t stream);   // util.cu: library-owned cudaMemPool (per device)

struct DetScale {
    const unsigned *maxbits;      // device: bit pattern of max|g| (a non-negative float)
    int headroom;                 // extra bits for contributions larger than max|g| (interpolation weights above 1)
    __device__ __forceinline__ int exponent() const {
        const float m = __uint_as_float(*maxbits);
        if (!(m > 0.f) || !isfinite(m)) return 0;
        return 40 - headroom - ilogbf(m);                        // |contribution| * 2^e < 2^41
    }
};

__device__ __forceinline__ void det_add(long long *acc, float v, int e) {
    const long long q = __double2ll_rn(ldexp((double)v, e));
    atomicAdd(reinterpret_cast<unsigned long long *>(acc), (unsigned long long)q);
}

__global__ void det_maxabs_kernel(const float *__restrict__ g, int64_t total, unsigned *__restrict__ maxbits);
__global__ void det_finalize_kernel(const long long *__restrict__ acc, int64_t total, DetScale sc, float *__restrict__ out);

// allocates {acc[total] zeroed, maxbits} and runs pass 1; the caller runs its scatter, then det_finish()
int det_begin(const float *grad_out, int64_t n_in, int64_t n_out, long long **acc, unsigned **maxbits, cudaStream_t stream);
int det_finish(long long *acc, unsigned *maxbits, int headroom, int64_t n_out, float *grad_points, cudaStream_t stream);
bool det_enabled();               // CAPTRA_GRAD_ATOMICS=1 restores the reference's fp32 atomics

}  // namespace captra
