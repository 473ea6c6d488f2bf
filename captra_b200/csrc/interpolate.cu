// interpolate.cu -- three_nn, knn, three_interpolate (+grad) and the fused
// three_nn -> weights -> interpolate used by feature propagation.
// Reference: interpolate_gpu.cu:9-214, pointnet2_utils.py:110-192, pointnet_utils.py:284-289.
//
// B200 design: one thread per query point; the `known` set (128..512 points in the backbone)
// is staged once per CTA in shared memory as float4 so the scan is one broadcast LDS.128 +
// the exact-order distance + a guarded cascade (the cascade body only runs when d < best3).
// The reference compares in double against 1e40 sentinels (interpolate_gpu.cu:102-120); a
// float +inf sentinel orders every float d identically and converts to the same float output.
#include "common.cuh"
#include "det_accum.cuh"

#include <math.h>

namespace captra {

constexpr int NN_THREADS = 256;
constexpr int NN_TILE = 2048;  // known points per shared-memory tile (32 KB)

struct Best3 {
    float d1, d2, d3;
    int i1, i2, i3;
};

// Scans known[b] for the query (ux,uy,uz).  All threads of the CTA must call it.
__device__ __forceinline__ Best3 scan_three_nn(int m, const float *__restrict__ known_b,
                                               float ux, float uy, float uz, bool active,
                                               float4 *tile) {
    Best3 r{INFINITY, INFINITY, INFINITY, 0, 0, 0};
    for (int t0 = 0; t0 < m; t0 += NN_TILE) {
        const int tn = min(NN_TILE, m - t0);
        if (t0 > 0) __syncthreads();
        for (int k = threadIdx.x; k < tn; k += NN_THREADS) {
            const float *p = known_b + (size_t)(t0 + k) * 3;
            tile[k] = make_float4(__ldg(p), __ldg(p + 1), __ldg(p + 2), 0.f);
        }
        __syncthreads();
        if (!active) continue;
#pragma unroll 4
        for (int k = 0; k < tn; ++k) {
            const float4 q = tile[k];
            const float d = sqdist_ref(ux, uy, uz, q.x, q.y, q.z);
            if (d < r.d3) {
                const int kk = t0 + k;
                if (d < r.d1) {
                    r.d3 = r.d2; r.i3 = r.i2; r.d2 = r.d1; r.i2 = r.i1; r.d1 = d; r.i1 = kk;
                } else if (d < r.d2) {
                    r.d3 = r.d2; r.i3 = r.i2; r.d2 = d; r.i2 = kk;
                } else {
                    r.d3 = d; r.i3 = kk;
                }
            }
        }
    }
    return r;
}

__global__ void __launch_bounds__(NN_THREADS)
three_nn_kernel(int n, int m, const float *__restrict__ unknown, const float *__restrict__ known,
                float *__restrict__ dist2, int *__restrict__ idx) {
    __shared__ float4 tile[NN_TILE];
    const int b = blockIdx.y;
    const int pt = blockIdx.x * NN_THREADS + threadIdx.x;
    const bool active = pt < n;
    float ux = 0, uy = 0, uz = 0;
    if (active) {
        const float *u = unknown + ((size_t)b * n + pt) * 3;
        ux = __ldg(u); uy = __ldg(u + 1); uz = __ldg(u + 2);
    }
    const Best3 r = scan_three_nn(m, known + (size_t)b * m * 3, ux, uy, uz, active, tile);
    if (!active) return;
    float *od = dist2 + ((size_t)b * n + pt) * 3;
    int *oi = idx + ((size_t)b * n + pt) * 3;
    od[0] = r.d1; od[1] = r.d2; od[2] = r.d3;
    oi[0] = r.i1; oi[1] = r.i2; oi[2] = r.i3;
}

// Inverse-distance weights exactly as the Python layer computes them in fp32:
// dist = sqrt(d2) (pointnet2_utils.py:134); r = 1/(dist+1e-8); w = r / (r0+r1+r2)
// (pointnet_utils.py:285-287).
__device__ __forceinline__ void idw_weights(const Best3 &r, float &dd1, float &dd2, float &dd3,
                                            float &w1, float &w2, float &w3) {
    dd1 = __fsqrt_rn(r.d1); dd2 = __fsqrt_rn(r.d2); dd3 = __fsqrt_rn(r.d3);
    const float r1 = __fdiv_rn(1.0f, __fadd_rn(dd1, 1e-8f));
    const float r2 = __fdiv_rn(1.0f, __fadd_rn(dd2, 1e-8f));
    const float r3 = __fdiv_rn(1.0f, __fadd_rn(dd3, 1e-8f));
    const float norm = __fadd_rn(__fadd_rn(r1, r2), r3);
    w1 = __fdiv_rn(r1, norm); w2 = __fdiv_rn(r2, norm); w3 = __fdiv_rn(r3, norm);
}

// out = fma(w2,p2, fma(w0,p0, w1*p1)) -- the reference's contraction
__device__ __forceinline__ float interp3(float w0, float p0, float w1, float p1, float w2, float p2) {
    return __fmaf_rn(w2, p2, __fmaf_rn(w0, p0, __fmul_rn(w1, p1)));
}

// Stage 1 of the fused FP path: 3-NN + weights, written as idx/weight [B,n,3].
__global__ void __launch_bounds__(NN_THREADS)
three_nn_weights_kernel(int n, int m, const float *__restrict__ unknown,
                        const float *__restrict__ known, float *__restrict__ dist,
                        int *__restrict__ idx, float *__restrict__ weight) {
    __shared__ float4 tile[NN_TILE];
    const int b = blockIdx.y;
    const int pt = blockIdx.x * NN_THREADS + threadIdx.x;
    const bool active = pt < n;
    float ux = 0, uy = 0, uz = 0;
    if (active) {
        const float *u = unknown + ((size_t)b * n + pt) * 3;
        ux = __ldg(u); uy = __ldg(u + 1); uz = __ldg(u + 2);
    }
    const Best3 r = scan_three_nn(m, known + (size_t)b * m * 3, ux, uy, uz, active, tile);
    if (!active) return;
    float d1, d2, d3, w1, w2, w3;
    idw_weights(r, d1, d2, d3, w1, w2, w3);
    const size_t o = ((size_t)b * n + pt) * 3;
    if (dist) { dist[o] = d1; dist[o + 1] = d2; dist[o + 2] = d3; }
    idx[o] = r.i1; idx[o + 1] = r.i2; idx[o + 2] = r.i3;
    weight[o] = w1; weight[o + 1] = w2; weight[o + 2] = w3;
}

// channel-major interpolate (reference layout): points [B,C,m] -> out [B,C,n]
constexpr int TI_CH_BLOCK = 8;
__global__ void __launch_bounds__(NN_THREADS)
three_interpolate_kernel(int c, int m, int n, const float *__restrict__ points,
                         const int *__restrict__ idx, const float *__restrict__ weight,
                         float *__restrict__ out) {
    const int b = blockIdx.z;
    const int pt = blockIdx.x * NN_THREADS + threadIdx.x;
    if (pt >= n) return;
    const size_t o = ((size_t)b * n + pt) * 3;
    const int i0 = __ldg(idx + o), i1 = __ldg(idx + o + 1), i2 = __ldg(idx + o + 2);
    const float w0 = __ldg(weight + o), w1 = __ldg(weight + o + 1), w2 = __ldg(weight + o + 2);
    const int cblk = blockIdx.y * TI_CH_BLOCK, cend = min(c, cblk + TI_CH_BLOCK);
    for (int ci = cblk; ci < cend; ++ci) {
        const float *src = points + ((size_t)b * c + ci) * m;
        st_stream(out + ((size_t)b * c + ci) * n + pt,
                  interp3(w0, __ldg(src + i0), w1, __ldg(src + i1), w2, __ldg(src + i2)));
    }
}

// point-major interpolate (internal layout): points [B,m,C] -> out rows of stride ldo at
// column offset col0.  One warp per query point, lanes across channels: three coalesced row
// reads and one coalesced row write per point.
__global__ void __launch_bounds__(NN_THREADS)
three_interpolate_pm_kernel(int c, int m, int n, const float *__restrict__ points,
                            const int *__restrict__ idx, const float *__restrict__ weight,
                            float *__restrict__ out, int64_t ldo, int col0) {
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pt = blockIdx.x * (NN_THREADS / 32) + warp;
    if (pt >= n) return;
    const size_t o = ((size_t)b * n + pt) * 3;
    const int i0 = __ldg(idx + o), i1 = __ldg(idx + o + 1), i2 = __ldg(idx + o + 2);
    const float w0 = __ldg(weight + o), w1 = __ldg(weight + o + 1), w2 = __ldg(weight + o + 2);
    const float *p0 = points + ((size_t)b * m + i0) * c;
    const float *p1 = points + ((size_t)b * m + i1) * c;
    const float *p2 = points + ((size_t)b * m + i2) * c;
    float *dst = out + ((size_t)b * n + pt) * ldo + col0;
    for (int ci = lane; ci < c; ci += 32)
        dst[ci] = interp3(w0, __ldg(p0 + ci), w1, __ldg(p1 + ci), w2, __ldg(p2 + ci));
}

__global__ void __launch_bounds__(NN_THREADS)
three_interpolate_grad_kernel(int c, int n, int m, const float *__restrict__ grad_out,
                              const int *__restrict__ idx, const float *__restrict__ weight,
                              float *__restrict__ grad_points) {
    const int b = blockIdx.z;
    const int pt = blockIdx.x * NN_THREADS + threadIdx.x;
    if (pt >= n) return;
    const size_t o = ((size_t)b * n + pt) * 3;
    const int i0 = __ldg(idx + o), i1 = __ldg(idx + o + 1), i2 = __ldg(idx + o + 2);
    const float w0 = __ldg(weight + o), w1 = __ldg(weight + o + 1), w2 = __ldg(weight + o + 2);
    const int cblk = blockIdx.y * TI_CH_BLOCK, cend = min(c, cblk + TI_CH_BLOCK);
    for (int ci = cblk; ci < cend; ++ci) {
        const float g = __ldg(grad_out + ((size_t)b * c + ci) * n + pt);
        float *dst = grad_points + ((size_t)b * c + ci) * m;
        atomicAdd(dst + i0, __fmul_rn(g, w0));
        atomicAdd(dst + i1, __fmul_rn(g, w1));
        atomicAdd(dst + i2, __fmul_rn(g, w2));
    }
}

// deterministic variant (det_accum.cuh): the fp32 products g * w are formed as in the reference, accumulated in 64-bit
// fixed point; max|grad_out| bounds the products of convex weights, and 8 bits of head-room cover callers whose weights
// are not (|w| <= 256; the resolution is still 2^-33 max|grad_out|, far below fp32's 2^-24)
__global__ void __launch_bounds__(NN_THREADS)
three_interpolate_grad_det_kernel(int c, int n, int m, const float *__restrict__ grad_out,
                                  const int *__restrict__ idx, const float *__restrict__ weight,
                                  long long *__restrict__ acc, DetScale sc) {
    const int b = blockIdx.z;
    const int pt = blockIdx.x * NN_THREADS + threadIdx.x;
    if (pt >= n) return;
    const int e = sc.exponent();
    const size_t o = ((size_t)b * n + pt) * 3;
    const int i0 = __ldg(idx + o), i1 = __ldg(idx + o + 1), i2 = __ldg(idx + o + 2);
    const float w0 = __ldg(weight + o), w1 = __ldg(weight + o + 1), w2 = __ldg(weight + o + 2);
    const int cblk = blockIdx.y * TI_CH_BLOCK, cend = min(c, cblk + TI_CH_BLOCK);
    for (int ci = cblk; ci < cend; ++ci) {
        const float g = __ldg(grad_out + ((size_t)b * c + ci) * n + pt);
        long long *dst = acc + ((size_t)b * c + ci) * m;
        det_add(dst + i0, __fmul_rn(g, w0), e);
        det_add(dst + i1, __fmul_rn(g, w1), e);
        det_add(dst + i2, __fmul_rn(g, w2), e);
    }
}

// knn (interpolate_gpu.cu:9-57): sorted insertion list of k <= 200 candidates per query.
// The list lives in shared memory (k floats + k ints per thread would spill 2.4 kB/thread in
// the reference); strict '<' keeps the earlier index on ties exactly like the reference.
constexpr int KNN_THREADS = 64;
__global__ void __launch_bounds__(KNN_THREADS)
knn_kernel(int n, int m, int k, const float *__restrict__ unknown, const float *__restrict__ known,
           float *__restrict__ dist2, int *__restrict__ idx) {
    extern __shared__ float sm[];  // best[k][KNN_THREADS], besti[k][KNN_THREADS]
    float *best = sm;
    int *besti = reinterpret_cast<int *>(sm + (size_t)k * KNN_THREADS);
    const int b = blockIdx.y;
    const int pt = blockIdx.x * KNN_THREADS + threadIdx.x;
    if (pt >= n) return;
    const int t = threadIdx.x;
    const float *u = unknown + ((size_t)b * n + pt) * 3;
    const float ux = __ldg(u), uy = __ldg(u + 1), uz = __ldg(u + 2);
    for (int i = 0; i < k; ++i) { best[i * KNN_THREADS + t] = INFINITY; besti[i * KNN_THREADS + t] = 0; }
    const float *kn = known + (size_t)b * m * 3;
    for (int i = 0; i < m; ++i) {
        const float d = sqdist_ref(ux, uy, uz, __ldg(kn + i * 3), __ldg(kn + i * 3 + 1), __ldg(kn + i * 3 + 2));
        if (!(d < best[(k - 1) * KNN_THREADS + t])) continue;
        int j = k - 1;  // shift larger entries down while the slot above is still > d
        while (j > 0 && d < best[(j - 1) * KNN_THREADS + t]) {
            best[j * KNN_THREADS + t] = best[(j - 1) * KNN_THREADS + t];
            besti[j * KNN_THREADS + t] = besti[(j - 1) * KNN_THREADS + t];
            --j;
        }
        best[j * KNN_THREADS + t] = d;
        besti[j * KNN_THREADS + t] = i;
    }
    float *od = dist2 + ((size_t)b * n + pt) * k;
    int *oi = idx + ((size_t)b * n + pt) * k;
    for (int i = 0; i < k; ++i) { od[i] = best[i * KNN_THREADS + t]; oi[i] = besti[i * KNN_THREADS + t]; }
}

}  // namespace captra

using namespace captra;

extern "C" int three_nn_kernel_launcher_fast(int b, int n, int m, const float *unknown,
                                             const float *known, float *dist2, int *idx,
                                             captra_stream_t stream) {
    CAPTRA_REQUIRE(b >= 0 && n >= 0 && m >= 0, "three_nn: negative size");
    if (b == 0 || n == 0) return CAPTRA_OK;
    CAPTRA_REQUIRE(unknown && dist2 && idx && (known || m == 0), "three_nn: null pointer");
    CAPTRA_REQUIRE(b <= 65535, "three_nn: batch exceeds grid limit");
    dim3 grid(ceil_div(n, NN_THREADS), b);
    three_nn_kernel<<<grid, NN_THREADS, 0, as_stream(stream)>>>(n, m, unknown, known, dist2, idx);
    CAPTRA_CHECK_LAUNCH("three_nn");
    return CAPTRA_OK;
}

extern "C" int knn_kernel_launcher_fast(int b, int n, int m, int k, const float *unknown,
                                        const float *known, float *dist2, int *idx,
                                        captra_stream_t stream) {
    CAPTRA_REQUIRE(b >= 0 && n >= 0 && m >= 0, "knn: negative size");
    CAPTRA_REQUIRE(k >= 1 && k <= 200, "knn: k=%d outside 1..200 (interpolate_gpu.cu:30)", k);
    if (b == 0 || n == 0) return CAPTRA_OK;
    CAPTRA_REQUIRE(unknown && dist2 && idx && (known || m == 0), "knn: null pointer");
    CAPTRA_REQUIRE(b <= 65535, "knn: batch exceeds grid limit");
    const size_t smem = (size_t)k * KNN_THREADS * 8;
    if (smem > 48 * 1024)
        CAPTRA_CUDA(cudaFuncSetAttribute(knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(n, KNN_THREADS), b);
    knn_kernel<<<grid, KNN_THREADS, smem, as_stream(stream)>>>(n, m, k, unknown, known, dist2, idx);
    CAPTRA_CHECK_LAUNCH("knn");
    return CAPTRA_OK;
}

extern "C" int three_interpolate_kernel_launcher_fast(int b, int c, int m, int n,
                                                      const float *points, const int *idx,
                                                      const float *weight, float *out,
                                                      captra_stream_t stream) {
    CAPTRA_REQUIRE(b >= 0 && c >= 0 && m >= 0 && n >= 0, "three_interpolate: negative size");
    if (b == 0 || c == 0 || n == 0) return CAPTRA_OK;
    CAPTRA_REQUIRE(points && idx && weight && out, "three_interpolate: null pointer");
    CAPTRA_REQUIRE(b <= 65535 && ceil_div(c, TI_CH_BLOCK) <= 65535, "three_interpolate: grid limit");
    dim3 grid(ceil_div(n, NN_THREADS), ceil_div(c, TI_CH_BLOCK), b);
    three_interpolate_kernel<<<grid, NN_THREADS, 0, as_stream(stream)>>>(c, m, n, points, idx, weight, out);
    CAPTRA_CHECK_LAUNCH("three_interpolate");
    return CAPTRA_OK;
}

extern "C" int three_interpolate_grad_kernel_launcher_fast(int b, int c, int n, int m,
                                                           const float *grad_out, const int *idx,
                                                           const float *weight, float *grad_points,
                                                           captra_stream_t stream) {
    CAPTRA_REQUIRE(b >= 0 && c >= 0 && m >= 0 && n >= 0, "three_interpolate_grad: negative size");
    if (b == 0 || c == 0 || n == 0) return CAPTRA_OK;
    CAPTRA_REQUIRE(grad_out && idx && weight && grad_points, "three_interpolate_grad: null pointer");
    CAPTRA_REQUIRE(b <= 65535 && ceil_div(c, TI_CH_BLOCK) <= 65535, "three_interpolate_grad: grid limit");
    dim3 grid(ceil_div(n, NN_THREADS), ceil_div(c, TI_CH_BLOCK), b);
    if (det_enabled()) {
        cudaStream_t s = as_stream(stream);
        long long *acc = nullptr;
        unsigned *maxbits = nullptr;
        int rc = det_begin(grad_out, (int64_t)b * c * n, (int64_t)b * c * m, &acc, &maxbits, s);
        if (rc) return rc;
        three_interpolate_grad_det_kernel<<<grid, NN_THREADS, 0, s>>>(c, n, m, grad_out, idx, weight, acc, DetScale{maxbits, 8});
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            cudaFreeAsync(acc, s);
            set_error("three_interpolate_grad: CUDA launch failed: %s", cudaGetErrorString(e));
            return CAPTRA_ERR_CUDA;
        }
        count_launch();
        return det_finish(acc, maxbits, 8, (int64_t)b * c * m, grad_points, s);
    }
    three_interpolate_grad_kernel<<<grid, NN_THREADS, 0, as_stream(stream)>>>(c, n, m, grad_out, idx, weight, grad_points);
    CAPTRA_CHECK_LAUNCH("three_interpolate_grad");
    return CAPTRA_OK;
}

extern "C" int captra_three_nn_interpolate(int b, int c, int n, int m, const float *unknown,
                                           const float *known, const float *points, float *out,
                                           float *dist, int *idx, float *weight, int point_major,
                                           int64_t out_row_stride, int out_col_offset,
                                           captra_stream_t stream) {
    CAPTRA_REQUIRE(b >= 0 && c >= 0 && m >= 1 && n >= 0, "three_nn_interpolate: bad size");
    if (b == 0 || n == 0) return CAPTRA_OK;
    CAPTRA_REQUIRE(idx && weight, "three_nn_interpolate: idx/weight buffers required");
    CAPTRA_REQUIRE(b <= 65535, "three_nn_interpolate: batch exceeds grid limit");
    cudaStream_t s = as_stream(stream);
    if (unknown) {   // unknown == NULL: idx/weight are inputs (geometry shared between two networks)
        CAPTRA_REQUIRE(known, "three_nn_interpolate: null known");
        dim3 g1(ceil_div(n, NN_THREADS), b);
        three_nn_weights_kernel<<<g1, NN_THREADS, 0, s>>>(n, m, unknown, known, dist, idx, weight);
        CAPTRA_CHECK_LAUNCH("three_nn_weights");
    }
    if (c == 0 || !out) return CAPTRA_OK;
    CAPTRA_REQUIRE(points, "three_nn_interpolate: null points");
    if (point_major) {
        dim3 g2(ceil_div(n, NN_THREADS / 32), b);
        three_interpolate_pm_kernel<<<g2, NN_THREADS, 0, s>>>(c, m, n, points, idx, weight, out,
                                                               out_row_stride > 0 ? out_row_stride : c, out_col_offset);
        CAPTRA_CHECK_LAUNCH("three_interpolate_pm");
    } else {
        CAPTRA_REQUIRE(ceil_div(c, TI_CH_BLOCK) <= 65535, "three_nn_interpolate: grid limit");
        dim3 g2(ceil_div(n, NN_THREADS), ceil_div(c, TI_CH_BLOCK), b);
        three_interpolate_kernel<<<g2, NN_THREADS, 0, s>>>(c, m, n, points, idx, weight, out);
        CAPTRA_CHECK_LAUNCH("three_interpolate");
    }
    return CAPTRA_OK;
}
