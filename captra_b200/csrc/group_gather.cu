// group_gather.cu -- index-driven copies and their scatter-add adjoints.
// Reference: group_points_gpu.cu:8-66, sampling_gpu.cu:8-63.
//
// B200 design: these are HBM-write-bound copies.  The reference launches one thread per
// (c, m, k) element and re-reads idx once per channel; here a thread owns 4 consecutive
// (m,k) slots, reads their indices once (int4), and walks a block of channels, so idx traffic
// drops by the channel-block factor and every store is a 16-byte streaming store.  The source
// rows points[b,c,:] (N floats) stay L1/L2 resident.  All offsets are 64-bit (the reference's
// int offsets overflow at B*C*M*K >= 2^31).
#include "common.cuh"

namespace captra {

constexpr int GG_THREADS = 256;
constexpr int GG_CH_BLOCK = 8;  // channels walked by one thread

// out[b,c,j] = points[b,c,idx[b,j]] for j in [0, L) ; L = npoints*nsample (group) or npoints
template <bool VEC4>
__global__ void __launch_bounds__(GG_THREADS)
gather_rows_kernel(int c, int n, int64_t L, const float *__restrict__ points,
                   const int *__restrict__ idx, float *__restrict__ out) {
    const int b = blockIdx.z;
    const int cblk = blockIdx.y * GG_CH_BLOCK;
    const int cend = min(c, cblk + GG_CH_BLOCK);
    const int *ix = idx + (size_t)b * L;
    if (VEC4) {
        const int64_t j4 = ((int64_t)blockIdx.x * GG_THREADS + threadIdx.x) * 4;
        if (j4 >= L) return;
        const int4 id = __ldg(reinterpret_cast<const int4 *>(ix + j4));
        for (int ci = cblk; ci < cend; ++ci) {
            const float *src = points + ((size_t)b * c + ci) * n;
            float4 v;
            v.x = __ldg(src + id.x); v.y = __ldg(src + id.y);
            v.z = __ldg(src + id.z); v.w = __ldg(src + id.w);
            st_stream4(reinterpret_cast<float4 *>(out + ((size_t)b * c + ci) * L + j4), v);
        }
    } else {
        const int64_t j = (int64_t)blockIdx.x * GG_THREADS + threadIdx.x;
        if (j >= L) return;
        const int id = __ldg(ix + j);
        for (int ci = cblk; ci < cend; ++ci)
            st_stream(out + ((size_t)b * c + ci) * L + j, __ldg(points + ((size_t)b * c + ci) * n + id));
    }
}

// grad_points[b,c,idx[b,j]] += grad_out[b,c,j]
__global__ void __launch_bounds__(GG_THREADS)
scatter_add_rows_kernel(int c, int n, int64_t L, const float *__restrict__ grad_out,
                        const int *__restrict__ idx, float *__restrict__ grad_points) {
    const int b = blockIdx.z;
    const int cblk = blockIdx.y * GG_CH_BLOCK;
    const int cend = min(c, cblk + GG_CH_BLOCK);
    const int64_t j = (int64_t)blockIdx.x * GG_THREADS + threadIdx.x;
    if (j >= L) return;
    const int id = __ldg(idx + (size_t)b * L + j);
    for (int ci = cblk; ci < cend; ++ci)
        atomicAdd(grad_points + ((size_t)b * c + ci) * n + id, __ldg(grad_out + ((size_t)b * c + ci) * L + j));
}

static int launch_gather(const char *name, int b, int c, int n, int64_t L, const float *points,
                         const int *idx, float *out, cudaStream_t stream) {
    CAPTRA_REQUIRE(b >= 0 && c >= 0 && n >= 0 && L >= 0, "%s: negative size", name);
    if (b == 0 || c == 0 || L == 0) return CAPTRA_OK;
    CAPTRA_REQUIRE(points && idx && out, "%s: null pointer", name);
    CAPTRA_REQUIRE(b <= 65535 && ceil_div(c, GG_CH_BLOCK) <= 65535, "%s: grid limit", name);
    const bool vec = (L % 4 == 0) && ((reinterpret_cast<uintptr_t>(idx) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    if (vec) {
        dim3 grid((unsigned)ceil_div<int64_t>(L / 4, GG_THREADS), ceil_div(c, GG_CH_BLOCK), b);
        gather_rows_kernel<true><<<grid, GG_THREADS, 0, stream>>>(c, n, L, points, idx, out);
    } else {
        dim3 grid((unsigned)ceil_div<int64_t>(L, GG_THREADS), ceil_div(c, GG_CH_BLOCK), b);
        gather_rows_kernel<false><<<grid, GG_THREADS, 0, stream>>>(c, n, L, points, idx, out);
    }
    CAPTRA_CHECK_LAUNCH(name);
    return CAPTRA_OK;
}

static int launch_scatter(const char *name, int b, int c, int n, int64_t L, const float *grad_out,
                          const int *idx, float *grad_points, cudaStream_t stream) {
    CAPTRA_REQUIRE(b >= 0 && c >= 0 && n >= 0 && L >= 0, "%s: negative size", name);
    if (b == 0 || c == 0 || L == 0) return CAPTRA_OK;
    CAPTRA_REQUIRE(grad_out && idx && grad_points, "%s: null pointer", name);
    CAPTRA_REQUIRE(b <= 65535 && ceil_div(c, GG_CH_BLOCK) <= 65535, "%s: grid limit", name);
    dim3 grid((unsigned)ceil_div<int64_t>(L, GG_THREADS), ceil_div(c, GG_CH_BLOCK), b);
    scatter_add_rows_kernel<<<grid, GG_THREADS, 0, stream>>>(c, n, L, grad_out, idx, grad_points);
    CAPTRA_CHECK_LAUNCH(name);
    return CAPTRA_OK;
}

}  // namespace captra

using namespace captra;

extern "C" int group_points_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample,
                                                 const float *points, const int *idx, float *out,
                                                 captra_stream_t stream) {
    return launch_gather("group_points", b, c, n, (int64_t)npoints * nsample, points, idx, out, as_stream(stream));
}

extern "C" int group_points_grad_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample,
                                                      const float *grad_out, const int *idx,
                                                      float *grad_points, captra_stream_t stream) {
    return launch_scatter("group_points_grad", b, c, n, (int64_t)npoints * nsample, grad_out, idx, grad_points, as_stream(stream));
}

extern "C" int gather_points_kernel_launcher_fast(int b, int c, int n, int npoints,
                                                  const float *points, const int *idx, float *out,
                                                  captra_stream_t stream) {
    return launch_gather("gather_points", b, c, n, npoints, points, idx, out, as_stream(stream));
}

extern "C" int gather_points_grad_kernel_launcher_fast(int b, int c, int n, int npoints,
                                                       const float *grad_out, const int *idx,
                                                       float *grad_points, captra_stream_t stream) {
    return launch_scatter("gather_points_grad", b, c, n, npoints, grad_out, idx, grad_points, as_stream(stream));
}
