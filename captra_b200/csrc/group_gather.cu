// group_gather.cu -- index-driven copies and their scatter-add adjoints.
// Reference: group_points_gpu.cu:8-66, sampling_gpu.cu:8-63.
//
// B200 design: these are HBM-write-bound copies.  The reference launches one thread per
// (c, m, k) element and re-reads idx once per channel; here a thread owns 4 consecutive
// (m,k) slots, reads their indices once (int4), and walks a block of channels, so idx traffic
// drops by the channel-block factor and every store is a 16-byte streaming store.  The source
// rows points[b,c,:] (N floats) stay L1/L2 resident.  All offsets are 64-bit (the reference's
// int offsets overflow at B*C*M*K >= 2^31).
#include <algorithm>

#include "common.cuh"
#include "det_accum.cuh"

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

// Shared-memory variant for the big calls.  With the rows in L1 the kernel above is bound by the
// load pipe (four divergent 4-byte loads per 16-byte store, ~12 cycles per warp-load measured on
// cfg5), not by HBM.  Here a CTA copies CB source rows (CB*n floats) into shared memory once and
// serves every gather from there -- a divergent LDS costs ~3.5 cycles (bank conflicts of 32 random
// lanes) -- so the streaming stores become the limit.  blockIdx.x splits the L slots of one
// (cloud, channel block) into chunks when there are too few CTAs otherwise.
constexpr int GS_THREADS = 512;

template <bool VEC4>
__global__ void __launch_bounds__(GS_THREADS)
gather_rows_smem_kernel(int c, int n, int64_t L, int cb, int64_t chunk, const float *__restrict__ points,
                        const int *__restrict__ idx, float *__restrict__ out) {
    extern __shared__ float rows[];            // [cb][n]
    const int b = blockIdx.z;
    const int cblk = blockIdx.y * cb;
    const int nc = min(c, cblk + cb) - cblk;
    const float *src = points + ((size_t)b * c + cblk) * n;
    const int64_t tot = (int64_t)nc * n;
    if (((reinterpret_cast<uintptr_t>(src) & 15) == 0) && (tot % 4 == 0)) {
        for (int64_t i = threadIdx.x; i < tot / 4; i += GS_THREADS)
            reinterpret_cast<float4 *>(rows)[i] = __ldg(reinterpret_cast<const float4 *>(src) + i);
    } else {
        for (int64_t i = threadIdx.x; i < tot; i += GS_THREADS) rows[i] = __ldg(src + i);
    }
    __syncthreads();
    const int *ix = idx + (size_t)b * L;
    float *dst = out + ((size_t)b * c + cblk) * L;
    const int64_t j0 = (int64_t)blockIdx.x * chunk, j1 = min(L, j0 + chunk);
    if (VEC4) {
        for (int64_t j4 = j0 + (int64_t)threadIdx.x * 4; j4 < j1; j4 += GS_THREADS * 4) {
            const int4 id = __ldg(reinterpret_cast<const int4 *>(ix + j4));
#pragma unroll 4
            for (int ci = 0; ci < nc; ++ci) {
                const float *r = rows + (size_t)ci * n;
                st_stream4(reinterpret_cast<float4 *>(dst + (size_t)ci * L + j4), make_float4(r[id.x], r[id.y], r[id.z], r[id.w]));
            }
        }
    } else {
        for (int64_t j = j0 + threadIdx.x; j < j1; j += GS_THREADS) {
            const int id = __ldg(ix + j);
            for (int ci = 0; ci < nc; ++ci) st_stream(dst + (size_t)ci * L + j, rows[(size_t)ci * n + id]);
        }
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

// deterministic variant (det_accum.cuh): same traversal, 64-bit fixed-point integer atomics
__global__ void __launch_bounds__(GG_THREADS)
scatter_add_rows_det_kernel(int c, int n, int64_t L, const float *__restrict__ grad_out,
                            const int *__restrict__ idx, long long *__restrict__ acc, DetScale sc) {
    const int b = blockIdx.z;
    const int cblk = blockIdx.y * GG_CH_BLOCK;
    const int cend = min(c, cblk + GG_CH_BLOCK);
    const int64_t j = (int64_t)blockIdx.x * GG_THREADS + threadIdx.x;
    if (j >= L) return;
    const int e = sc.exponent();
    const int id = __ldg(idx + (size_t)b * L + j);
    for (int ci = cblk; ci < cend; ++ci)
        det_add(acc + ((size_t)b * c + ci) * n + id, __ldg(grad_out + ((size_t)b * c + ci) * L + j), e);
}

static int launch_gather(const char *name, int b, int c, int n, int64_t L, const float *points,
                         const int *idx, float *out, cudaStream_t stream) {
    CAPTRA_REQUIRE(b >= 0 && c >= 0 && n >= 0 && L >= 0, "%s: negative size", name);
    if (b == 0 || c == 0 || L == 0) return CAPTRA_OK;
    CAPTRA_REQUIRE(points && idx && out, "%s: null pointer", name);
    CAPTRA_REQUIRE(b <= 65535 && ceil_div(c, GG_CH_BLOCK) <= 65535, "%s: grid limit", name);
    const bool vec = (L % 4 == 0) && ((reinterpret_cast<uintptr_t>(idx) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    // big calls: rows staged in shared memory (about 96 KB per CTA, two CTAs per SM)
    constexpr size_t kSmemBudget = 96u << 10, kSmemMax = 200u << 10;
    const int64_t work = (int64_t)b * c * L;
    if (work >= (int64_t)1 << 22 && L >= 8 * (int64_t)n && (size_t)n * 4 <= kSmemMax) {
        int cb = (int)std::max<size_t>(1, kSmemBudget / ((size_t)n * 4));
        cb = std::min(std::min(cb, 32), c);
        if (c <= 4 && (size_t)c * n * 4 <= kSmemMax) cb = c;   // xyz-like inputs: idx read once for all channels
        const int nblk = ceil_div(c, cb);
        cb = ceil_div(c, nblk);                // even out the channel blocks
        const size_t smem = (size_t)cb * n * 4;
        // Split L into chunks (each >= 2 rows' worth of slots, so the staging copy stays a small part of a CTA's
        // work) such that the CTA count fills whole waves: slots = SMs x resident CTAs per SM for this smem
        // footprint.  (cfg5 level 1 -- 64 clouds, 196 KB per CTA -- ran 320 CTAs in 3 waves at 72 % fill.)
        int splits = 1;
        const int64_t ctas = (int64_t)b * nblk;
        const int occ = (int)std::max<size_t>(1, std::min<size_t>((size_t)(227u << 10) / (smem + 1024), 2048 / GS_THREADS));
        const int64_t slots = (int64_t)sm_count() * occ;
        const int64_t max_splits = std::max<int64_t>(1, std::min<int64_t>(L / (2 * (int64_t)n), 64));
        if (ctas < 4 * slots) {
            double best = -1.0;
            for (int64_t sp = 1; sp <= max_splits; ++sp) {
                const int64_t t = ctas * sp, waves = ceil_div<int64_t>(t, slots);
                if (waves > 4 && sp > 1) break;
                const double fill = (double)t / (double)(waves * slots);
                if (fill > best + 1e-9) { best = fill; splits = (int)sp; }
            }
        }
        int64_t chunk = ceil_div<int64_t>(ceil_div<int64_t>(L, splits), 4) * 4;
        splits = (int)ceil_div<int64_t>(L, chunk);
        if (nblk <= 65535) {
            auto kern = vec ? gather_rows_smem_kernel<true> : gather_rows_smem_kernel<false>;
            // per-device attribute: set on every launch (no process-wide flag)
            CAPTRA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemMax));
            kern<<<dim3(splits, nblk, b), GS_THREADS, smem, stream>>>(c, n, L, cb, chunk, points, idx, out);
            CAPTRA_CHECK_LAUNCH(name);
            return CAPTRA_OK;
        }
    }
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
    if (det_enabled()) {
        long long *acc = nullptr;
        unsigned *maxbits = nullptr;
        int rc = det_begin(grad_out, (int64_t)b * c * L, (int64_t)b * c * n, &acc, &maxbits, stream);
        if (rc) return rc;
        scatter_add_rows_det_kernel<<<grid, GG_THREADS, 0, stream>>>(c, n, L, grad_out, idx, acc, DetScale{maxbits, 0});
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) {
            cudaFreeAsync(acc, stream);
            set_error("%s: CUDA launch failed: %s", name, cudaGetErrorString(e));
            return CAPTRA_ERR_CUDA;
        }
        count_launch();
        return det_finish(acc, maxbits, 0, (int64_t)b * c * n, grad_points, stream);
    }
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
