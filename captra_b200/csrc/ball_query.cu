// ball_query.cu -- first-K-in-index-order radius search (reference: ball_query_gpu.cu:9-45).
//
// B200 design: the reference gives every centroid ONE thread that walks the whole cloud with
// stride-3 scalar loads and row-strided idx stores.  Here a WARP owns CPW centroids and walks
// the cloud 32 points at a time out of a shared-memory SoA copy of the cloud tile (one
// coalesced pass over xyz per CTA).  Each step is: 3 conflict-free LDS, one distance per
// centroid (exact reference FMA order), one ballot per (centroid, radius); hits are compacted
// with popc(ballot & lanemask_lt) so the K indices of a row are produced in ascending index
// order, exactly as the serial scan does.  Up to CAPTRA_MAX_RADII radii that share the
// centroids (the MSG loop, pointnet_utils.py:228-233) are answered by the same scan.
#include "common.cuh"

namespace captra {

constexpr int BQ_THREADS = 256;
constexpr int BQ_WARPS = BQ_THREADS / 32;
constexpr int BQ_TILE = 4096;  // points staged per tile: 3 * 4096 * 4 B = 48 KB
// component planes are offset by 11 banks so the AoS->SoA transpose store is conflict-free
constexpr int BQ_PLANE = BQ_TILE + 11;

struct BQParams {
    float radius2[CAPTRA_MAX_RADII];
    int nsample[CAPTRA_MAX_RADII];
    int *idx[CAPTRA_MAX_RADII];
};

template <int NR, int CPW>
__global__ void __launch_bounds__(BQ_THREADS)
ball_query_kernel(int n, int m, BQParams prm, const float *__restrict__ new_xyz,
                  const float *__restrict__ xyz) {
    extern __shared__ float smem[];
    float *sx = smem, *sy = smem + BQ_PLANE, *sz = smem + 2 * BQ_PLANE;

    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = lanemask_lt();
    const float *cloud = xyz + (size_t)b * n * 3;

    // centroids of this warp
    float cx[CPW], cy[CPW], cz[CPW];
    int cnt[CPW][NR], first[CPW][NR];
    int *row[CPW][NR];
    const int c0 = (blockIdx.x * BQ_WARPS + warp) * CPW;
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
        const int ci = c0 + c;
        const bool ok = ci < m;
        const float *p = new_xyz + ((size_t)b * m + (ok ? ci : 0)) * 3;
        cx[c] = p[0]; cy[c] = p[1]; cz[c] = p[2];
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            cnt[c][r] = ok ? 0 : prm.nsample[r];  // out-of-range centroids start "full"
            first[c][r] = 0;
            row[c][r] = prm.idx[r] + ((size_t)b * m + (ok ? ci : 0)) * prm.nsample[r];
        }
    }

    for (int tile0 = 0; tile0 < n; tile0 += BQ_TILE) {
        const int tn = min(BQ_TILE, n - tile0);
        if (tile0 > 0) __syncthreads();
        // coalesced AoS read, SoA store
        const float *src = cloud + (size_t)tile0 * 3;
        for (int e = threadIdx.x; e < tn * 3; e += BQ_THREADS) {
            const int pt = e / 3, comp = e - pt * 3;
            smem[comp * BQ_PLANE + pt] = __ldg(src + e);
        }
        __syncthreads();

        bool warp_active = false;
#pragma unroll
        for (int c = 0; c < CPW; ++c)
#pragma unroll
            for (int r = 0; r < NR; ++r) warp_active |= cnt[c][r] < prm.nsample[r];

        for (int base = 0; base < tn && warp_active; base += 32) {
            const int k = base + lane;
            const bool in = k < tn;
            const float x = in ? sx[k] : 0.f, y = in ? sy[k] : 0.f, z = in ? sz[k] : 0.f;
            warp_active = false;
#pragma unroll
            for (int c = 0; c < CPW; ++c) {
                bool act = false;
#pragma unroll
                for (int r = 0; r < NR; ++r) act |= cnt[c][r] < prm.nsample[r];
                if (!act) continue;  // warp-uniform
                const float d2 = sqdist_ref(cx[c], cy[c], cz[c], x, y, z);
#pragma unroll
                for (int r = 0; r < NR; ++r) {
                    const int K = prm.nsample[r];
                    const int have = cnt[c][r];
                    const bool hit = in && (d2 < prm.radius2[r]) && (have < K);
                    const unsigned bal = __ballot_sync(kFull, hit);
                    if (bal) {
                        const int pos = have + __popc(bal & lt);
                        if (hit && pos < K) row[c][r][pos] = tile0 + k;
                        if (have == 0) first[c][r] = tile0 + base + __ffs(bal) - 1;
                        cnt[c][r] = min(K, have + __popc(bal));
                    }
                    warp_active |= cnt[c][r] < K;
                }
            }
        }
    }

    // pad the tail of every non-empty row with its first hit (ball_query_gpu.cu:35-39);
    // empty rows stay as the caller left them.
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
        if (c0 + c >= m) continue;
#pragma unroll
        for (int r = 0; r < NR; ++r) {
            const int have = cnt[c][r], K = prm.nsample[r];
            if (have > 0)
                for (int l = have + lane; l < K; l += 32) row[c][r][l] = first[c][r];
        }
    }
}

template <int NR>
static int launch_bq(int b, int n, int m, const BQParams &prm, const float *new_xyz,
                     const float *xyz, cudaStream_t stream) {
    constexpr int CPW = 4;
    auto kern = ball_query_kernel<NR, CPW>;
    const size_t smem = sizeof(float) * 3 * BQ_PLANE;
    static bool attr_done = false;
    if (!attr_done) {
        CAPTRA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done = true;
    }
    dim3 grid(ceil_div(m, BQ_WARPS * CPW), b);
    kern<<<grid, BQ_THREADS, smem, stream>>>(n, m, prm, new_xyz, xyz);
    CAPTRA_CHECK_LAUNCH("ball_query");
    return CAPTRA_OK;
}

}  // namespace captra

using namespace captra;

extern "C" int captra_ball_query_multi(int b, int n, int m, int nradii, const float *radii_host,
                                       const int *nsamples_host, const float *new_xyz,
                                       const float *xyz, int *const *idx_host_ptrs,
                                       captra_stream_t stream) {
    CAPTRA_REQUIRE(b >= 0 && n >= 0 && m >= 0, "ball_query: negative size");
    CAPTRA_REQUIRE(nradii >= 1 && nradii <= CAPTRA_MAX_RADII, "ball_query: nradii must be 1..%d", CAPTRA_MAX_RADII);
    CAPTRA_REQUIRE(b <= 65535, "ball_query: batch %d exceeds grid.y limit", b);
    if (b == 0 || m == 0 || n == 0) return CAPTRA_OK;
    CAPTRA_REQUIRE(new_xyz && xyz, "ball_query: null input");
    BQParams prm;
    for (int r = 0; r < CAPTRA_MAX_RADII; ++r) {
        const int rr = r < nradii ? r : 0;
        // radius is the Python float cast to fp32, squared in fp32 (ball_query_gpu.cu:23)
        const float rad = radii_host[rr];
        prm.radius2[r] = rad * rad;
        prm.nsample[r] = nsamples_host[rr];
        prm.idx[r] = idx_host_ptrs[rr];
        CAPTRA_REQUIRE(prm.nsample[r] >= 0 && (prm.nsample[r] == 0 || prm.idx[r]), "ball_query: bad nsample/idx for radius %d", rr);
    }
    cudaStream_t s = as_stream(stream);
    switch (nradii) {
        case 1: return launch_bq<1>(b, n, m, prm, new_xyz, xyz, s);
        case 2: return launch_bq<2>(b, n, m, prm, new_xyz, xyz, s);
        case 3: return launch_bq<3>(b, n, m, prm, new_xyz, xyz, s);
        default: return launch_bq<4>(b, n, m, prm, new_xyz, xyz, s);
    }
}

extern "C" int ball_query_kernel_launcher_fast(int b, int n, int m, float radius, int nsample,
                                               const float *new_xyz, const float *xyz, int *idx,
                                               captra_stream_t stream) {
    return captra_ball_query_multi(b, n, m, 1, &radius, &nsample, new_xyz, xyz, &idx, stream);
}
