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

#include <math.h>
#include <stdlib.h>

#include <mutex>

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

// PIPED: the centroids are being produced by a furthest-point-sampling kernel that is still running
// (captra_fps_ball_query): the grid is laid out (cloud, centroid block) so that blocks become resident in the order
// FPS finishes them, a CTA stages its cloud tile first and only then waits until `progress[b]` covers its centroids,
// and the centroids are read past the (non-coherent) L1.
template <int NR, int CPW, bool PIPED>
__global__ void __launch_bounds__(BQ_THREADS)
ball_query_kernel(int n, int m, BQParams prm, const float *__restrict__ new_xyz_,
                  const float *__restrict__ xyz, const int *progress) {
    extern __shared__ float smem[];
    float *sx = smem, *sy = smem + BQ_PLANE, *sz = smem + 2 * BQ_PLANE;
    const float *new_xyz = new_xyz_;

    const int b = PIPED ? blockIdx.x : blockIdx.y;
    const int cblock = PIPED ? blockIdx.y : blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt = lanemask_lt();
    const float *cloud = xyz + (size_t)b * n * 3;

    // centroids of this warp
    float cx[CPW], cy[CPW], cz[CPW];
    int cnt[CPW][NR], first[CPW][NR];
    int *row[CPW][NR];
    const int c0 = (cblock * BQ_WARPS + warp) * CPW;
    if (PIPED) {
        // stage the first cloud tile while FPS is still picking this block's centroids
        const int tn0 = min(BQ_TILE, n);
        for (int e = threadIdx.x; e < tn0 * 3; e += BQ_THREADS) {
            const int pt = e / 3, comp = e - pt * 3;
            smem[comp * BQ_PLANE + pt] = __ldg(cloud + e);
        }
        if (threadIdx.x == 0) {
            const int need = min(m, (cblock + 1) * BQ_WARPS * CPW);
            int have;
            do {
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(have) : "l"(progress + b) : "memory");
                if (have < need) __nanosleep(500);
            } while (have < need);
        }
        __syncthreads();
    }
#pragma unroll
    for (int c = 0; c < CPW; ++c) {
        const int ci = c0 + c;
        const bool ok = ci < m;
        const float *p = new_xyz + ((size_t)b * m + (ok ? ci : 0)) * 3;
        if (PIPED) { cx[c] = __ldcg(p); cy[c] = __ldcg(p + 1); cz[c] = __ldcg(p + 2); }
        else { cx[c] = p[0]; cy[c] = p[1]; cz[c] = p[2]; }
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
        if (!PIPED || tile0 > 0) {      // (PIPED: the first tile was staged before the wait)
            // coalesced AoS read, SoA store
            const float *src = cloud + (size_t)tile0 * 3;
            for (int e = threadIdx.x; e < tn * 3; e += BQ_THREADS) {
                const int pt = e / 3, comp = e - pt * 3;
                smem[comp * BQ_PLANE + pt] = __ldg(src + e);
            }
            __syncthreads();
        }

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
                     const float *xyz, cudaStream_t stream, const int *progress = nullptr) {
    constexpr int CPW = 4;
    const size_t smem = sizeof(float) * 3 * BQ_PLANE;
    if (progress) {
        // two centroids per warp here (16 per CTA): the blocks that become ready last -- when FPS has already finished --
        // are the tail of the pipeline, and a block's latency is proportional to the centroids a warp walks serially
        constexpr int CPW = 2;
        // programmatic dependent launch: this grid may start while the preceding kernel of the stream (FPS, which
        // has signalled griddepcontrol.launch_dependents) is still running; the data dependency is the progress counter
        auto kern = ball_query_kernel<NR, CPW, true>;
        CAPTRA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(b, ceil_div(m, BQ_WARPS * CPW));
        cfg.blockDim = dim3(BQ_THREADS);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        CAPTRA_CUDA(cudaLaunchKernelEx(&cfg, kern, n, m, prm, new_xyz, xyz, progress));
        CAPTRA_CHECK_LAUNCH("ball_query(piped)");
        return CAPTRA_OK;
    }
    auto kern = ball_query_kernel<NR, CPW, false>;
    // per-device attribute, cheap host-side call: set it on every launch (no process-wide flag)
    CAPTRA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(m, BQ_WARPS * CPW), b);
    kern<<<grid, BQ_THREADS, smem, stream>>>(n, m, prm, new_xyz, xyz, nullptr);
    CAPTRA_CHECK_LAUNCH("ball_query");
    return CAPTRA_OK;
}


// ------------------------------------------------------------------------------------------------
// Grid-accelerated query for large clouds (single radius).  The brute-force scan costs O(M*N)
// distance tests; at N = 16384 (BASELINE cfg5) that is 4.3 G tests per call and two orders of magnitude
// away from the HBM roofline of the op.  Here each cloud is binned into a uniform grid with cell
// size h >= r(1+1e-5) (so every hit lies in the 27 cells around the centroid), a warp tests only
// those cells with the SAME exact-order fp32 distance, and the hit indices are sorted ascending in
// a per-warp bitmap and read back in ascending order, which reproduces "the first K hits in index
// order" bit for bit.  Dense balls
// (more than BQG_CAP candidates, where the serial scan would exit early anyway) fall back to the
// early-exit scan inside the same warp.
// ------------------------------------------------------------------------------------------------
constexpr int BQG_DIM = 32;                    // max cells per axis
constexpr int BQG_MAXCELL = BQG_DIM * BQG_DIM * BQG_DIM;
constexpr int BQG_CAP = 256;                   // more hits than this: the early-exit scan is cheaper
constexpr int BQG_MAX_SMEM = 200 << 10;        // bitmap budget of the one-warp-per-CTA variant

static inline int bq_grid_wpl_log2(int n) {    // bitmap words per lane (power of two) covering n bits
    int l = 0;
    while ((1024ll << l) < n) ++l;
    return l;
}
// A lane's run of wpl bitmap words is padded so that the lanes' runs start in different banks: wpl + 4 keeps a run 16-byte
// aligned (the clear and the prefix pass move four words per instruction; conflict-free per quarter warp), wpl | 1 below that.
__host__ __device__ static inline int bq_grid_stride(int wpl) { return wpl >= 4 ? wpl + 4 : (wpl | 1); }
// per-warp working set: bitmap words + per-word prefix counts (16 bit) + per-lane offsets + hit list + staged row
static inline size_t bq_grid_warp_bytes(int n, int nsample) {
    const size_t words = 32 * (size_t)bq_grid_stride(1 << bq_grid_wpl_log2(n));
    return (sizeof(int) * (words + 32 + (BQG_CAP + 32) + (size_t)nsample) + sizeof(unsigned short) * words + 15) & ~(size_t)15;
}
static inline bool bq_grid_fits(int n, int nsample) { return bq_grid_warp_bytes(n, nsample) <= (size_t)BQG_MAX_SMEM; }

struct BQGrid {          // per cloud, written by the build kernel
    float minx, miny, minz, inv_h;
    int dx, dy, dz, ncell;
};

// One CTA per cloud: bbox -> cell histogram (smem) -> exclusive scan -> scatter of (x,y,z,index).
__global__ void __launch_bounds__(1024)
bq_grid_build_kernel(int n, float radius, const float *__restrict__ xyz, BQGrid *__restrict__ grids,
                     int *__restrict__ cell_start /* [B][BQG_MAXCELL+1] */, float4 *__restrict__ sorted /* [B][n] */) {
    extern __shared__ int counts[];            // [BQG_MAXCELL + 1]
    __shared__ float red[6][32];
    __shared__ BQGrid g;
    __shared__ int warp_tot[32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *cloud = xyz + (size_t)b * n * 3;
    // ---- bounding box over finite coordinates
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int k = tid; k < n; k += 1024) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = __ldg(cloud + k * 3 + c);
            if (isfinite(v)) { lo[c] = fminf(lo[c], v); hi[c] = fmaxf(hi[c], v); }
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(kFull, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(kFull, hi[c], o));
        }
        if (lane == 0) { red[c][warp] = lo[c]; red[3 + c][warp] = hi[c]; }
    }
    for (int i = tid; i <= BQG_MAXCELL; i += 1024) counts[i] = 0;
    __syncthreads();
    if (tid == 0) {
        float mn[3], mx[3];
        for (int c = 0; c < 3; ++c) {
            mn[c] = INFINITY; mx[c] = -INFINITY;
            for (int w = 0; w < 32; ++w) { mn[c] = fminf(mn[c], red[c][w]); mx[c] = fmaxf(mx[c], red[3 + c][w]); }
            if (!(mn[c] <= mx[c])) { mn[c] = 0.f; mx[c] = 0.f; }   // no finite coordinate at all
        }
        const float ext = fmaxf(fmaxf(mx[0] - mn[0], mx[1] - mn[1]), mx[2] - mn[2]);
        float h = radius * (1.0f + 1e-5f);
        h = fmaxf(h, ext / (float)(BQG_DIM - 1));      // at most BQG_DIM cells per axis
        if (!(h > 0.f) || !isfinite(h)) h = 1.f;
        g.minx = mn[0]; g.miny = mn[1]; g.minz = mn[2]; g.inv_h = 1.0f / h;
        g.dx = min(BQG_DIM, (int)((mx[0] - mn[0]) * g.inv_h) + 1);
        g.dy = min(BQG_DIM, (int)((mx[1] - mn[1]) * g.inv_h) + 1);
        g.dz = min(BQG_DIM, (int)((mx[2] - mn[2]) * g.inv_h) + 1);
        g.ncell = g.dx * g.dy * g.dz;
        grids[b] = g;
    }
    __syncthreads();
    auto cell_of = [&](float x, float y, float z) -> int {
        // non-finite coordinates land in cell 0; they can never pass the distance test
        int ix = isfinite(x) ? (int)((x - g.minx) * g.inv_h) : 0;
        int iy = isfinite(y) ? (int)((y - g.miny) * g.inv_h) : 0;
        int iz = isfinite(z) ? (int)((z - g.minz) * g.inv_h) : 0;
        ix = min(max(ix, 0), g.dx - 1); iy = min(max(iy, 0), g.dy - 1); iz = min(max(iz, 0), g.dz - 1);
        return (iz * g.dy + iy) * g.dx + ix;
    };
    for (int k = tid; k < n; k += 1024)
        atomicAdd(&counts[cell_of(__ldg(cloud + k * 3), __ldg(cloud + k * 3 + 1), __ldg(cloud + k * 3 + 2))], 1);
    __syncthreads();
    // ---- exclusive scan of counts[0..ncell) in place (each thread owns a contiguous run)
    const int ncell = g.ncell;
    const int per = (ncell + 1023) / 1024;
    const int beg = min(tid * per, ncell), end = min(beg + per, ncell);
    int sum = 0;
    for (int i = beg; i < end; ++i) sum += counts[i];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int v = warp_tot[lane], inc2 = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFull, inc2, o);
            if (lane >= o) inc2 += t;
        }
        warp_tot[lane] = inc2 - v;   // exclusive prefix of warp totals
    }
    __syncthreads();
    int run = warp_tot[warp] + incl - sum;
    int *cs = cell_start + (size_t)b * (BQG_MAXCELL + 1);
    for (int i = beg; i < end; ++i) {
        const int c = counts[i];
        counts[i] = run;       // becomes the scatter cursor
        cs[i] = run;
        run += c;
    }
    if (tid == 1023) cs[ncell] = n;
    __syncthreads();
    float4 *dst = sorted + (size_t)b * n;
    for (int k = tid; k < n; k += 1024) {
        const float x = __ldg(cloud + k * 3), y = __ldg(cloud + k * 3 + 1), z = __ldg(cloud + k * 3 + 2);
        const int pos = atomicAdd(&counts[cell_of(x, y, z)], 1);
        dst[pos] = make_float4(x, y, z, __int_as_float(k));
    }
}

// One warp per centroid.  Hits are recorded as bits of a per-warp index bitmap in shared memory, so
// the rank of a hit in index order is the number of set bits below its own -- no sort.  Lane l owns
// the wpl consecutive words [l*wpl, (l+1)*wpl) (padded, bq_grid_stride).
// GC: channels grouped by the query warp itself (0: none, 3: compile-time, -1: gc at run time, <= 8).
template <int GC>
__device__ __forceinline__ void bq_grid_centroid(int b, int ci, int lane, unsigned char *ws, int n, int m, float radius2, int nsample,
                                                 int wpl_log2, const float *__restrict__ new_xyz, const float *__restrict__ xyz,
                                                 const BQGrid *__restrict__ grids, const int *__restrict__ cell_start,
                                                 const float4 *__restrict__ sorted, int *__restrict__ idx, int gc,
                                                 const float *__restrict__ gpoints, float *__restrict__ grouped) {
    const int wpl = 1 << wpl_log2, stride = bq_grid_stride(wpl);
    unsigned *bm = reinterpret_cast<unsigned *>(ws);                                     // one bit per point index
    unsigned short *pre = reinterpret_cast<unsigned short *>(bm + 32 * stride);          // set bits before each word, within its lane's run
    int *loff = reinterpret_cast<int *>(pre + 32 * stride);                              // set bits before each lane's run of words
    int *hits = loff + 32;                                                               // the hits in the order they were found
    int *stage = hits + BQG_CAP + 32;                                                    // the row, in index order
    unsigned *mine = bm + lane * stride;
    if (wpl >= 4)
        for (int j = 0; j < wpl; j += 4) *reinterpret_cast<uint4 *>(mine + j) = make_uint4(0u, 0u, 0u, 0u);
    else
        for (int j = 0; j < wpl; ++j) mine[j] = 0u;
    const unsigned lt = lanemask_lt();
    const BQGrid g = grids[b];
    const float *cp = new_xyz + ((size_t)b * m + ci) * 3;
    const float cx = __ldg(cp), cy = __ldg(cp + 1), cz = __ldg(cp + 2);
    int *row = idx + ((size_t)b * m + ci) * nsample;
    const int ngc = GC >= 0 ? GC : gc;
    // Writes the row (keep staged indices, padded with the first hit) and, fused grouping (captra_ball_query_group, few
    // channels), grouped[b, c, ci, :] = points[b, c, row[:]] -- K contiguous floats per channel, no second launch, no idx
    // re-read.  An empty ball keeps the caller's zeros and groups index 0.
    auto emit = [&](int keep) {
        const int first = keep ? stage[0] : 0;
        const float *src = GC ? gpoints + (size_t)b * ngc * n : nullptr;
        float *dst = GC ? grouped + ((size_t)b * ngc * m + ci) * nsample : nullptr;
        const size_t dstride = (size_t)m * nsample;
        for (int l0 = lane; l0 < nsample; l0 += 64) {      // two slots per lane and pass: all their loads in flight before the first store
            const int l1 = l0 + 32;
            const bool two = l1 < nsample;
            const int k0 = l0 < keep ? stage[l0] : first, k1 = two && l1 < keep ? stage[l1] : first;
            if (keep) {
                row[l0] = k0;
                if (two) row[l1] = k1;
            }
            if (GC) {
                constexpr int MAXC = GC > 0 ? GC : 8;
                float v0[MAXC], v1[MAXC];
#pragma unroll
                for (int c = 0; c < MAXC; ++c)
                    if (c < ngc) {
                        v0[c] = __ldg(src + (size_t)c * n + k0);
                        v1[c] = __ldg(src + (size_t)c * n + k1);
                    }
#pragma unroll
                for (int c = 0; c < MAXC; ++c)
                    if (c < ngc) {
                        st_stream(dst + c * dstride + l0, v0[c]);
                        if (two) st_stream(dst + c * dstride + l1, v1[c]);
                    }
            }
        }
    };
    const int *cs = cell_start + (size_t)b * (BQG_MAXCELL + 1);
    const float4 *pts = sorted + (size_t)b * n;
    __syncwarp();

    int cnt = 0;
    bool overflow = false;
    // the centroid's own cell, unclamped: a centroid outside the box still sees the right neighbours (computed ahead of
    // the branch so that the grid descriptor and the centroid are fetched in the same round trip)
    const int ix = (int)floorf((cx - g.minx) * g.inv_h), iy = (int)floorf((cy - g.miny) * g.inv_h),
              iz = (int)floorf((cz - g.minz) * g.inv_h);
    if (isfinite(cx) && isfinite(cy) && isfinite(cz)) {
        // Cells x0..x1 of one (y,z) row are contiguous in the binned array, so the 27 cells are nine
        // segments.  Lane r < 9 fetches segment r's bounds (one round trip for all nine), and the
        // candidates are then walked as one flat list, 32 per step, with the next step's points in flight.
        int seg_beg = 0, seg_len = 0;
        if (lane < 9) {
            const int zz = iz - 1 + lane / 3, yy = iy - 1 + lane % 3;
            const int x0 = max(ix - 1, 0), x1 = min(ix + 1, g.dx - 1);
            if (zz >= 0 && zz < g.dz && yy >= 0 && yy < g.dy && x0 <= x1) {
                const int rowc = (zz * g.dy + yy) * g.dx;
                seg_beg = __ldg(cs + rowc + x0);
                seg_len = __ldg(cs + rowc + x1 + 1) - seg_beg;
            }
        }
        int incl = seg_len;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            const int t = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += t;
        }
        const int total = __shfl_sync(kFull, incl, 8);
        // flat position f lies in segment r iff se[r-1] <= f < se[r]; its binned slot is f + dl[r]
        const int delta = seg_beg - (incl - seg_len);
        int dl[9], se[8];
#pragma unroll
        for (int r = 0; r < 9; ++r) dl[r] = __shfl_sync(kFull, delta, r);
#pragma unroll
        for (int r = 0; r < 8; ++r) se[r] = __shfl_sync(kFull, incl, r);
        auto fetch = [&](int f, float4 &p) -> bool {
            if (f >= total) return false;
            int d = dl[0];
#pragma unroll
            for (int r = 1; r < 9; ++r) d = f >= se[r - 1] ? dl[r] : d;
            p = __ldg(pts + f + d);
            return true;
        };
        // two steps of candidates in flight: the warp is latency-bound (shared memory caps it at ~40 warps per SM)
        float4 p_n1 = make_float4(0.f, 0.f, 0.f, 0.f), p_n2 = p_n1;
        bool have_n1 = fetch(lane, p_n1), have_n2 = fetch(32 + lane, p_n2);
        for (int t = 0; t < total; t += 32) {
            const float4 p = p_n1;
            const bool have = have_n1;
            p_n1 = p_n2;
            have_n1 = have_n2;
            have_n2 = fetch(t + 64 + lane, p_n2);
            const bool hit = have && sqdist_ref(cx, cy, cz, p.x, p.y, p.z) < radius2;
            const unsigned bal = __ballot_sync(kFull, hit);
            if (hit) {
                const int k = __float_as_int(p.w), w = k >> 5;
                atomicOr(bm + (w >> wpl_log2) * stride + (w & (wpl - 1)), 1u << (k & 31));
                hits[cnt + __popc(bal & lt)] = k;
            }
            cnt += __popc(bal);
            if (cnt > BQG_CAP) { overflow = true; break; }
        }
    }
    if (overflow) {
        // dense ball: the reference's serial scan exits early here; do exactly that (warp-wide)
        const float *cloud = xyz + (size_t)b * n * 3;
        int have = 0, first = 0;
        for (int base = 0; base < n && have < nsample; base += 32) {
            const int k = base + lane;
            bool hit = false;
            if (k < n) hit = sqdist_ref(cx, cy, cz, __ldg(cloud + k * 3), __ldg(cloud + k * 3 + 1), __ldg(cloud + k * 3 + 2)) < radius2;
            const unsigned bal = __ballot_sync(kFull, hit);
            if (bal) {
                const int pos = have + __popc(bal & lt);
                if (hit && pos < nsample) row[pos] = k;
                if (have == 0) first = base + __ffs(bal) - 1;
                have = min(nsample, have + __popc(bal));
            }
        }
        __syncwarp();
        if (have == 0) { emit(0); return; }
        // (rows of this rare path are re-read from global: the same warp wrote them)
        for (int l = have + lane; l < nsample; l += 32) row[l] = first;
        if (GC) {
            __syncwarp();
            for (int l = lane; l < nsample; l += 32) {
                const int k = row[l];
                for (int c = 0; c < ngc; ++c)
                    st_stream(grouped + (((size_t)b * ngc + c) * m + ci) * nsample + l, __ldg(gpoints + ((size_t)b * ngc + c) * n + k));
            }
        }
        return;
    }
    if (cnt == 0) { emit(0); return; }            // empty ball: the caller's zeros stay
    __syncwarp();
    // ---- index order without a sort: the rank of hit k is the number of set bits below bit k.  Per-word prefix counts
    // within each lane's run, a warp scan across the runs, then every recorded hit looks its rank up and drops itself
    // into the staged row.
    unsigned short *pmine = pre + lane * stride;
    int c = 0;
    if (wpl >= 4) {
        for (int j = 0; j < wpl; j += 4) {
            const uint4 w = *reinterpret_cast<const uint4 *>(mine + j);
            const int c1 = c + __popc(w.x), c2 = c1 + __popc(w.y), c3 = c2 + __popc(w.z);
            *reinterpret_cast<uint2 *>(pmine + j) = make_uint2((unsigned)c | ((unsigned)c1 << 16), (unsigned)c2 | ((unsigned)c3 << 16));
            c = c3 + __popc(w.w);
        }
    } else {
        for (int j = 0; j < wpl; ++j) {
            pmine[j] = (unsigned short)c;
            c += __popc(mine[j]);
        }
    }
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFull, incl, o);
        if (lane >= o) incl += t;
    }
    loff[lane] = incl - c;
    __syncwarp();
    for (int i = lane; i < cnt; i += 32) {
        const int k = hits[i], w = k >> 5, own = w >> wpl_log2, slot = own * stride + (w & (wpl - 1));
        const int r = loff[own] + pre[slot] + __popc(bm[slot] & ((1u << (k & 31)) - 1u));
        if (r < nsample) stage[r] = k;
    }
    __syncwarp();
    emit(min(cnt, nsample));
}

// One warp per centroid (a persistent variant -- fewer CTAs, each warp striding over the centroids -- measured 4 % slower on
// BASELINE cfg5 level 1, and capping the registers for 10 CTAs per SM spills: 364 -> 462 us).
template <int QW, int GC>
__global__ void __launch_bounds__(QW * 32)
bq_grid_query_kernel(int n, int m, float radius2, int nsample, int wpl_log2, const float *__restrict__ new_xyz,
                     const float *__restrict__ xyz, const BQGrid *__restrict__ grids,
                     const int *__restrict__ cell_start, const float4 *__restrict__ sorted, int *__restrict__ idx,
                     int gc, const float *__restrict__ gpoints, float *__restrict__ grouped, int warp_bytes) {
    extern __shared__ __align__(16) unsigned char bq_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ci = blockIdx.x * QW + warp;
    if (ci < m)
        bq_grid_centroid<GC>(blockIdx.y, ci, lane, bq_smem + (size_t)warp * warp_bytes, n, m, radius2, nsample, wpl_log2, new_xyz, xyz,
                             grids, cell_start, sorted, idx, gc, gpoints, grouped);
}

static int launch_bq_grid(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz,
                          int *idx, cudaStream_t stream, int gc = 0, const float *gpoints = nullptr, float *grouped = nullptr) {
    // stream-ordered scratch: grid descriptors, cell starts, binned points
    const size_t sz_g = sizeof(BQGrid) * (size_t)b;
    const size_t sz_c = sizeof(int) * (size_t)b * (BQG_MAXCELL + 1);
    const size_t sz_s = sizeof(float4) * (size_t)b * n;
    const size_t off_c = (sz_g + 255) & ~(size_t)255, off_s = (off_c + sz_c + 255) & ~(size_t)255;
    // Scratch comes from a stream-ordered pool OWNED by this library (one per device, created on first use, release
    // threshold raised so freed scratch is reused by the next frame instead of going back to the driver at every
    // synchronisation).  The process's default pool -- torch's, or anybody else's -- is not touched.
    static std::mutex pool_mu;
    static cudaMemPool_t pools[64] = {};
    int dev = 0;
    CAPTRA_CUDA(cudaGetDevice(&dev));
    CAPTRA_REQUIRE(dev >= 0 && dev < 64, "ball_query: device index %d out of range", dev);
    cudaMemPool_t pool;
    {
        std::lock_guard<std::mutex> lock(pool_mu);
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
    uint8_t *ws = nullptr;
    CAPTRA_CUDA(cudaMallocFromPoolAsync(reinterpret_cast<void **>(&ws), off_s + sz_s, pool, stream));
    struct Guard {              // every early return below hands the scratch back
        uint8_t *p; cudaStream_t s;
        ~Guard() { if (p) cudaFreeAsync(p, s); }
    } guard{ws, stream};
    BQGrid *grids = reinterpret_cast<BQGrid *>(ws);
    int *cell_start = reinterpret_cast<int *>(ws + off_c);
    float4 *sorted = reinterpret_cast<float4 *>(ws + off_s);
    const size_t smem = sizeof(int) * (BQG_MAXCELL + 1);
    CAPTRA_CUDA(cudaFuncSetAttribute(bq_grid_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bq_grid_build_kernel<<<b, 1024, smem, stream>>>(n, radius, xyz, grids, cell_start, sorted);
    CAPTRA_CHECK_LAUNCH("ball_query(grid build)");
    const int wpl_log2 = bq_grid_wpl_log2(n);
    const size_t per_warp = bq_grid_warp_bytes(n, nsample);
    auto launch = [&](auto kern, int qw) -> int {
        CAPTRA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, qw == 4 ? (64 << 10) : BQG_MAX_SMEM));
        kern<<<dim3(ceil_div(m, qw), b), qw * 32, qw * per_warp, stream>>>(n, m, radius * radius, nsample, wpl_log2, new_xyz, xyz, grids,
                                                                         cell_start, sorted, idx, gc, gpoints, grouped, (int)per_warp);
        return CAPTRA_OK;
    };
    const bool four = 4 * per_warp <= (64u << 10);
    int rc;
    if (!grouped || gc == 0) rc = four ? launch(bq_grid_query_kernel<4, 0>, 4) : launch(bq_grid_query_kernel<1, 0>, 1);
    else if (gc == 3) rc = four ? launch(bq_grid_query_kernel<4, 3>, 4) : launch(bq_grid_query_kernel<1, 3>, 1);
    else rc = four ? launch(bq_grid_query_kernel<4, -1>, 4) : launch(bq_grid_query_kernel<1, -1>, 1);
    if (rc) return rc;
    CAPTRA_CHECK_LAUNCH("ball_query(grid query)");
    return CAPTRA_OK;           // ~Guard frees the scratch (stream-ordered, after the query kernel)
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
    // larger clouds, one radius: binned search (same hits, same order).  Measured crossover against the
    // scan on B200 is between N = 1024 and 4096 (profiles/r01_stress_cfg5_*.jsonl); a huge
    // CAPTRA_BQ_GRID_MIN_N forces the scan.
    static const int grid_min_n = [] { const char *e = getenv("CAPTRA_BQ_GRID_MIN_N"); return e ? atoi(e) : 2048; }();
    if (nradii == 1 && n >= grid_min_n && prm.nsample[0] > 0 && radii_host[0] > 0.f && isfinite(radii_host[0]) &&
        bq_grid_fits(n, prm.nsample[0]))
        return launch_bq_grid(b, n, m, radii_host[0], prm.nsample[0], new_xyz, xyz, prm.idx[0], s);
    switch (nradii) {
        case 1: return launch_bq<1>(b, n, m, prm, new_xyz, xyz, s);
        case 2: return launch_bq<2>(b, n, m, prm, new_xyz, xyz, s);
        case 3: return launch_bq<3>(b, n, m, prm, new_xyz, xyz, s);
        default: return launch_bq<4>(b, n, m, prm, new_xyz, xyz, s);
    }
}

// in fps.cu: FPS that publishes its progress (see captra_fps_ball_query)
namespace captra {
int fps_gather_progress(int b, int n, int m, const float *dataset, int *idxs, float *new_xyz, int *progress, cudaStream_t stream);
}

extern "C" int captra_fps_ball_query(int b, int n, int m, const float *xyz, int *fps_idx, float *new_xyz, int nradii,
                                     const float *radii_host, const int *nsamples_host, int *const *idx_host_ptrs,
                                     int *progress, captra_stream_t stream) {
    CAPTRA_REQUIRE(b >= 0 && n >= 1 && m >= 1, "fps_ball_query: bad sizes");
    CAPTRA_REQUIRE(nradii >= 1 && nradii <= CAPTRA_MAX_RADII, "fps_ball_query: nradii must be 1..%d", CAPTRA_MAX_RADII);
    CAPTRA_REQUIRE(b <= 65535 && n <= 8192, "fps_ball_query: at most 65535 clouds of at most 8192 points (use the two separate calls beyond)");
    if (b == 0) return CAPTRA_OK;
    CAPTRA_REQUIRE(xyz && fps_idx && new_xyz && progress, "fps_ball_query: null pointer");
    BQParams prm;
    for (int r = 0; r < CAPTRA_MAX_RADII; ++r) {
        const int rr = r < nradii ? r : 0;
        const float rad = radii_host[rr];
        prm.radius2[r] = rad * rad;
        prm.nsample[r] = nsamples_host[rr];
        prm.idx[r] = idx_host_ptrs[rr];
        CAPTRA_REQUIRE(prm.nsample[r] >= 0 && (prm.nsample[r] == 0 || prm.idx[r]), "fps_ball_query: bad nsample/idx for radius %d", rr);
    }
    cudaStream_t s = as_stream(stream);
    int rc = fps_gather_progress(b, n, m, xyz, fps_idx, new_xyz, progress, s);
    if (rc) return rc;
    switch (nradii) {
        case 1: return launch_bq<1>(b, n, m, prm, new_xyz, xyz, s, progress);
        case 2: return launch_bq<2>(b, n, m, prm, new_xyz, xyz, s, progress);
        case 3: return launch_bq<3>(b, n, m, prm, new_xyz, xyz, s, progress);
        default: return launch_bq<4>(b, n, m, prm, new_xyz, xyz, s, progress);
    }
}

// in group_gather.cu
extern "C" int group_points_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample, const float *points,
                                                 const int *idx, float *out, captra_stream_t stream);

extern "C" int captra_ball_query_group(int b, int n, int m, int c, float radius, int nsample, const float *new_xyz,
                                       const float *xyz, const float *points, int *idx, float *grouped,
                                       captra_stream_t stream) {
    CAPTRA_REQUIRE(b >= 0 && n >= 0 && m >= 0 && c >= 0 && nsample >= 0, "ball_query_group: negative size");
    CAPTRA_REQUIRE(b <= 65535, "ball_query_group: batch %d exceeds grid.y limit", b);
    if (b == 0 || m == 0 || n == 0 || nsample == 0) return CAPTRA_OK;
    CAPTRA_REQUIRE(new_xyz && xyz && idx && (c == 0 || (points && grouped)), "ball_query_group: null pointer");
    static const int grid_min_n = [] { const char *e = getenv("CAPTRA_BQ_GRID_MIN_N"); return e ? atoi(e) : 2048; }();
    // few channels (SA level 1 groups the coordinates themselves): the query warp writes the grouped rows itself
    if (c >= 1 && c <= 8 && n >= grid_min_n && radius > 0.f && isfinite(radius) && bq_grid_fits(n, nsample))
        return launch_bq_grid(b, n, m, radius, nsample, new_xyz, xyz, idx, as_stream(stream), c, points, grouped);
    // wide rows: the query, then the shared-memory-staged gather (the HBM-bound part, group_gather.cu)
    int rc = captra_ball_query_multi(b, n, m, 1, &radius, &nsample, new_xyz, xyz, &idx, stream);
    if (rc != CAPTRA_OK || c == 0) return rc;
    return group_points_kernel_launcher_fast(b, c, n, m, nsample, points, idx, grouped, stream);
}

extern "C" int ball_query_kernel_launcher_fast(int b, int n, int m, float radius, int nsample,
                                               const float *new_xyz, const float *xyz, int *idx,
                                               captra_stream_t stream) {
    return captra_ball_query_multi(b, n, m, 1, &radius, &nsample, new_xyz, xyz, &idx, stream);
}
