// crop.cu -- the data-side crop of one tracking frame on the device:
//   datasets/nocs_data/nocs_utils.py:5-33           backproject (depth pixels -> camera points, float64)
//   datasets/nocs_data/nocs_data_process.py:92-109  crop_ball_from_pts (ball test, radius growth x1.10 until >= 10
//                                                   points, tiling of small crops, FPS resample)
//   datasets/nocs_data/nocs_data_process.py:148-164 crop_ball_from_depth_image
//   datasets/data_utils.py:138-158                  farthest_point_sample (random 5*npoint subset, then FPS from index 0)
// The reference does this in numpy on the host with one H2D / D2H round trip for the FPS; in `nocs_otf` tracking
// (B = 1) it is the serial bottleneck of a frame.  Here the depth window never leaves the device:
//   crop_classify   per pixel: back-projection in fp64 (same expression order as nocs_utils.py:22-31), distance to
//                   the centre, the first of the ten candidate radii that contains it; per-row histogram
//   crop_select     one CTA: which radius the growth loop stops at (nocs_data_process.py:95-99), the "take
//                   everything" fallback (:101-102), exclusive row offsets, the count n
//   crop_compact    selected pixels -> (point fp64, mask value, pixel index) in the row-major order np.where gives
//   crop_subset     tiled / permuted subset -> the float32 cloud FPS runs on (data_utils.py:147-149)
//   crop_gather     FPS picks -> output points, mask values, raw indices
// Index order, the selected radius and the FPS picks are those of the reference for a given permutation.
#include "common.cuh"

#include <math.h>

namespace captra {

constexpr int CROP_LEVELS = 10;           // for i in range(10): ... radius *= 1.10
constexpr int CROP_THREADS = 128;
constexpr uint8_t CROP_INVALID = 255;     // depth == 0 (or outside the window)

struct CropGeom {
    int H, W, r0, r1, c0, c1;             // image size, window rows r0..r1 and columns c0..c1 (inclusive)
    double kinv[9];                       // inverse intrinsics (numpy.linalg.inv on the host, as the reference)
    double center[3];
    double radii[CROP_LEVELS];            // max(radius, 0.05) * 1.10^i, multiplied sequentially like the loop
    double scale;                         // 0.001
};

// nocs_utils.py:18-33: grid = (col, H - row); xyz = Kinv @ (u, v, 1); pts = xyz * z / xyz[2]; pts[2] = -pts[2]; * scale
__device__ __forceinline__ void backproject_px(const CropGeom &g, int row, int col, float depth, double (&p)[3]) {
    const double u = (double)col, v = (double)(g.H - row);
    double xyz[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) xyz[i] = g.kinv[3 * i] * u + g.kinv[3 * i + 1] * v + g.kinv[3 * i + 2] * 1.0;
    const double z = (double)depth;
#pragma unroll
    for (int i = 0; i < 3; ++i) p[i] = xyz[i] * z / xyz[2];
    p[2] = -p[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) p[i] = p[i] * g.scale;
}

// grid.x = window rows.  level[pixel of the window] = first radius index that contains the point (CROP_LEVELS:
// valid but outside all ten; CROP_INVALID: no depth).  hist[row][0..10] = pixels of the row per level.
__global__ void __launch_bounds__(CROP_THREADS) crop_classify_kernel(CropGeom g, const float *__restrict__ depth,
                                                                     uint8_t *__restrict__ level, int *__restrict__ hist) {
    __shared__ int h[CROP_LEVELS + 1];
    const int wr = blockIdx.x, row = g.r0 + wr, wcols = g.c1 - g.c0 + 1;
    if (threadIdx.x <= CROP_LEVELS) h[threadIdx.x] = 0;
    __syncthreads();
    for (int wc = threadIdx.x; wc < wcols; wc += CROP_THREADS) {
        const int col = g.c0 + wc;
        const float d = __ldg(depth + (size_t)row * g.W + col);
        uint8_t lv = CROP_INVALID;
        if (d > 0.f) {
            double p[3];
            backproject_px(g, row, col, d, p);
            // nocs_data_process.py:93: sqrt(sum((pts - center)^2)), sum over the 3 components in order
            const double dx = p[0] - g.center[0], dy = p[1] - g.center[1], dz = p[2] - g.center[2];
            const double dist = sqrt(dx * dx + dy * dy + dz * dz);
            lv = CROP_LEVELS;
#pragma unroll
            for (int i = CROP_LEVELS - 1; i >= 0; --i)
                if (dist <= g.radii[i]) lv = (uint8_t)i;
            atomicAdd(&h[lv], 1);
        }
        level[(size_t)wr * wcols + wc] = lv;
    }
    __syncthreads();
    if (threadIdx.x <= CROP_LEVELS) hist[wr * (CROP_LEVELS + 1) + threadIdx.x] = h[threadIdx.x];
}

// meta[0] = n (selected points), meta[1] = level threshold (pixels with level <= meta[1] are selected; CROP_LEVELS =
// every valid pixel), meta[2] = index of the radius the loop stopped at, row_off[wr] = exclusive offset of window row wr
__global__ void __launch_bounds__(1024) crop_select_kernel(int nrows, int grow, const int *__restrict__ hist,
                                                           int *__restrict__ row_off, int *__restrict__ meta) {
    __shared__ int tot[CROP_LEVELS + 1];
    __shared__ int thr_s;
    __shared__ int part[1024];
    const int tid = threadIdx.x;
    if (tid <= CROP_LEVELS) {
        int s = 0;
        for (int r = 0; r < nrows; ++r) s += hist[r * (CROP_LEVELS + 1) + tid];
        tot[tid] = s;
    }
    __syncthreads();
    if (tid == 0) {
        // nocs_data_process.py:95-99: stop at the first radius with >= 10 points (only radius 0 is tried when the
        // caller does not resample: num_points is None); after ten tries the last one tested stays
        int cum = 0, stop = CROP_LEVELS - 1;
        for (int i = 0; i < CROP_LEVELS; ++i) {
            cum += tot[i];
            if (cum >= 10 || !grow) { stop = i; break; }
        }
        int n = 0;
        for (int i = 0; i <= stop; ++i) n += tot[i];
        int thr = stop;
        if (n == 0 && grow) {            // :101-102: nothing inside the ball -> every back-projected point
            thr = CROP_LEVELS;
            for (int i = 0; i <= CROP_LEVELS; ++i) n += tot[i];
        }
        thr_s = thr;
        meta[0] = n; meta[1] = thr; meta[2] = stop;
    }
    __syncthreads();
    const int thr = thr_s;
    // exclusive scan of the per-row selected counts (nrows <= a few hundred: chunked serial scan per thread + block scan)
    const int per = (nrows + 1023) / 1024;
    int mine = 0;
    for (int k = 0; k < per; ++k) {
        const int r = tid * per + k;
        if (r < nrows)
            for (int i = 0; i <= thr; ++i) mine += hist[r * (CROP_LEVELS + 1) + i];
    }
    part[tid] = mine;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int v = tid >= o ? part[tid - o] : 0;
        __syncthreads();
        part[tid] += v;
        __syncthreads();
    }
    int run = part[tid] - mine;
    for (int k = 0; k < per; ++k) {
        const int r = tid * per + k;
        if (r < nrows) {
            row_off[r] = run;
            for (int i = 0; i <= thr; ++i) run += hist[r * (CROP_LEVELS + 1) + i];
        }
    }
}

// grid.x = window rows, one warp per row: ordered compaction (np.where order: row-major)
__global__ void __launch_bounds__(32) crop_compact_kernel(CropGeom g, const float *__restrict__ depth, const int *__restrict__ mask,
                                                          const uint8_t *__restrict__ level, const int *__restrict__ row_off,
                                                          const int *__restrict__ meta, double *__restrict__ pts,
                                                          int *__restrict__ pmask, int *__restrict__ pix) {
    const int wr = blockIdx.x, row = g.r0 + wr, wcols = g.c1 - g.c0 + 1, lane = threadIdx.x;
    const int thr = meta[1];
    int base = row_off[wr];
    const unsigned lt = lanemask_lt();
    for (int w0 = 0; w0 < wcols; w0 += 32) {
        const int wc = w0 + lane;
        const bool sel = wc < wcols && level[(size_t)wr * wcols + wc] <= thr;     // CROP_INVALID = 255 is never selected
        const unsigned bal = __ballot_sync(kFull, sel);
        if (sel) {
            const int o = base + __popc(bal & lt), col = g.c0 + wc;
            double p[3];
            backproject_px(g, row, col, __ldg(depth + (size_t)row * g.W + col), p);
            pts[(size_t)o * 3 + 0] = p[0]; pts[(size_t)o * 3 + 1] = p[1]; pts[(size_t)o * 3 + 2] = p[2];
            pmask[o] = mask ? __ldg(mask + (size_t)row * g.W + col) : 0;
            pix[o] = row * g.W + col;
        }
        base += __popc(bal);
    }
}

// position j of the (tiled, optionally permuted) list -> source point: sel ? sel[j] % n : j % n (the doubling concat of
// nocs_data_process.py:105-106 makes the tiled list idx[q] = idx[q % n]); writes the float32 cloud FPS runs on
__global__ void crop_subset_kernel(int n, int count, const int64_t *__restrict__ sel, const double *__restrict__ pts,
                                   float *__restrict__ cloud) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    const int64_t src = (sel ? sel[j] : (int64_t)j) % n;
#pragma unroll
    for (int c = 0; c < 3; ++c) cloud[(size_t)j * 3 + c] = (float)pts[(size_t)src * 3 + c];      // torch.tensor(xyz).float()
}

__global__ void crop_gather_kernel(int n, int count, const int *__restrict__ fps_idx, const int64_t *__restrict__ sel,
                                   const double *__restrict__ pts, const int *__restrict__ pmask, const int *__restrict__ pix,
                                   double *__restrict__ out_pts, int *__restrict__ out_mask, int64_t *__restrict__ out_idx,
                                   int *__restrict__ out_pix) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= count) return;
    const int64_t q = fps_idx ? (int64_t)fps_idx[j] : (int64_t)j;      // position in the subset (no resample: identity)
    const int64_t src = (sel ? sel[q] : q) % n;                          // ... in the tiled list ... in the selected list
#pragma unroll
    for (int c = 0; c < 3; ++c) out_pts[(size_t)j * 3 + c] = pts[(size_t)src * 3 + c];
    out_mask[j] = pmask[src];
    out_idx[j] = src;
    out_pix[j] = pix[src];
}

}  // namespace captra

using namespace captra;

static int fill_geom(CropGeom &g, int h, int w, const int *window, const double *kinv, const double *center, double radius) {
    CAPTRA_REQUIRE(h >= 1 && w >= 1 && window && kinv && center, "crop: bad image size or null host argument");
    g.H = h; g.W = w; g.r0 = window[0]; g.c0 = window[1]; g.r1 = window[2]; g.c1 = window[3];
    CAPTRA_REQUIRE(g.r0 >= 0 && g.c0 >= 0 && g.r1 < h && g.c1 < w, "crop: window outside the image");
    for (int i = 0; i < 9; ++i) g.kinv[i] = kinv[i];
    for (int i = 0; i < 3; ++i) g.center[i] = center[i];
    double r = radius > 0.05 ? radius : 0.05;                 // radius = max(radius, 0.05)
    for (int i = 0; i < CROP_LEVELS; ++i) { g.radii[i] = r; r *= 1.10; }
    g.scale = 0.001;
    return CAPTRA_OK;
}

extern "C" int captra_crop_select(int h, int w, const float *depth, const int *mask, const int *window_host,
                                  const double *kinv_host, const double *center_host, double radius, int grow,
                                  uint8_t *level, int *hist, int *row_off, int *meta, double *pts, int *pmask, int *pix,
                                  captra_stream_t stream) {
    CropGeom g;
    int rc = fill_geom(g, h, w, window_host, kinv_host, center_host, radius);
    if (rc) return rc;
    CAPTRA_REQUIRE(depth && level && hist && row_off && meta && pts && pmask && pix, "crop_select: null pointer");
    const int nrows = g.r1 - g.r0 + 1, ncols = g.c1 - g.c0 + 1;
    cudaStream_t s = as_stream(stream);
    if (nrows <= 0 || ncols <= 0) {
        CAPTRA_CUDA(cudaMemsetAsync(meta, 0, 3 * sizeof(int), s));
        return CAPTRA_OK;
    }
    crop_classify_kernel<<<nrows, CROP_THREADS, 0, s>>>(g, depth, level, hist);
    CAPTRA_CHECK_LAUNCH("crop_classify");
    crop_select_kernel<<<1, 1024, 0, s>>>(nrows, grow, hist, row_off, meta);
    CAPTRA_CHECK_LAUNCH("crop_select");
    crop_compact_kernel<<<nrows, 32, 0, s>>>(g, depth, mask, level, row_off, meta, pts, pmask, pix);
    CAPTRA_CHECK_LAUNCH("crop_compact");
    return CAPTRA_OK;
}

extern "C" int captra_crop_subset(int n, int count, const int64_t *sel, const double *pts, float *cloud, captra_stream_t stream) {
    CAPTRA_REQUIRE(n >= 1 && count >= 0, "crop_subset: bad sizes");
    if (count == 0) return CAPTRA_OK;
    CAPTRA_REQUIRE(pts && cloud, "crop_subset: null pointer");
    crop_subset_kernel<<<ceil_div(count, 256), 256, 0, as_stream(stream)>>>(n, count, sel, pts, cloud);
    CAPTRA_CHECK_LAUNCH("crop_subset");
    return CAPTRA_OK;
}

extern "C" int captra_crop_gather(int n, int count, const int *fps_idx, const int64_t *sel, const double *pts, const int *pmask,
                                  const int *pix, double *out_pts, int *out_mask, int64_t *out_idx, int *out_pix,
                                  captra_stream_t stream) {
    CAPTRA_REQUIRE(n >= 1 && count >= 0, "crop_gather: bad sizes");
    if (count == 0) return CAPTRA_OK;
    CAPTRA_REQUIRE(pts && pmask && pix && out_pts && out_mask && out_idx && out_pix, "crop_gather: null pointer");
    crop_gather_kernel<<<ceil_div(count, 256), 256, 0, as_stream(stream)>>>(n, count, fps_idx, sel, pts, pmask, pix, out_pts,
                                                                            out_mask, out_idx, out_pix);
    CAPTRA_CHECK_LAUNCH("crop_gather");
    return CAPTRA_OK;
}
