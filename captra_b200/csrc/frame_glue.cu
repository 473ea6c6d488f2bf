// frame_glue.cu -- the per-frame elementwise / small-reduction steps between the networks and the pose fit,
// each as ONE launch (the reference spends ~150 small torch kernels per frame on them):
//   captra_canonicalize     networks.py:38-41 / :184-187   (cam + mean - t) -> R^T . -> / s, written point-major
//   captra_coord_head_post  networks.py:44-46, model.py:458 softmax -> argmax labels, sigmoid - 0.5
//   captra_rot_head_post    blocks.py:181-193, networks.py:127-141, pose_utils/rotations.py:300-387,
//                           part_dof_utils.py:124-141      per-point 6-D / 3-D -> matrix, masked mean, default,
//                                                          Gram-Schmidt (or y-axis frame), R_prev . dR
// All arithmetic is fp32 in the reference's operation order where that order is defined by its Python
// (sums over points are the exception: fp32 partials + fp64 block reduction here, torch's pairwise sum there).
#include "common.cuh"

#include <math.h>

namespace captra {

// ---- canonicalise -----------------------------------------------------------------------------
// points [B,3,N] (mean-subtracted), mean [B,3], per (cloud, part) pose: rotation [B*P,3,3], translation [B*P,3],
// scale [B*P].  out row (b*P+p)*N + i = R^T ((x + mean) - t) / s.
__global__ void __launch_bounds__(256) canonicalize_kernel(int p, int n, const float *__restrict__ points,
                                                           const float *__restrict__ mean, const float *__restrict__ rot,
                                                           const float *__restrict__ trans, const float *__restrict__ scale,
                                                           float *__restrict__ out_pm, float *__restrict__ out_cm,
                                                           float *__restrict__ out_dup) {
    const int bp = blockIdx.y, b = bp / p;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float *pt = points + (size_t)b * 3 * n;
    const float *R = rot + (size_t)bp * 9;
    // networks.py:38-39: cam = cam + points_mean; cam = cam - translation (two rounded fp32 ops)
    const float x = __fsub_rn(__fadd_rn(__ldg(pt + i), __ldg(mean + b * 3 + 0)), __ldg(trans + bp * 3 + 0));
    const float y = __fsub_rn(__fadd_rn(__ldg(pt + n + i), __ldg(mean + b * 3 + 1)), __ldg(trans + bp * 3 + 1));
    const float z = __fsub_rn(__fadd_rn(__ldg(pt + 2 * n + i), __ldg(mean + b * 3 + 2)), __ldg(trans + bp * 3 + 2));
    const float s = __ldg(scale + bp);
    float o[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        // :40 matmul(R^T, cam): row c of R^T = column c of R; k ascending, fused multiply-adds
        float acc = __fmul_rn(__ldg(R + c), x);
        acc = __fmaf_rn(__ldg(R + 3 + c), y, acc);
        acc = __fmaf_rn(__ldg(R + 6 + c), z, acc);
        o[c] = __fdiv_rn(acc, s);                                   // :41
    }
    const size_t row = (size_t)bp * n + i;
    if (out_pm) { out_pm[row * 3 + 0] = o[0]; out_pm[row * 3 + 1] = o[1]; out_pm[row * 3 + 2] = o[2]; }
    if (out_dup) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { out_dup[row * 6 + c] = o[c]; out_dup[row * 6 + 3 + c] = o[c]; }
    }
    if (out_cm) {
#pragma unroll
        for (int c = 0; c < 3; ++c) out_cm[((size_t)bp * 3 + c) * n + i] = o[c];
    }
}

// ---- CoordNet head post-processing ---------------------------------------------------------------
constexpr int CH_MAX_SEG = 8;
__global__ void __launch_bounds__(256) coord_head_post_kernel(int n, int nseg, int nnocs, const float *__restrict__ seg_raw,
                                                              int64_t ld_seg, const float *__restrict__ nocs_raw, int64_t ld_nocs,
                                                              int64_t *__restrict__ labels, float *__restrict__ nocs,
                                                              float *__restrict__ seg) {
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const size_t row = (size_t)b * n + i;
    // F.softmax(seg, dim=1) (networks.py:45): exp(x - max) / sum, then torch.max(seg, dim=-2)[1] (model.py:458):
    // first index among equal maxima of the PROBABILITIES
    float v[CH_MAX_SEG];
    float mx = -INFINITY;
    for (int c = 0; c < nseg; ++c) { v[c] = __ldg(seg_raw + row * ld_seg + c); mx = fmaxf(mx, v[c]); }
    float sum = 0.f;
    for (int c = 0; c < nseg; ++c) { v[c] = expf(v[c] - mx); sum += v[c]; }
    int best = 0;
    float bestp = -1.f;
    bool any_nan = false;
    for (int c = 0; c < nseg; ++c) {
        const float pr = v[c] / sum;
        if (seg) seg[((size_t)b * nseg + c) * n + i] = pr;
        if (pr != pr && !any_nan) { best = c; any_nan = true; }     // torch.max propagates NaN: the first NaN wins
        if (!any_nan && pr > bestp) { bestp = pr; best = c; }
    }
    labels[row] = best;
    // nocs_head ends in Sigmoid (blocks.py:118-135), then - 0.5 (networks.py:46)
    for (int c = 0; c < nnocs; ++c) {
        const float x = __ldg(nocs_raw + row * ld_nocs + c);
        const float sg = 1.f / (1.f + expf(-x));
        nocs[((size_t)b * nnocs + c) * n + i] = sg - 0.5f;
    }
}

// ---- RotationRegressor head post-processing ---------------------------------------------------------
__device__ __forceinline__ void normalize3(const float v[3], float o[3]) {
    // rotations.py:300-312: v / max(|v|, 1e-8) * valid + (1,0,0) * (1 - valid), valid = |v| > 1e-8
    const float mag = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    const float valid = mag > 1e-8f ? 1.f : 0.f;
    const float den = fmaxf(mag, 1e-8f);
    o[0] = (v[0] / den) * valid + 1.f * (1.f - valid);
    o[1] = (v[1] / den) * valid + 0.f * (1.f - valid);
    o[2] = (v[2] / den) * valid + 0.f * (1.f - valid);
}

__device__ __forceinline__ void cross3f(const float a[3], const float b[3], float c[3]) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

constexpr int RH_THREADS = 256;

struct RotHeadArgs {
    int p, n, sym;
    const float *raw[8];          // per part: head p's output on copy p, point-major [B*N, D] (D = 3 sym / 6)
    int64_t ld;
    const int64_t *labels;        // [B,N]
    const float *rot_prev;        // [B,P,3,3]
    float *rotation;              // [B,P,3,3] = rot_prev . dR
    float *rtvec;                 // [B,P,D'] (D' = 3 sym / 9) or null: the masked mean (+ default)
};

__global__ void __launch_bounds__(RH_THREADS) rot_head_post_kernel(RotHeadArgs a) {
    __shared__ double red[10][RH_THREADS / 32];
    const int b = blockIdx.x, part = blockIdx.y, bp = b * a.p + part;
    const int D = a.sym ? 3 : 6, DO = a.sym ? 3 : 9;
    const float *raw = a.raw[part] + (size_t)b * a.n * a.ld;
    const int64_t *lab = a.labels + (size_t)b * a.n;
    float acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, cnt = 0.f;
    for (int i = threadIdx.x; i < a.n; i += RH_THREADS) {
        if (__ldg(lab + i) != (int64_t)part) continue;
        cnt += 1.f;
        float v[6];
        for (int c = 0; c < D; ++c) v[c] = __ldg(raw + (size_t)i * a.ld + c);
        if (a.sym) {                      // blocks.py:189-192: normalize_vector of the 3-vector
            float o[3];
            normalize3(v, o);
            acc[0] += o[0]; acc[1] += o[1]; acc[2] += o[2];
        } else {                          // blocks.py:184-188 -> rotations.py:330-343, columns (x, y, z) stacked on dim 2
            float x[3], zr[3], z[3], y[3];
            normalize3(v, x);
            cross3f(x, v + 3, zr);
            normalize3(zr, z);
            cross3f(z, x, y);
#pragma unroll
            for (int r = 0; r < 3; ++r) { acc[r * 3 + 0] += x[r]; acc[r * 3 + 1] += y[r]; acc[r * 3 + 2] += z[r]; }
        }
    }
    // block reduction in fp64
    double v[10];
#pragma unroll
    for (int k = 0; k < 9; ++k) v[k] = acc[k];
    v[9] = cnt;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 10; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(kFull, v[k], o);
        if (lane == 0) red[k][warp] = v[k];
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    double tot[10];
    for (int k = 0; k < 10; ++k) {
        double s = 0.0;
        for (int w = 0; w < RH_THREADS / 32; ++w) s += red[k][w];
        tot[k] = s;
    }
    // networks.py:129-139: masked mean, default where the part has no point
    const float count = (float)tot[9];
    float w[9];
    const float valid = count > 0.f ? 1.f : 0.f;
    for (int k = 0; k < DO; ++k) {
        const float mean = (float)tot[k] / fmaxf(count, 1.f);
        const float dflt = a.sym ? (k == 1 ? 1.f : 0.f) : ((k == 0 || k == 4 || k == 8) ? 1.f : 0.f);
        w[k] = valid * mean + (1.f - valid) * dflt;
        if (a.rtvec) a.rtvec[(size_t)bp * DO + k] = w[k];
    }
    float dR[3][3];
    if (a.sym) {
        // rotations.py:375-387: y = v/|v|, z = normalize((1,0,0) x y), x = y x z; columns (x, y, z)
        float y[3], zr[3], z[3], x[3];
        const float e0[3] = {1.f, 0.f, 0.f};
        normalize3(w, y);
        cross3f(e0, y, zr);
        normalize3(zr, z);
        cross3f(y, z, x);
        for (int r = 0; r < 3; ++r) { dR[r][0] = x[r]; dR[r][1] = y[r]; dR[r][2] = z[r]; }
    } else {
        // rotations.py:354-372: Gram-Schmidt on the columns of the mean matrix
        float a1[3], a2[3], a3[3];
        for (int r = 0; r < 3; ++r) { a1[r] = w[r * 3 + 0]; a2[r] = w[r * 3 + 1]; a3[r] = w[r * 3 + 2]; }
        auto proj = [](const float u[3], const float q[3], float o[3]) {      // rotations.py:344-351
            const float top = u[0] * q[0] + u[1] * q[1] + u[2] * q[2];
            const float bottom = fmaxf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2], 1e-8f);
            const float f = top / bottom;
            o[0] = f * u[0]; o[1] = f * u[1]; o[2] = f * u[2];
        };
        float p12[3], p13[3], p23[3], u2[3], u3[3];
        proj(a1, a2, p12);
        for (int r = 0; r < 3; ++r) u2[r] = a2[r] - p12[r];
        proj(a1, a3, p13);
        proj(u2, a3, p23);
        for (int r = 0; r < 3; ++r) u3[r] = a3[r] - p13[r] - p23[r];
        float n1[3], n2[3], n3[3];
        normalize3(a1, n1); normalize3(u2, n2); normalize3(u3, n3);
        for (int r = 0; r < 3; ++r) { dR[r][0] = n1[r]; dR[r][1] = n2[r]; dR[r][2] = n3[r]; }
    }
    // part_dof_utils.py:124-128: rotation = R_prev . dR
    const float *Rp = a.rot_prev + (size_t)bp * 9;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            float s = Rp[i * 3 + 0] * dR[0][j];
            s = fmaf(Rp[i * 3 + 1], dR[1][j], s);
            s = fmaf(Rp[i * 3 + 2], dR[2][j], s);
            a.rotation[(size_t)bp * 9 + i * 3 + j] = s;
        }
}

// ---- GroupNorm + ReLU applied in place (cloud sizes that are not a multiple of the 128-row tile) ----------------
__global__ void __launch_bounds__(256) gn_relu_rows_kernel(int64_t rows, int c, int rows_per_cloud, float *__restrict__ y, int64_t ld,
                                                           const float *__restrict__ scale, const float *__restrict__ shift) {
    const int64_t total = rows * c;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / c;
        const int ch = (int)(e - r * c);
        const int64_t cloud = r / rows_per_cloud;
        float *p = y + r * ld + ch;
        *p = fmaxf(fmaf(*p, __ldg(scale + cloud * c + ch), __ldg(shift + cloud * c + ch)), 0.f);
    }
}

}  // namespace captra

using namespace captra;

extern "C" int captra_group_norm_relu_rows(int64_t rows, int c, int rows_per_cloud, float *y, int64_t ldy, const float *scale,
                                           const float *shift, captra_stream_t stream) {
    CAPTRA_REQUIRE(rows >= 0 && c >= 1 && rows_per_cloud >= 1 && ldy >= c, "group_norm_relu_rows: bad sizes");
    if (rows == 0) return CAPTRA_OK;
    CAPTRA_REQUIRE(y && scale && shift, "group_norm_relu_rows: null pointer");
    const int64_t blocks = ceil_div<int64_t>(rows * c, 256);
    gn_relu_rows_kernel<<<(unsigned)(blocks < 148 * 16 ? blocks : 148 * 16), 256, 0, as_stream(stream)>>>(rows, c, rows_per_cloud, y, ldy, scale, shift);
    CAPTRA_CHECK_LAUNCH("group_norm_relu_rows");
    return CAPTRA_OK;
}

extern "C" int captra_canonicalize(int b, int p, int n, const float *points, const float *points_mean,
                                   const float *rotation, const float *translation, const float *scale,
                                   float *out_pm, float *out_cm, float *out_dup, captra_stream_t stream) {
    CAPTRA_REQUIRE(b >= 0 && p >= 1 && n >= 0, "canonicalize: bad sizes");
    if (b == 0 || n == 0) return CAPTRA_OK;
    CAPTRA_REQUIRE((int64_t)b * p <= 65535, "canonicalize: too many clouds");
    CAPTRA_REQUIRE(points && points_mean && rotation && translation && scale, "canonicalize: null pointer");
    CAPTRA_REQUIRE(out_pm || out_cm || out_dup, "canonicalize: no output requested");
    canonicalize_kernel<<<dim3(ceil_div(n, 256), b * p), 256, 0, as_stream(stream)>>>(p, n, points, points_mean, rotation,
                                                                                      translation, scale, out_pm, out_cm, out_dup);
    CAPTRA_CHECK_LAUNCH("canonicalize");
    return CAPTRA_OK;
}

extern "C" int captra_coord_head_post(int b, int n, int nseg, int nnocs, const float *seg_raw, int64_t ld_seg,
                                      const float *nocs_raw, int64_t ld_nocs, int64_t *labels, float *nocs, float *seg,
                                      captra_stream_t stream) {
    CAPTRA_REQUIRE(b >= 0 && n >= 0 && nseg >= 1 && nseg <= CH_MAX_SEG && nnocs >= 0, "coord_head_post: bad sizes (nseg <= %d)", CH_MAX_SEG);
    if (b == 0 || n == 0) return CAPTRA_OK;
    CAPTRA_REQUIRE(b <= 65535, "coord_head_post: too many clouds");
    CAPTRA_REQUIRE(seg_raw && labels && (nnocs == 0 || (nocs_raw && nocs)), "coord_head_post: null pointer");
    coord_head_post_kernel<<<dim3(ceil_div(n, 256), b), 256, 0, as_stream(stream)>>>(n, nseg, nnocs, seg_raw, ld_seg, nocs_raw,
                                                                                      ld_nocs, labels, nocs, seg);
    CAPTRA_CHECK_LAUNCH("coord_head_post");
    return CAPTRA_OK;
}

extern "C" int captra_rot_head_post(int b, int p, int n, int sym, const float *const *raw_host_ptrs, int64_t ld,
                                    const int64_t *labels, const float *rot_prev, float *rotation, float *rtvec,
                                    captra_stream_t stream) {
    CAPTRA_REQUIRE(b >= 0 && p >= 1 && p <= 8 && n >= 0, "rot_head_post: bad sizes (at most 8 parts)");
    if (b == 0) return CAPTRA_OK;
    CAPTRA_REQUIRE(raw_host_ptrs && labels && rot_prev && rotation, "rot_head_post: null pointer");
    CAPTRA_REQUIRE(ld >= (sym ? 3 : 6), "rot_head_post: row stride %lld below the head width", (long long)ld);
    RotHeadArgs a;
    a.p = p; a.n = n; a.sym = sym; a.ld = ld; a.labels = labels; a.rot_prev = rot_prev; a.rotation = rotation; a.rtvec = rtvec;
    for (int i = 0; i < 8; ++i) a.raw[i] = i < p ? raw_host_ptrs[i] : nullptr;
    for (int i = 0; i < p; ++i) CAPTRA_REQUIRE(a.raw[i], "rot_head_post: null head output for part %d", i);
    rot_head_post_kernel<<<dim3(b, p), RH_THREADS, 0, as_stream(stream)>>>(a);
    CAPTRA_CHECK_LAUNCH("rot_head_post");
    return CAPTRA_OK;
}
