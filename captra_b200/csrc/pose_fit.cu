// pose_fit.cu -- per-part NOCS->camera pose fit on the device.
// Reference: pose_utils/procrustes.py:25-56 (3x3), :110-164 (masked fit), :167-228 (2-D fit for
// symmetric categories), pose_utils/pose_fit.py:26-53.  The reference computes a dozen small
// torch ops per frame and runs torch.svd on the CPU after an M.cpu() copy (a device sync in
// the middle of every frame, procrustes.py:27,170).
//
// B200 design: one CTA per (cloud, part).  Two coalesced passes over the part's points give the
// 17 sufficient statistics (count, two centroids, 3x3 centred cross-covariance H, sum|s_c|^2);
// fp32 per-thread partials, fp64 block reduction.  Everything else (optional 3x3 Procrustes,
// optional 2-D symmetric refinement, scale, translation, validity) is closed-form on those 17
// numbers and is solved by one thread in fp64 -- no SVD library, no host round trip.
//
// Closed forms used (U,S,V = SVD of M, columns u_i, v_i):
//   3x3:  U diag(1,1,det(U V^T)) V^T = u1 v1^T + u2 v2^T + (u1 x u2)(v1 x v2)^T
//   2x2:  U diag(1,det(U V^T)) V^T   = u1 v1^T + perp(u1) perp(v1)^T
// (det(U) u3 = u1 x u2 and det(V) v3 = v1 x v2 for any orthogonal U, V), so only the two leading
// singular pairs are needed and the result does not depend on LAPACK's sign conventions.
#include "common.cuh"

#include <math.h>

namespace captra {

constexpr double kEps = 1e-6;  // procrustes.py:5

// --- tiny fp64 linear algebra ---------------------------------------------------------------
__device__ inline void jacobi_eig3(double A[3][3], double V[3][3], double lam[3]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) V[i][j] = (i == j);
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        const double diag = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
        if (off <= 1e-300 || off <= 1e-18 * diag) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (A[p][q] == 0.0) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {  // A <- A J
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {  // A <- J^T A
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk;
                    A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < 3; ++i) lam[i] = A[i][i];
}

__device__ inline void cross3(const double a[3], const double b[3], double c[3]) {
    c[0] = a[1] * b[2] - a[2] * b[1];
    c[1] = a[2] * b[0] - a[0] * b[2];
    c[2] = a[0] * b[1] - a[1] * b[0];
}

__device__ inline double norm3(const double a[3]) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

// any unit vector orthogonal to unit vector a
__device__ inline void any_orthogonal(const double a[3], double o[3]) {
    int k = 0;
    if (fabs(a[1]) < fabs(a[k])) k = 1;
    if (fabs(a[2]) < fabs(a[k])) k = 2;
    double e[3] = {0, 0, 0};
    e[k] = 1.0;
    cross3(a, e, o);
    const double n = norm3(o);
    o[0] /= n; o[1] /= n; o[2] /= n;
}

// R = U diag(1,1,det(UV^T)) V^T for M (row-major 3x3).  procrustes.py:30-54.
__device__ inline void procrustes_rot3(const double M[3][3], double R[3][3]) {
    double A[3][3], V[3][3], lam[3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A[i][j] = M[0][i] * M[0][j] + M[1][i] * M[1][j] + M[2][i] * M[2][j];
    jacobi_eig3(A, V, lam);
    int o0 = 0, o1 = 1, o2 = 2, t;  // sort eigenvalues descending
    if (lam[o0] < lam[o1]) { t = o0; o0 = o1; o1 = t; }
    if (lam[o0] < lam[o2]) { t = o0; o0 = o2; o2 = t; }
    if (lam[o1] < lam[o2]) { t = o1; o1 = o2; o2 = t; }
    double v1[3] = {V[0][o0], V[1][o0], V[2][o0]}, v2[3] = {V[0][o1], V[1][o1], V[2][o1]}, v3[3];
    cross3(v1, v2, v3);
    double u1[3], u2[3], u3[3];
    for (int i = 0; i < 3; ++i) {
        u1[i] = M[i][0] * v1[0] + M[i][1] * v1[1] + M[i][2] * v1[2];
        u2[i] = M[i][0] * v2[0] + M[i][1] * v2[1] + M[i][2] * v2[2];
    }
    const double s1 = norm3(u1);
    if (!(s1 > 0.0)) {  // M == 0 (or NaN): LAPACK returns U = V = I for the zero matrix
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) R[i][j] = (s1 == 0.0) ? (double)(i == j) : NAN;
        return;
    }
    for (int i = 0; i < 3; ++i) u1[i] /= s1;
    const double d12 = u1[0] * u2[0] + u1[1] * u2[1] + u1[2] * u2[2];
    for (int i = 0; i < 3; ++i) u2[i] -= d12 * u1[i];
    const double s2 = norm3(u2);
    if (s2 > 1e-150 * s1 && s2 > 0.0) {
        for (int i = 0; i < 3; ++i) u2[i] /= s2;
    } else {
        any_orthogonal(u1, u2);  // rank-1 M: rotation about u1 is not determined by the data
    }
    cross3(u1, u2, u3);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) R[i][j] = u1[i] * v1[j] + u2[i] * v2[j] + u3[i] * v3[j];
}

// R = U diag(1,det(UV^T)) V^T for M (2x2), then the reference's orthogonality check
// (mean|R^T R - I| < 1e-5, else identity; procrustes.py:183-204).
__device__ inline void procrustes_rot2(const double M[2][2], double R[2][2]) {
    const double a00 = M[0][0] * M[0][0] + M[1][0] * M[1][0];
    const double a01 = M[0][0] * M[0][1] + M[1][0] * M[1][1];
    const double a11 = M[0][1] * M[0][1] + M[1][1] * M[1][1];
    const double phi = 0.5 * atan2(2.0 * a01, a00 - a11);
    const double v[2] = {cos(phi), sin(phi)};
    double u[2] = {M[0][0] * v[0] + M[0][1] * v[1], M[1][0] * v[0] + M[1][1] * v[1]};
    const double s = sqrt(u[0] * u[0] + u[1] * u[1]);
    if (s == 0.0) {
        R[0][0] = R[1][1] = 1.0; R[0][1] = R[1][0] = 0.0;
        return;
    }
    u[0] /= s; u[1] /= s;
    // u v^T + perp(u) perp(v)^T, perp(a) = (-a_y, a_x)
    R[0][0] = u[0] * v[0] + u[1] * v[1];
    R[0][1] = u[0] * v[1] - u[1] * v[0];
    R[1][0] = u[1] * v[0] - u[0] * v[1];
    R[1][1] = u[1] * v[1] + u[0] * v[0];
    // validity in fp32 like the reference (a NaN fails the test and 0*NaN poisons the blend)
    const float r00 = (float)R[0][0], r01 = (float)R[0][1], r10 = (float)R[1][0], r11 = (float)R[1][1];
    const float e = (fabsf(r00 * r00 + r10 * r10 - 1.f) + fabsf(r00 * r01 + r10 * r11) * 2.f +
                     fabsf(r01 * r01 + r11 * r11 - 1.f)) * 0.25f;
    if (isnan(e)) { R[0][0] = R[0][1] = R[1][0] = R[1][1] = NAN; return; }
    if (!(e < 1e-5f)) { R[0][0] = R[1][1] = 1.0; R[0][1] = R[1][0] = 0.0; }
}

__global__ void rot3_kernel(int64_t count, const float *__restrict__ M, float *__restrict__ R) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    double m[3][3], r[3][3];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) m[a][b] = (double)M[i * 9 + a * 3 + b];
    procrustes_rot3(m, r);
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) R[i * 9 + a * 3 + b] = (float)r[a][b];
}

__global__ void rot2_kernel(int64_t count, const float *__restrict__ M, float *__restrict__ R) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    double m[2][2] = {{(double)M[i * 4], (double)M[i * 4 + 1]}, {(double)M[i * 4 + 2], (double)M[i * 4 + 3]}}, r[2][2];
    procrustes_rot2(m, r);
    R[i * 4] = (float)r[0][0]; R[i * 4 + 1] = (float)r[0][1];
    R[i * 4 + 2] = (float)r[1][0]; R[i * 4 + 3] = (float)r[1][1];
}

// --- fused masked fit -----------------------------------------------------------------------
constexpr int PF_THREADS = 256;

template <int NV>
__device__ __forceinline__ void block_reduce_sum(double (&v)[NV], double *smem /* [NV][8] */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(kFull, v[i], o);
    }
    __syncthreads();  // protects smem reuse between calls
    if (lane == 0)
#pragma unroll
        for (int i = 0; i < NV; ++i) smem[i * (PF_THREADS / 32) + warp] = v[i];
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < PF_THREADS / 32; ++w) s += smem[i * (PF_THREADS / 32) + w];
        v[i] = s;
    }
}

struct PFArgs {
    int p, n, sym;
    const int64_t *labels;   // [B,N] or null
    const float *mask;       // [B,P,N] binary float, used when labels == null
    const float *source; int64_t ssb, ssp, ssn, ssc;
    const float *target; int64_t tsb, tsp, tsn, tsc;
    const float *rotation;     // [B,P,3,3] or null
    const float *given_scale;  // [B,P] or null
    float *scale, *translation; uint8_t *valid; float *rot_out;
    // tracker extras (captra_part_fit_track): target = target + target_add[b] on load (the cam + points_mean of
    // networks.py:220-221), and the validity blend with the previous pose (networks.py:230-232)
    const float *target_add;   // [B,3] or null
    const float *prev_scale;   // [B,P] or null
    const float *prev_translation;  // [B,P,3] or null
};

__global__ void __launch_bounds__(PF_THREADS) part_fit_kernel(PFArgs a) {
    __shared__ double red[16 * (PF_THREADS / 32)];
    const int bp = blockIdx.x, b = bp / a.p, part = bp % a.p;
    const float *src = a.source + b * a.ssb + part * a.ssp;
    const float *tgt = a.target + b * a.tsb + part * a.tsp;
    const int64_t *lab = a.labels ? a.labels + (size_t)b * a.n : nullptr;
    const float *msk = a.mask ? a.mask + (size_t)bp * a.n : nullptr;
    const bool add = a.target_add != nullptr;
    const float ta0 = add ? a.target_add[b * 3 + 0] : 0.f, ta1 = add ? a.target_add[b * 3 + 1] : 0.f,
                ta2 = add ? a.target_add[b * 3 + 2] : 0.f;
    // x + 0.f is exact, so one code path serves both entry points
    auto T = [&](int i, int c, float off) { return __fadd_rn(__ldg(tgt + i * a.tsn + c * a.tsc), off); };

    // pass 1: count and centroids (procrustes.py:137-138)
    float c = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f, t0 = 0.f, t1 = 0.f, t2 = 0.f;
    for (int i = threadIdx.x; i < a.n; i += PF_THREADS) {
        const bool in = lab ? (__ldg(lab + i) == (int64_t)part) : (__ldg(msk + i) != 0.f);
        if (in) {
            c += 1.f;
            s0 += __ldg(src + i * a.ssn); s1 += __ldg(src + i * a.ssn + a.ssc); s2 += __ldg(src + i * a.ssn + 2 * a.ssc);
            t0 += T(i, 0, ta0); t1 += T(i, 1, ta1); t2 += T(i, 2, ta2);
        }
    }
    double v1[7] = {c, s0, s1, s2, t0, t1, t2};
    block_reduce_sum<7>(v1, red);
    const double cnt = v1[0], den = fmax(cnt, 1.0);
    // the reference forms the centres in fp32; round them the same way before centring
    const float sc0 = (float)(v1[1] / den), sc1 = (float)(v1[2] / den), sc2 = (float)(v1[3] / den);
    const float tc0 = (float)(v1[4] / den), tc1 = (float)(v1[5] / den), tc2 = (float)(v1[6] / den);

    // pass 2: centred cross-covariance H[i][j] = sum m * t_c,i * s_c,j and sum m |s_c|^2
    float h[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, ss = 0.f;
    for (int i = threadIdx.x; i < a.n; i += PF_THREADS) {
        const bool in = lab ? (__ldg(lab + i) == (int64_t)part) : (__ldg(msk + i) != 0.f);
        if (in) {
            const float x0 = __ldg(src + i * a.ssn) - sc0, x1 = __ldg(src + i * a.ssn + a.ssc) - sc1,
                        x2 = __ldg(src + i * a.ssn + 2 * a.ssc) - sc2;
            const float y0 = T(i, 0, ta0) - tc0, y1 = T(i, 1, ta1) - tc1, y2 = T(i, 2, ta2) - tc2;
            h[0] += y0 * x0; h[1] += y0 * x1; h[2] += y0 * x2;
            h[3] += y1 * x0; h[4] += y1 * x1; h[5] += y1 * x2;
            h[6] += y2 * x0; h[7] += y2 * x1; h[8] += y2 * x2;
            ss += x0 * x0 + x1 * x1 + x2 * x2;
        }
    }
    double v2[10] = {h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], ss};
    block_reduce_sum<10>(v2, red);
    if (threadIdx.x != 0) return;

    double H[3][3] = {{v2[0], v2[1], v2[2]}, {v2[3], v2[4], v2[5]}, {v2[6], v2[7], v2[8]}};
    double R[3][3];
    if (a.rotation) {
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) R[i][j] = (double)a.rotation[(size_t)bp * 9 + i * 3 + j];
    } else {
        // rotate_pts_mask weights both sides by sqrt(w+EPS): a uniform (1+EPS) factor on the
        // masked rows, which does not change the rotation (procrustes.py:110-114)
        procrustes_rot3(H, R);
    }
    if (a.sym) {
        // 2-D fit between source xz and (R^T target) xz: M2 = (R^T H)[{0,2}][{0,2}]
        double M2[2][2], R2[2][2];
        const int ax[2] = {0, 2};
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j)
                M2[i][j] = R[0][ax[i]] * H[0][ax[j]] + R[1][ax[i]] * H[1][ax[j]] + R[2][ax[i]] * H[2][ax[j]];
        procrustes_rot2(M2, R2);
        // R <- R * Ry, Ry = [[xx,0,xz],[0,1,0],[zx,0,zz]] (procrustes.py:69-75,151)
        for (int i = 0; i < 3; ++i) {
            const double r0 = R[i][0], r2 = R[i][2];
            R[i][0] = r0 * R2[0][0] + r2 * R2[1][0];
            R[i][2] = r0 * R2[0][1] + r2 * R2[1][1];
        }
    }
    double scale;
    if (a.given_scale) {
        scale = (double)a.given_scale[bp];
    } else {
        double num = 0.0;
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) num += R[i][j] * H[i][j];
        scale = num / (v2[9] + kEps);  // procrustes.py:117-120
    }
    const double sc[3] = {sc0, sc1, sc2}, tc[3] = {tc0, tc1, tc2};
    const double f = cnt / den;  // 1 if the part has points, 0 otherwise (procrustes.py:123-129)
    double tr[3];
    for (int i = 0; i < 3; ++i)
        tr[i] = f * (tc[i] - scale * (R[i][0] * sc[0] + R[i][1] * sc[1] + R[i][2] * sc[2]));

    const float fs = (float)scale;
    const float ft[3] = {(float)tr[0], (float)tr[1], (float)tr[2]};
    // pose_fit.py:26-35,46 -- the *returned* rotation is the input one, so its finiteness is
    // what filter_model_valid sees when a rotation is given
    float rsum = 0.f;
    if (a.rotation)
        for (int i = 0; i < 9; ++i) rsum += a.rotation[(size_t)bp * 9 + i];
    else
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) rsum += (float)R[i][j];
    const bool ok = (cnt > 3.0) && isfinite(fs) && isfinite(ft[0] + ft[1] + ft[2]) && isfinite(rsum);
    a.valid[bp] = ok ? 1 : 0;
    if (a.prev_scale) {
        // networks.py:230-232: valid * new + (1 - valid) * previous, in fp32 exactly as written there (a NaN fit
        // stays NaN: 0 * NaN)
        const float v = ok ? 1.f : 0.f;
        a.scale[bp] = v * fs + (1.f - v) * a.prev_scale[bp];
        for (int i = 0; i < 3; ++i) a.translation[(size_t)bp * 3 + i] = v * ft[i] + (1.f - v) * a.prev_translation[(size_t)bp * 3 + i];
    } else {
        a.scale[bp] = fs;
        for (int i = 0; i < 3; ++i) a.translation[(size_t)bp * 3 + i] = ft[i];
    }
    if (a.rot_out)
        for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 3; ++j) a.rot_out[(size_t)bp * 9 + i * 3 + j] = (float)R[i][j];
}

}  // namespace captra

using namespace captra;

extern "C" int captra_procrustes_rot3(int64_t count, const float *M, float *R, captra_stream_t stream) {
    CAPTRA_REQUIRE(count >= 0, "procrustes_rot3: negative count");
    if (count == 0) return CAPTRA_OK;
    CAPTRA_REQUIRE(M && R, "procrustes_rot3: null pointer");
    rot3_kernel<<<(unsigned)ceil_div<int64_t>(count, 64), 64, 0, as_stream(stream)>>>(count, M, R);
    CAPTRA_CHECK_LAUNCH("procrustes_rot3");
    return CAPTRA_OK;
}

extern "C" int captra_procrustes_rot2(int64_t count, const float *M, float *R, captra_stream_t stream) {
    CAPTRA_REQUIRE(count >= 0, "procrustes_rot2: negative count");
    if (count == 0) return CAPTRA_OK;
    CAPTRA_REQUIRE(M && R, "procrustes_rot2: null pointer");
    rot2_kernel<<<(unsigned)ceil_div<int64_t>(count, 64), 64, 0, as_stream(stream)>>>(count, M, R);
    CAPTRA_CHECK_LAUNCH("procrustes_rot2");
    return CAPTRA_OK;
}

extern "C" int captra_part_fit_st(int b, int p, int n, const int64_t *labels, const float *mask,
                                  const float *source, int64_t ssb, int64_t ssp, int64_t ssn,
                                  int64_t ssc, const float *target, int64_t tsb, int64_t tsp,
                                  int64_t tsn, int64_t tsc, const float *rotation,
                                  const float *given_scale, int sym, float *scale,
                                  float *translation, uint8_t *valid, float *rot_out,
                                  captra_stream_t stream) {
    CAPTRA_REQUIRE(b >= 0 && p >= 0 && n >= 0, "part_fit_st: negative size");
    if (b == 0 || p == 0) return CAPTRA_OK;
    CAPTRA_REQUIRE((labels != nullptr) != (mask != nullptr), "part_fit_st: exactly one of labels / mask");
    CAPTRA_REQUIRE(source && target && scale && translation && valid, "part_fit_st: null pointer");
    PFArgs a;
    a.p = p; a.n = n; a.sym = sym; a.labels = labels; a.mask = mask;
    a.source = source; a.ssb = ssb; a.ssp = ssp; a.ssn = ssn; a.ssc = ssc;
    a.target = target; a.tsb = tsb; a.tsp = tsp; a.tsn = tsn; a.tsc = tsc;
    a.rotation = rotation; a.given_scale = given_scale;
    a.scale = scale; a.translation = translation; a.valid = valid; a.rot_out = rot_out;
    a.target_add = nullptr; a.prev_scale = nullptr; a.prev_translation = nullptr;
    part_fit_kernel<<<b * p, PF_THREADS, 0, as_stream(stream)>>>(a);
    CAPTRA_CHECK_LAUNCH("part_fit_st");
    return CAPTRA_OK;
}

extern "C" int captra_part_fit_track(int b, int p, int n, const int64_t *labels, const float *nocs, const float *points,
                                     const float *points_mean, const float *rotation, int sym, const float *prev_scale,
                                     const float *prev_translation, float *scale, float *translation, uint8_t *valid,
                                     captra_stream_t stream) {
    CAPTRA_REQUIRE(b >= 0 && p >= 0 && n >= 0, "part_fit_track: negative size");
    if (b == 0 || p == 0) return CAPTRA_OK;
    CAPTRA_REQUIRE(labels && nocs && points && points_mean && rotation && prev_scale && prev_translation && scale &&
                       translation && valid, "part_fit_track: null pointer");
    PFArgs a;
    a.p = p; a.n = n; a.sym = sym; a.labels = labels; a.mask = nullptr;
    // source = pred_npcs [B,P,3,N] read as [B,P,N,3] (networks.py:227 transposes the view); target = the cloud
    // [B,3,N] + points_mean, shared by all parts (the `repeat` of networks.py:221 is a zero part stride)
    a.source = nocs; a.ssb = (int64_t)p * 3 * n; a.ssp = (int64_t)3 * n; a.ssn = 1; a.ssc = n;
    a.target = points; a.tsb = (int64_t)3 * n; a.tsp = 0; a.tsn = 1; a.tsc = n;
    a.rotation = rotation; a.given_scale = nullptr;
    a.scale = scale; a.translation = translation; a.valid = valid; a.rot_out = nullptr;
    a.target_add = points_mean; a.prev_scale = prev_scale; a.prev_translation = prev_translation;
    part_fit_kernel<<<b * p, PF_THREADS, 0, as_stream(stream)>>>(a);
    CAPTRA_CHECK_LAUNCH("part_fit_track");
    return CAPTRA_OK;
}
