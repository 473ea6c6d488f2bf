/*
 * oracle/cpu_ref.c -- CPU restatement of CAPTRA's pointnet_lib CUDA kernels.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (captra_b200/) may import, link or
 * call this file; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs use it, as the checker or as the timed CPU baseline.
 *
 * Every function restates the *CUDA-extension* semantics of the reference (the thing the
 * tracker actually runs on a GPU), NOT the torch CPU fallback in pointnet_utils.py, which is
 * numerically different.  File:line citations are into
 * /root/reference/network/models/pointnet_lib/src/.
 *
 * Arithmetic contract: nvcc contracts
 *     dx*dx + dy*dy + dz*dz   ->   fma(dz,dz, fma(dx,dx, dy*dy))
 * so all distances here are computed with explicit fmaf in that order and the file must be
 * compiled with -ffp-contract=off so gcc adds no contraction of its own.
 *
 * Parity pin: the reference ships no golden vectors.  This oracle is
 * pinned against the reference's own kernels compiled from /root/reference into
 * oracle/_ref/libpointnet2_ref.so (see oracle/Makefile) and run on the GPU box
 * (the `refcu` comparisons of tests/test_ops_gpu.py: test_fps_bit_exact, test_ball_query_bit_exact, test_group_gather_bit_exact, test_three_nn_interpolate_bit_exact, test_knn_bit_exact, test_backward_ops), and against fixtures
 * generated here by importing the Python reference (tests/golden/).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* (c - p) squared distance in the reference's contracted order.
 * ball_query_gpu.cu:33, interpolate_gpu.cu:108 (d = centre - point),
 * sampling_gpu.cu:133 (d = point - old); the sign does not matter for the squares but we
 * keep the operand order anyway. */
static inline float sqdist(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = ax - bx, dy = ay - by, dz = az - bz;
    float t = dy * dy;
    t = fmaf(dx, dx, t);
    return fmaf(dz, dz, t);
}

int ref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void ref_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ball_query_gpu.cu:9-45.  idx rows with no hit are left untouched (caller zeroes them,
 * pointnet2_utils.py:261). */
void ref_ball_query(int b, int n, int m, float radius, int nsample,
                    const float *new_xyz, const float *xyz, int *idx) {
    const float radius2 = radius * radius;
#pragma omp parallel for collapse(2) schedule(static)
    for (int bs = 0; bs < b; ++bs) {
        for (int pt = 0; pt < m; ++pt) {
            const float *c = new_xyz + ((size_t)bs * m + pt) * 3;
            const float *p = xyz + (size_t)bs * n * 3;
            int *o = idx + ((size_t)bs * m + pt) * nsample;
            const float cx = c[0], cy = c[1], cz = c[2];
            int cnt = 0;
            for (int k = 0; k < n; ++k) {
                float d2 = sqdist(cx, cy, cz, p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2]);
                if (d2 < radius2) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; ++l) o[l] = k;
                    o[cnt] = k;
                    if (++cnt >= nsample) break;
                }
            }
        }
    }
}

/* group_points_gpu.cu:47-66.  64-bit offsets (the reference's int offsets overflow at
 * B*C*M*K >= 2^31). */
void ref_group_points(int b, int c, int n, int npoints, int nsample,
                      const float *points, const int *idx, float *out) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int bs = 0; bs < b; ++bs)
        for (int ci = 0; ci < c; ++ci) {
            const float *src = points + ((size_t)bs * c + ci) * n;
            float *dst = out + ((size_t)bs * c + ci) * npoints * nsample;
            const int *ix = idx + (size_t)bs * npoints * nsample;
            for (size_t j = 0; j < (size_t)npoints * nsample; ++j) dst[j] = src[ix[j]];
        }
}

/* group_points_gpu.cu:8-25 (atomicAdd scatter; sequential order here). */
void ref_group_points_grad(int b, int c, int n, int npoints, int nsample,
                           const float *grad_out, const int *idx, float *grad_points) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int bs = 0; bs < b; ++bs)
        for (int ci = 0; ci < c; ++ci) {
            float *dst = grad_points + ((size_t)bs * c + ci) * n;
            const float *src = grad_out + ((size_t)bs * c + ci) * npoints * nsample;
            const int *ix = idx + (size_t)bs * npoints * nsample;
            for (size_t j = 0; j < (size_t)npoints * nsample; ++j) dst[ix[j]] += src[j];
        }
}

/* sampling_gpu.cu:8-24 */
void ref_gather_points(int b, int c, int n, int npoints,
                       const float *points, const int *idx, float *out) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int bs = 0; bs < b; ++bs)
        for (int ci = 0; ci < c; ++ci) {
            const float *src = points + ((size_t)bs * c + ci) * n;
            float *dst = out + ((size_t)bs * c + ci) * npoints;
            const int *ix = idx + (size_t)bs * npoints;
            for (int j = 0; j < npoints; ++j) dst[j] = src[ix[j]];
        }
}

/* sampling_gpu.cu:46-63 */
void ref_gather_points_grad(int b, int c, int n, int npoints,
                            const float *grad_out, const int *idx, float *grad_points) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int bs = 0; bs < b; ++bs)
        for (int ci = 0; ci < c; ++ci) {
            float *dst = grad_points + ((size_t)bs * c + ci) * n;
            const float *src = grad_out + ((size_t)bs * c + ci) * npoints;
            const int *ix = idx + (size_t)bs * npoints;
            for (int j = 0; j < npoints; ++j) dst[ix[j]] += src[j];
        }
}

/* cuda_utils.h:10-14 -- largest power of two <= work_size, capped at 1024. */
int ref_opt_n_threads(int work_size) {
    const int pow_2 = (int)(log((double)work_size) / log(2.0));
    int t = 1 << pow_2;
    if (t > 1024) t = 1024;
    if (t < 1) t = 1;
    return t;
}

/* sampling_gpu.cu:86-209: thread-level emulation of the kernel, one "CTA" per cloud.
 *   - per-thread strided scan, strict '>' against best=-1 / besti=0       (:119-138)
 *   - shared-memory tournament: slot tid vs tid+stride for stride=block/2..1,
 *     value = max(v1,v2), index = v2 > v1 ? i2 : i1                         (:86-91,:143-203)
 * temp is read and updated in place exactly like the kernel (caller pre-fills 1e10,
 * pointnet2_utils.py:27).  CUDA min/max on floats are fminf/fmaxf (NaN-dropping). */
void ref_furthest_point_sampling(int b, int n, int m, const float *dataset, float *temp,
                                 int *idxs) {
    if (m <= 0) return;
    const int block = ref_opt_n_threads(n);
#pragma omp parallel for schedule(dynamic, 1)
    for (int bs = 0; bs < b; ++bs) {
        const float *p = dataset + (size_t)bs * n * 3;
        float *tmp = temp + (size_t)bs * n;
        int *out = idxs + (size_t)bs * m;
        float *dists = (float *)malloc(sizeof(float) * block);
        int *dists_i = (int *)malloc(sizeof(int) * block);
        int old = 0;
        out[0] = old;
        for (int j = 1; j < m; ++j) {
            const float x1 = p[old * 3 + 0], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
            for (int tid = 0; tid < block; ++tid) {
                int besti = 0;
                float best = -1.0f;
                for (int k = tid; k < n; k += block) {
                    float d = sqdist(p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2], x1, y1, z1);
                    float d2 = fminf(d, tmp[k]);
                    tmp[k] = d2;
                    besti = d2 > best ? k : besti;
                    best = d2 > best ? d2 : best;
                }
                dists[tid] = best;
                dists_i[tid] = besti;
            }
            for (int stride = block / 2; stride >= 1; stride >>= 1)
                for (int tid = 0; tid < stride; ++tid) {
                    const float v1 = dists[tid], v2 = dists[tid + stride];
                    const int i1 = dists_i[tid], i2 = dists_i[tid + stride];
                    dists[tid] = fmaxf(v1, v2);
                    dists_i[tid] = v2 > v1 ? i2 : i1;
                }
            old = dists_i[0];
            out[j] = old;
        }
        free(dists);
        free(dists_i);
    }
}

/* interpolate_gpu.cu:81-124.  best* are double initialised to 1e40, d is float. */
void ref_three_nn(int b, int n, int m, const float *unknown, const float *known,
                  float *dist2, int *idx) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int bs = 0; bs < b; ++bs)
        for (int pt = 0; pt < n; ++pt) {
            const float *u = unknown + ((size_t)bs * n + pt) * 3;
            const float *kn = known + (size_t)bs * m * 3;
            const float ux = u[0], uy = u[1], uz = u[2];
            double best1 = 1e40, best2 = 1e40, best3 = 1e40;
            int besti1 = 0, besti2 = 0, besti3 = 0;
            for (int k = 0; k < m; ++k) {
                float d = sqdist(ux, uy, uz, kn[k * 3 + 0], kn[k * 3 + 1], kn[k * 3 + 2]);
                if (d < best1) {
                    best3 = best2; besti3 = besti2;
                    best2 = best1; besti2 = besti1;
                    best1 = d; besti1 = k;
                } else if (d < best2) {
                    best3 = best2; besti3 = besti2;
                    best2 = d; besti2 = k;
                } else if (d < best3) {
                    best3 = d; besti3 = k;
                }
            }
            float *od = dist2 + ((size_t)bs * n + pt) * 3;
            int *oi = idx + ((size_t)bs * n + pt) * 3;
            od[0] = (float)best1; od[1] = (float)best2; od[2] = (float)best3;
            oi[0] = besti1; oi[1] = besti2; oi[2] = besti3;
        }
}

/* interpolate_gpu.cu:9-57: insertion into a sorted list of k (<=200) doubles. */
void ref_knn(int b, int n, int m, int k, const float *unknown, const float *known,
             float *dist2, int *idx) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int bs = 0; bs < b; ++bs)
        for (int pt = 0; pt < n; ++pt) {
            const float *u = unknown + ((size_t)bs * n + pt) * 3;
            const float *kn = known + (size_t)bs * m * 3;
            const float ux = u[0], uy = u[1], uz = u[2];
            double best[200];
            int besti[200];
            for (int i = 0; i < k; ++i) { best[i] = 1e40; besti[i] = 0; }
            for (int i = 0; i < m; ++i) {
                float d = sqdist(ux, uy, uz, kn[i * 3 + 0], kn[i * 3 + 1], kn[i * 3 + 2]);
                for (int j = 0; j < k; ++j) {
                    if (d < best[j]) {
                        for (int l = k - 1; l > j; --l) { best[l] = best[l - 1]; besti[l] = besti[l - 1]; }
                        best[j] = d; besti[j] = i;
                        break;
                    }
                }
            }
            float *od = dist2 + ((size_t)bs * n + pt) * k;
            int *oi = idx + ((size_t)bs * n + pt) * k;
            for (int i = 0; i < k; ++i) { oi[i] = besti[i]; od[i] = (float)best[i]; }
        }
}

/* interpolate_gpu.cu:149-169; contraction fma(w2,p2, fma(w0,p0, w1*p1)) */
void ref_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx,
                           const float *weight, float *out) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int bs = 0; bs < b; ++bs)
        for (int ci = 0; ci < c; ++ci) {
            const float *src = points + ((size_t)bs * c + ci) * m;
            float *dst = out + ((size_t)bs * c + ci) * n;
            const int *ix = idx + (size_t)bs * n * 3;
            const float *w = weight + (size_t)bs * n * 3;
            for (int pt = 0; pt < n; ++pt) {
                float t = w[pt * 3 + 1] * src[ix[pt * 3 + 1]];
                t = fmaf(w[pt * 3 + 0], src[ix[pt * 3 + 0]], t);
                dst[pt] = fmaf(w[pt * 3 + 2], src[ix[pt * 3 + 2]], t);
            }
        }
}

/* interpolate_gpu.cu:192-214 (atomicAdd scatter; sequential order here). */
void ref_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out,
                                const int *idx, const float *weight, float *grad_points) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int bs = 0; bs < b; ++bs)
        for (int ci = 0; ci < c; ++ci) {
            float *dst = grad_points + ((size_t)bs * c + ci) * m;
            const float *g = grad_out + ((size_t)bs * c + ci) * n;
            const int *ix = idx + (size_t)bs * n * 3;
            const float *w = weight + (size_t)bs * n * 3;
            for (int pt = 0; pt < n; ++pt)
                for (int j = 0; j < 3; ++j) dst[ix[pt * 3 + j]] += g[pt] * w[pt * 3 + j];
        }
}

/* Exhaustive check that the reference's floating-point log2 (cuda_utils.h:10-14) equals the
 * integer floor(log2 n) for every n in [1, max_n]; returns the first n that differs, or 0. */
int ref_check_opt_n_threads(int max_n) {
    for (int n = 1; n <= max_n; ++n) {
        int bits = 0;
        while ((2 << bits) <= n) ++bits;
        int want = 1 << bits;
        if (want > 1024) want = 1024;
        if (ref_opt_n_threads(n) != want) return n;
    }
    return 0;
}
