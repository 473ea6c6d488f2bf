// mlp_simt.cu -- fused shared per-point MLP on the fp32 CUDA cores (impl 0).
// Reference: pointnet_utils.py:234-246 (SA-MSG: group, centre, concat, 3x conv1x1+BN+ReLU, max
// over nsample), :319-343 (group-all), :291-298 (FP MLP), backbones.py:68 (conv1+bn1+relu).
//
// The reference materialises the grouped tensor [B,C,S,K] in HBM, then runs cuDNN conv, BN,
// ReLU and max as separate kernels (172 MB of activations per cloud round-trip HBM, SURVEY
// App. B).  Here ONE kernel per SA scale / FP stage does everything: a CTA owns a 64-row tile,
// assembles the layer-0 input straight from the index lists (gather + centre-subtract +
// concat on the fly), keeps every inter-layer activation in shared memory (transposed,
// [channel][row]), streams the BN-folded weights through a 16-deep K chunk, and reduces the
// max over nsample before anything is written.  HBM traffic = inputs + weights + final output.
//
// This is the exact-fp32 path: it is the parity reference for the tcgen05 path (impl 1) and the
// fallback for shapes that one does not cover.  Register tile: 4 rows x (4|8) columns/thread.
#include "mlp_common.cuh"

#include <math.h>

namespace captra {

constexpr int MS_THREADS = 256;
constexpr int MS_LDR = MLP_TR + 4;  // row stride of transposed activation tiles (keeps float4 alignment)

struct MlpSimtArgs {
    int nlayers, cin0, relu_last;
    int cin_pad[CAPTRA_MAX_MLP_LAYERS], cout[CAPTRA_MAX_MLP_LAYERS], coutp[CAPTRA_MAX_MLP_LAYERS], cw[CAPTRA_MAX_MLP_LAYERS];
    const float *wt[CAPTRA_MAX_MLP_LAYERS], *bias[CAPTRA_MAX_MLP_LAYERS];
    int64_t rows;      // total rows
    int group;         // max over each `group` consecutive rows (0: none)
    int cpt, tpc;      // groups per tile (group <= 64) / tiles per group (group > 64)
    int64_t ngroups;
    float *out; int64_t ldo; int col_off;
    // layer-0 input, SA mode
    int n, s, cfeat;
    const float *xyz, *new_xyz, *feats; const int *idx;
    // layer-0 input, dense mode: row = [segA row (ca) | segB row (cb)], segB row = row / bcast if bcast
    const float *segA; int64_t ldA; int ca;
    const float *segB; int64_t ldB; int cb; int bcast;
    int actA_rows, actB_rows, runmax_n;
};

template <int MODE>  // 0: SA gather, 1: dense
__global__ void __launch_bounds__(MS_THREADS) mlp_simt_kernel(MlpSimtArgs a) {
    extern __shared__ float4 smem4[];
    float *smem = reinterpret_cast<float *>(smem4);
    float *Xs = smem;                                  // [16][MS_LDR]
    float *Ws = Xs + MLP_KC * MS_LDR;                  // [16][128]
    float *act0 = Ws + MLP_KC * 128;                   // [actA_rows][MS_LDR]
    float *act1 = act0 + (size_t)a.actA_rows * MS_LDR; // [actB_rows][MS_LDR]
    float *runmax = act1 + (size_t)a.actB_rows * MS_LDR;
    __shared__ long long s_row[MLP_TR];   // global row, -1 = padding
    __shared__ int s_pt[MLP_TR];          // SA: b*n + point index
    __shared__ float s_ctr[MLP_TR * 3];   // SA: centroid
    __shared__ long long s_grp[MLP_TR];   // group id per group slot of this tile

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int K = a.group;

    for (int sub = 0; sub < a.tpc; ++sub) {
        __syncthreads();
        // ---- tile row metadata ----------------------------------------------------------
        if (tid < MLP_TR) {
            long long grow = -1;
            const int r = tid;
            if (K == 0) {
                const long long g = (long long)blockIdx.x * MLP_TR + r;
                if (g < a.rows) grow = g;
            } else if (a.tpc == 1) {
                const int slot = r / K, kk = r - slot * K;
                const long long gid = (long long)blockIdx.x * a.cpt + slot;
                if (slot < a.cpt && gid < a.ngroups) grow = gid * K + kk;
                if (kk == 0 && slot < a.cpt) s_grp[slot] = gid < a.ngroups ? gid : -1;
            } else {
                const int kk = sub * MLP_TR + r;
                if (kk < K) grow = (long long)blockIdx.x * K + kk;
                if (r == 0) s_grp[0] = blockIdx.x;
            }
            s_row[r] = grow;
            if (MODE == 0 && grow >= 0) {
                const long long cen = grow / K;                 // b*s + centroid
                const long long b = cen / a.s;
                s_pt[r] = (int)(b * a.n + __ldg(a.idx + grow));
                s_ctr[r * 3 + 0] = __ldg(a.new_xyz + cen * 3 + 0);
                s_ctr[r * 3 + 1] = __ldg(a.new_xyz + cen * 3 + 1);
                s_ctr[r * 3 + 2] = __ldg(a.new_xyz + cen * 3 + 2);
            }
        }
        if (sub == 0 && a.tpc > 1)
            for (int i = tid; i < a.runmax_n; i += MS_THREADS) runmax[i] = -INFINITY;
        __syncthreads();

        for (int l = 0; l < a.nlayers; ++l) {
            const bool last = (l == a.nlayers - 1);
            const bool relu = !last || a.relu_last;
            const float *in = (l == 0) ? Xs : ((l - 1) & 1 ? act1 : act0);
            float *outb = (l & 1) ? act1 : act0;
            const int Kpad = a.cin_pad[l], coutp = a.coutp[l], cout = a.cout[l];
            const bool wide = a.cw[l] == 128;
            const int CW = wide ? 128 : 64;
            const float *wt = a.wt[l];

            for (int col0 = 0; col0 < coutp; col0 += CW) {
                float acc[4][8];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

                for (int k0 = 0; k0 < Kpad; k0 += MLP_KC) {
                    __syncthreads();
                    // weights chunk [16][CW]
                    for (int i = tid; i < MLP_KC * CW / 4; i += MS_THREADS) {
                        const int kk = i / (CW / 4), c4 = i - kk * (CW / 4);
                        const float4 v = __ldg(reinterpret_cast<const float4 *>(wt + (size_t)(k0 + kk) * coutp + col0) + c4);
                        *reinterpret_cast<float4 *>(Ws + kk * 128 + c4 * 4) = v;
                    }
                    if (l == 0) {  // assemble the layer-0 input chunk [16][64] (transposed)
                        const int kk = tid & 15, r0 = tid >> 4, c = k0 + kk;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int r = r0 + 16 * i;
                            float v = 0.f;
                            const long long grow = s_row[r];
                            if (grow >= 0) {
                                if (MODE == 0) {
                                    if (c < a.cfeat) v = __ldg(a.feats + (size_t)s_pt[r] * a.cfeat + c);
                                    else if (c < a.cfeat + 3)
                                        v = __fsub_rn(__ldg(a.xyz + (size_t)s_pt[r] * 3 + (c - a.cfeat)), s_ctr[r * 3 + (c - a.cfeat)]);
                                } else {
                                    if (c < a.ca) v = __ldg(a.segA + grow * a.ldA + c);
                                    else if (c < a.ca + a.cb)
                                        v = __ldg(a.segB + (a.bcast ? grow / a.bcast : grow) * a.ldB + (c - a.ca));
                                }
                            }
                            Xs[kk * MS_LDR + r] = v;
                        }
                    }
                    __syncthreads();
                    const float *xin = (l == 0) ? Xs : in + (size_t)k0 * MS_LDR;
#pragma unroll
                    for (int kk = 0; kk < MLP_KC; ++kk) {
                        const float4 av = *reinterpret_cast<const float4 *>(xin + kk * MS_LDR + ty * 4);
                        const float4 w0 = *reinterpret_cast<const float4 *>(Ws + kk * 128 + tx * 4);
                        const float ar[4] = {av.x, av.y, av.z, av.w};
                        const float wr0[4] = {w0.x, w0.y, w0.z, w0.w};
#pragma unroll
                        for (int i = 0; i < 4; ++i)
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], wr0[j], acc[i][j]);
                        if (wide) {
                            const float4 w1 = *reinterpret_cast<const float4 *>(Ws + kk * 128 + 64 + tx * 4);
                            const float wr1[4] = {w1.x, w1.y, w1.z, w1.w};
#pragma unroll
                            for (int i = 0; i < 4; ++i)
#pragma unroll
                                for (int j = 0; j < 4; ++j) acc[i][4 + j] = fmaf(ar[i], wr1[j], acc[i][4 + j]);
                        }
                    }
                }

                // ---- epilogue of this column pass ---------------------------------------
                const int nhalf = wide ? 2 : 1;
                const bool vec_out = ((a.ldo | a.col_off) & 3) == 0 && (reinterpret_cast<uintptr_t>(a.out) & 15) == 0;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (h >= nhalf) break;
                    const int cl0 = h * 64 + tx * 4;  // first of this thread's 4 columns within the pass
                    float v[4][4];                    // [row][col]
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float bz = __ldg(a.bias[l] + col0 + cl0 + j);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const float t = acc[i][h * 4 + j] + bz;
                            v[i][j] = relu ? fmaxf(t, 0.f) : t;
                        }
                    }
                    if (!last || K > 0) {  // keep on chip, transposed [channel][row]
                        const int cbase = last ? cl0 : col0 + cl0;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            *reinterpret_cast<float4 *>(outb + (size_t)(cbase + j) * MS_LDR + ty * 4) =
                                make_float4(v[0][j], v[1][j], v[2][j], v[3][j]);
                    } else {               // final rows straight to global, point-major
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const long long grow = s_row[ty * 4 + i];
                            if (grow < 0) continue;
                            float *dst = a.out + grow * a.ldo + a.col_off + col0 + cl0;
                            if (vec_out && col0 + cl0 + 3 < cout) {
                                *reinterpret_cast<float4 *>(dst) = make_float4(v[i][0], v[i][1], v[i][2], v[i][3]);
                            } else {
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    if (col0 + cl0 + j < cout) dst[j] = v[i][j];
                            }
                        }
                    }
                }
                if (last && K > 0) {  // max over the rows of each group present in this tile
                    __syncthreads();
                    const int nslots = (a.tpc == 1) ? a.cpt : 1;
                    for (int o = tid; o < nslots * CW; o += MS_THREADS) {
                        const int slot = o / CW, cl = o - slot * CW;
                        const long long gid = s_grp[slot];
                        if (gid < 0 || col0 + cl >= cout) continue;
                        int rbeg, rend;
                        if (a.tpc == 1) { rbeg = slot * K; rend = rbeg + K; }
                        else { rbeg = 0; rend = min(MLP_TR, K - sub * MLP_TR); }
                        float m = -INFINITY;
                        for (int r = rbeg; r < rend; ++r) m = fmaxf(m, outb[(size_t)cl * MS_LDR + r]);
                        if (a.tpc > 1) {
                            m = fmaxf(m, runmax[col0 + cl]);
                            runmax[col0 + cl] = m;
                            if (sub != a.tpc - 1) continue;
                        }
                        a.out[gid * a.ldo + a.col_off + col0 + cl] = m;
                    }
                }
            }
        }
    }
}

__global__ void pack_simt_kernel(int cin, int cout, int cin_pad, int coutp, const float *__restrict__ w,
                                 const float *__restrict__ bias, float *__restrict__ wt, float *__restrict__ bp) {
    const int total = cin_pad * coutp;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int k = i / coutp, c = i - k * coutp;
        wt[i] = (k < cin && c < cout) ? w[(size_t)c * cin + k] : 0.f;
    }
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < coutp; c += gridDim.x * blockDim.x)
        bp[c] = (c < cout && bias) ? bias[c] : 0.f;
}

static int fill_common(MlpSimtArgs &a, const captra_mlp_desc *d, const void *packed, size_t *smem_bytes) {
    const SimtLayout L = simt_layout(*d);
    a.nlayers = d->nlayers; a.cin0 = d->cin; a.relu_last = d->relu_last;
    int actA = 128, actB = 128;
    for (int l = 0; l < d->nlayers; ++l) {
        a.cin_pad[l] = L.cin_pad[l]; a.cout[l] = L.cout[l]; a.coutp[l] = L.coutp[l]; a.cw[l] = L.cw[l];
        a.wt[l] = reinterpret_cast<const float *>(packed) + L.off_w[l];
        a.bias[l] = reinterpret_cast<const float *>(packed) + L.off_b[l];
        if (l < d->nlayers - 1) {
            if (l & 1) actB = max(actB, L.coutp[l]); else actA = max(actA, L.coutp[l]);
        }
    }
    a.actA_rows = actA; a.actB_rows = actB;
    a.runmax_n = L.coutp[d->nlayers - 1];
    *smem_bytes = sizeof(float) * ((size_t)MLP_KC * MS_LDR + MLP_KC * 128 + (size_t)(actA + actB) * MS_LDR + a.runmax_n);
    CAPTRA_REQUIRE(*smem_bytes <= 227 * 1024 - 2048, "mlp(simt): layer widths need %zu B of shared memory", *smem_bytes);
    return CAPTRA_OK;
}

template <int MODE>
static int launch_simt(const MlpSimtArgs &a, int64_t nblocks, size_t smem, cudaStream_t stream) {
    auto kern = mlp_simt_kernel<MODE>;
    CAPTRA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CAPTRA_REQUIRE(nblocks <= 0x7fffffffLL, "mlp(simt): too many tiles");
    kern<<<(unsigned)nblocks, MS_THREADS, smem, stream>>>(a);
    CAPTRA_CHECK_LAUNCH("mlp_simt");
    return CAPTRA_OK;
}

static void set_grouping(MlpSimtArgs &a, int64_t rows, int group, int64_t *nblocks) {
    a.rows = rows; a.group = group;
    if (group <= 0) {
        a.cpt = 1; a.tpc = 1; a.ngroups = 0;
        *nblocks = ceil_div<int64_t>(rows, MLP_TR);
    } else if (group <= MLP_TR) {
        a.cpt = MLP_TR / group; a.tpc = 1; a.ngroups = rows / group;
        *nblocks = ceil_div<int64_t>(a.ngroups, a.cpt);
    } else {
        a.cpt = 1; a.tpc = ceil_div(group, MLP_TR); a.ngroups = rows / group;
        *nblocks = a.ngroups;
    }
}

int simt_sa_mlp_max(int b, int n, int s, int k, int cfeat, const float *xyz, const float *new_xyz,
                    const float *feats, const int *idx, const captra_mlp_desc *d, const void *packed,
                    float *out, int64_t ldo, int col_off, cudaStream_t stream) {
    MlpSimtArgs a{};
    size_t smem;
    int rc = fill_common(a, d, packed, &smem);
    if (rc) return rc;
    int64_t nblocks;
    set_grouping(a, (int64_t)b * s * k, k, &nblocks);
    a.out = out; a.ldo = ldo; a.col_off = col_off;
    a.n = n; a.s = s; a.cfeat = cfeat; a.xyz = xyz; a.new_xyz = new_xyz; a.feats = feats; a.idx = idx;
    return launch_simt<0>(a, nblocks, smem, stream);
}

int simt_point_mlp(int64_t rows, const float *segA, int64_t ldA, int ca, const float *segB, int64_t ldB,
                   int cb, int bcast, const captra_mlp_desc *d, const void *packed, float *y, int64_t ldy,
                   int col_off, int group, cudaStream_t stream) {
    MlpSimtArgs a{};
    size_t smem;
    int rc = fill_common(a, d, packed, &smem);
    if (rc) return rc;
    int64_t nblocks;
    set_grouping(a, rows, group, &nblocks);
    a.out = y; a.ldo = ldy; a.col_off = col_off;
    a.segA = segA; a.ldA = ldA; a.ca = ca; a.segB = segB; a.ldB = ldB; a.cb = cb; a.bcast = bcast;
    return launch_simt<1>(a, nblocks, smem, stream);
}

int simt_pack(const captra_mlp_desc *d, void *packed, cudaStream_t stream) {
    const SimtLayout L = simt_layout(*d);
    for (int l = 0; l < d->nlayers; ++l) {
        float *base = reinterpret_cast<float *>(packed);
        pack_simt_kernel<<<64, 256, 0, stream>>>(L.cin[l], L.cout[l], L.cin_pad[l], L.coutp[l], d->w[l], d->bias[l],
                                                 base + L.off_w[l], base + L.off_b[l]);
        CAPTRA_CHECK_LAUNCH("mlp_pack(simt)");
    }
    return CAPTRA_OK;
}

}  // namespace captra
