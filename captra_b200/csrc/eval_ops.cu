// eval_ops.cu -- the per-frame loss / eval reductions of the tracking loop on the device,
// two launches per frame, deterministic (no atomics), CUDA-graph capturable:
//   pose_utils/part_dof_utils.py:40-67   eval_part_model / eval_part_full: sdiff, tdiff, rdiff (degrees; y axis only
//                                        for symmetric categories, metrics.py:5-34), 5deg5cm, 10deg10cm per part
//   network/models/loss.py:42-70,122-134 compute_nocs_loss (l2, self_supervise=False) and compute_miou_loss, as
//                                        EvalTrackModel.compute_loss calls them (model.py:546-561)
// Everything is emitted as SUMS + counts, so one all-reduce(SUM) over ranks followed by the divisions gives the
// batch means the reference logs (test.py:87-99) whatever the sharding.
#include "common.cuh"

#include <math.h>

namespace captra {

constexpr int EV_THREADS = 256;
constexpr int EV_MAX_SEG = 8;
constexpr float kLossEps = 1e-6f;     // loss.py:9

// per cloud: sum_c mIoU[b,c], sum of masked l2 NOCS errors, mask count
__global__ void __launch_bounds__(EV_THREADS) track_loss_partials_kernel(int n, int nseg, int p, const float *__restrict__ seg,
                                                                         const float *__restrict__ nocs,
                                                                         const int64_t *__restrict__ gt_labels,
                                                                         const float *__restrict__ gt_nocs,
                                                                         const int64_t *__restrict__ pred_labels,
                                                                         float *__restrict__ partial /* [B][3] */) {
    __shared__ float red[2 * EV_MAX_SEG + 2][EV_THREADS / 32];
    const int b = blockIdx.x;
    float I[EV_MAX_SEG], S[EV_MAX_SEG], nsum = 0.f, msum = 0.f;
#pragma unroll
    for (int c = 0; c < EV_MAX_SEG; ++c) I[c] = S[c] = 0.f;
    for (int i = threadIdx.x; i < n; i += EV_THREADS) {
        if (gt_labels) {
            // loss.py:122-128: I = sum_n pred * onehot(gt), U = sum_n (pred + onehot(gt)) - I
            const int g = (int)__ldg(gt_labels + (size_t)b * n + i);
#pragma unroll
            for (int c = 0; c < EV_MAX_SEG; ++c)
                if (c < nseg) {
                    const float pr = __ldg(seg + ((size_t)b * nseg + c) * n + i);
                    const float oh = g == c ? 1.f : 0.f;
                    I[c] += pr * oh;
                    S[c] += pr + oh;
                }
        }
        if (gt_nocs) {
            // loss.py:52-66: pick the predicted part's coordinates (labels >= P read zeros and are masked out when
            // P > 1; with a single part there is no mask), l2 norm of the difference
            const int l = (int)__ldg(pred_labels + (size_t)b * n + i);
            const bool pick = p > 1;
            const bool in = !pick || l < p;
            float d2 = 0.f;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float pv = pick ? (l < p ? __ldg(nocs + ((size_t)b * 3 * p + 3 * l + c) * n + i) : 0.f)
                                      : __ldg(nocs + ((size_t)b * 3 + c) * n + i);
                const float d = pv - __ldg(gt_nocs + ((size_t)b * 3 + c) * n + i);
                d2 += d * d;
            }
            if (in) { nsum += sqrtf(d2); msum += 1.f; }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    auto wsum = [](float v) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
        return v;
    };
#pragma unroll
    for (int c = 0; c < EV_MAX_SEG; ++c) {
        const float a = wsum(I[c]), s = wsum(S[c]);
        if (lane == 0) { red[c][warp] = a; red[EV_MAX_SEG + c][warp] = s; }
    }
    {
        const float a = wsum(nsum), s = wsum(msum);
        if (lane == 0) { red[2 * EV_MAX_SEG][warp] = a; red[2 * EV_MAX_SEG + 1][warp] = s; }
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    auto tot = [&](int k) { float s = 0.f; for (int w = 0; w < EV_THREADS / 32; ++w) s += red[k][w]; return s; };
    float miou = 0.f;
    for (int c = 0; c < nseg; ++c) {
        const float i_ = tot(c), u_ = tot(EV_MAX_SEG + c) - i_;
        miou += i_ / (u_ + kLossEps);
    }
    partial[b * 3 + 0] = miou;
    partial[b * 3 + 1] = tot(2 * EV_MAX_SEG);
    partial[b * 3 + 2] = tot(2 * EV_MAX_SEG + 1);
}

struct EvalArgs {
    int b, p, sym, nseg, n, have_seg, have_nocs, accumulate;
    const float *gt_R, *gt_t, *gt_s, *pr_R, *pr_t, *pr_s;
    const float *partial;        // [B][3] or null
    float *per_instance;         // [B,P,5] or null: sdiff, tdiff, rdiff, 5deg5cm, 10deg10cm
    float *sums;                 // [5P + 1 + 4]
};

__global__ void __launch_bounds__(EV_THREADS) eval_part_full_kernel(EvalArgs a) {
    extern __shared__ float diffs[];      // [B*P][5]
    const int total = a.b * a.p;
    for (int e = threadIdx.x; e < total; e += EV_THREADS) {
        const float *R1 = a.gt_R + (size_t)e * 9, *R2 = a.pr_R + (size_t)e * 9;
        float d;
        if (a.sym) {                     // metrics.py:12-16: angle between the y axes (column 1)
            d = R1[1] * R2[1] + R1[4] * R2[4] + R1[7] * R2[7];
        } else {                         // metrics.py:24-28: (trace(R1 R2^T) - 1) / 2
            float tr = 0.f;
            for (int i = 0; i < 3; ++i)
                for (int k = 0; k < 3; ++k) tr += R1[i * 3 + k] * R2[i * 3 + k];
            d = (tr - 1.f) / 2.f;
        }
        d = fminf(fmaxf(d, -1.f), 1.f);
        const float rdiff = acosf(d) / 3.14159265358979323846f * 180.f;
        float t2 = 0.f;
        for (int c = 0; c < 3; ++c) { const float dt = a.gt_t[(size_t)e * 3 + c] - a.pr_t[(size_t)e * 3 + c]; t2 += dt * dt; }
        const float tdiff = sqrtf(t2), sdiff = fabsf(a.gt_s[e] - a.pr_s[e]);
        float *o = diffs + (size_t)e * 5;
        o[0] = sdiff; o[1] = tdiff; o[2] = rdiff;
        o[3] = (rdiff <= 5.f && tdiff <= 0.05f) ? 1.f : 0.f;          // part_dof_utils.py:56-57
        o[4] = (rdiff <= 10.f && tdiff <= 0.10f) ? 1.f : 0.f;
        if (a.per_instance)
            for (int k = 0; k < 5; ++k) a.per_instance[(size_t)e * 5 + k] = o[k];
    }
    __syncthreads();
    // fixed-order sums (deterministic): thread k < 5P owns one (part, quantity)
    if (threadIdx.x < 5 * a.p) {
        const int part = threadIdx.x / 5, q = threadIdx.x % 5;
        double s = 0.0;
        for (int bb = 0; bb < a.b; ++bb) s += diffs[((size_t)bb * a.p + part) * 5 + q];
        a.sums[part * 5 + q] = (a.accumulate ? a.sums[part * 5 + q] : 0.f) + (float)s;
    }
    if (threadIdx.x == EV_THREADS - 1) {
        float *o = a.sums + 5 * a.p;
        const float keep = a.accumulate ? 1.f : 0.f;
        double miou = 0.0, ns = 0.0, ms = 0.0;
        if (a.partial)
            for (int bb = 0; bb < a.b; ++bb) { miou += a.partial[bb * 3]; ns += a.partial[bb * 3 + 1]; ms += a.partial[bb * 3 + 2]; }
        o[0] = keep * o[0] + (float)a.b;                                       // clouds (x frames when accumulating)
        o[1] = keep * o[1] + (a.have_seg ? (float)miou : 0.f);                 // sum over (cloud, class) of mIoU
        o[2] = keep * o[2] + (a.have_seg ? (float)(a.b * a.nseg) : 0.f);       //   ... and how many there are
        o[3] = keep * o[3] + (a.have_nocs ? (float)ns : 0.f);                  // sum of (masked) l2 NOCS errors
        o[4] = keep * o[4] + (a.have_nocs ? (float)ms : 0.f);                  //   ... and the mask count (B*N for one part)
    }
}

}  // namespace captra

using namespace captra;

extern "C" int captra_track_eval(int b, int p, int n, int nseg, int sym, const float *gt_rotation, const float *gt_translation,
                                 const float *gt_scale, const float *rotation, const float *translation, const float *scale,
                                 const float *seg, const float *nocs, const int64_t *pred_labels, const int64_t *gt_labels,
                                 const float *gt_nocs, float *scratch, float *per_instance, float *sums,
                                 int accumulate, captra_stream_t stream) {
    CAPTRA_REQUIRE(b >= 0 && p >= 1 && p <= 8 && n >= 0 && nseg >= 0 && nseg <= EV_MAX_SEG, "track_eval: bad sizes");
    CAPTRA_REQUIRE(sums, "track_eval: null output");
    CAPTRA_REQUIRE(b == 0 || (gt_rotation && gt_translation && gt_scale && rotation && translation && scale), "track_eval: null pose");
    CAPTRA_REQUIRE((size_t)b * p * 5 * sizeof(float) <= 200 * 1024, "track_eval: batch too large for one block (%d x %d)", b, p);
    const bool have_seg = gt_labels && seg && nseg > 0, have_nocs = gt_nocs && nocs && pred_labels;
    cudaStream_t s = as_stream(stream);
    if (b > 0 && (have_seg || have_nocs)) {
        CAPTRA_REQUIRE(scratch, "track_eval: scratch [B*3] floats needed for the loss partials");
        track_loss_partials_kernel<<<b, EV_THREADS, 0, s>>>(n, have_seg ? nseg : 0, p, seg, nocs, have_seg ? gt_labels : nullptr,
                                                            have_nocs ? gt_nocs : nullptr, pred_labels, scratch);
        CAPTRA_CHECK_LAUNCH("track_eval(loss partials)");
    }
    EvalArgs a;
    a.b = b; a.p = p; a.sym = sym; a.nseg = nseg; a.n = n; a.have_seg = have_seg; a.have_nocs = have_nocs;
    a.accumulate = accumulate;
    a.gt_R = gt_rotation; a.gt_t = gt_translation; a.gt_s = gt_scale; a.pr_R = rotation; a.pr_t = translation; a.pr_s = scale;
    a.partial = (have_seg || have_nocs) ? scratch : nullptr; a.per_instance = per_instance; a.sums = sums;
    const size_t smem = (size_t)b * p * 5 * sizeof(float);
    if (smem > 48 * 1024) CAPTRA_CUDA(cudaFuncSetAttribute(eval_part_full_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    eval_part_full_kernel<<<1, EV_THREADS, smem, s>>>(a);
    CAPTRA_CHECK_LAUNCH("track_eval");
    return CAPTRA_OK;
}
