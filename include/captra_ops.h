/*
 * captra_ops.h -- C ABI of libcaptra_ops.so, the B200-native (sm_100a) replacement for the
 * launcher layer of CAPTRA's `pointnet2_cuda` extension plus the device-side pose fit.
 *
 * Conventions (all entry points):
 *   - plain pointers + sizes, no torch types; every pointer is a DEVICE pointer unless the
 *     name ends in `_host`; `stream` is a cudaStream_t passed as void*;
 *   - the caller allocates every output (same ownership as the reference,
 *     pointnet2_utils.py:26-27,58,98-99,130-131,167,213,261);
 *   - launches are asynchronous on `stream`, no internal sync, no global state, re-entrant;
 *   - return value: CAPTRA_OK, or an error code -- the reference launchers print and
 *     exit(-1) on a CUDA error (e.g. sampling_gpu.cu:248-252); we return the code and let the
 *     binding raise.  captra_last_error() gives a thread-local message.
 *
 * Citations are into /root/reference/network/models/pointnet_lib/src/ (the FFI a reference
 * maintainer would re-bind).
 */
#ifndef CAPTRA_OPS_H_
#define CAPTRA_OPS_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAPTRA_OK 0
#define CAPTRA_ERR_INVALID_ARG 1
#define CAPTRA_ERR_CUDA 2
#define CAPTRA_ERR_UNSUPPORTED 3

typedef void *captra_stream_t; /* cudaStream_t */

const char *captra_last_error(void);
int captra_abi_version(void);
/* number of kernels launched by this library in this process (bench.py's gpu_launches) */
int64_t captra_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * 1. Drop-in launchers: same names, argument order and meaning as the reference's
 *    *_kernel_launcher_fast functions; only the return type changes (void -> int status).
 * ---------------------------------------------------------------------------------------- */

/* ball_query_gpu.cu:48-49 (positional order is new_xyz, xyz -- ball_query_gpu.h:12 swaps the
 * names).  new_xyz [B,M,3], xyz [B,N,3] -> idx [B,M,nsample] int32; rows of empty balls are
 * left untouched (caller zeroes, pointnet2_utils.py:261). */
int ball_query_kernel_launcher_fast(int b, int n, int m, float radius, int nsample,
                                    const float *new_xyz, const float *xyz, int *idx,
                                    captra_stream_t stream);

/* group_points_gpu.h:13-14.  points [B,C,N], idx [B,npoints,nsample] -> out [B,C,npoints,nsample] */
int group_points_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample,
                                      const float *points, const int *idx, float *out,
                                      captra_stream_t stream);
/* group_points_gpu.h:19-20.  grad_points [B,C,N] must be zeroed by the caller (:231) */
int group_points_grad_kernel_launcher_fast(int b, int c, int n, int npoints, int nsample,
                                           const float *grad_out, const int *idx,
                                           float *grad_points, captra_stream_t stream);

/* sampling_gpu.h:12-13.  points [B,C,N], idx [B,npoints] -> out [B,C,npoints] */
int gather_points_kernel_launcher_fast(int b, int c, int n, int npoints, const float *points,
                                       const int *idx, float *out, captra_stream_t stream);
/* sampling_gpu.h:19-20 */
int gather_points_grad_kernel_launcher_fast(int b, int c, int n, int npoints,
                                            const float *grad_out, const int *idx,
                                            float *grad_points, captra_stream_t stream);

/* sampling_gpu.h:26-27.  dataset [B,N,3], temp [B,N] (caller pre-fills 1e10; read at entry and
 * written back at exit exactly like the reference's in-place scratch) -> idxs [B,M] int32.
 * Tie rule of the reference's shared-memory tournament is reproduced exactly
 * (sampling_gpu.cu:86-91,119-203). */
int furthest_point_sampling_kernel_launcher(int b, int n, int m, const float *dataset,
                                            float *temp, int *idxs, captra_stream_t stream);

/* interpolate_gpu.h:13-14.  unknown [B,n,3], known [B,m,3] -> dist2 [B,n,3] (SQUARED), idx [B,n,3] */
int three_nn_kernel_launcher_fast(int b, int n, int m, const float *unknown,
                                  const float *known, float *dist2, int *idx,
                                  captra_stream_t stream);
/* interpolate_gpu.h:19-20 (k <= 200 as in the reference, interpolate_gpu.cu:30) */
int knn_kernel_launcher_fast(int b, int n, int m, int k, const float *unknown,
                             const float *known, float *dist2, int *idx,
                             captra_stream_t stream);
/* interpolate_gpu.h:26-27.  points [B,C,m], idx/weight [B,n,3] -> out [B,C,n] */
int three_interpolate_kernel_launcher_fast(int b, int c, int m, int n, const float *points,
                                           const int *idx, const float *weight, float *out,
                                           captra_stream_t stream);
/* interpolate_gpu.h:33-34.  NOTE (b,c,n,m) order, unlike the forward's (b,c,m,n). */
int three_interpolate_grad_kernel_launcher_fast(int b, int c, int n, int m,
                                                const float *grad_out, const int *idx,
                                                const float *weight, float *grad_points,
                                                captra_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * 2. Fused entry points used by the host-side mirror of network/models/pointnet_utils.py.
 *    Same per-element arithmetic as the ops above; they only remove HBM round trips.
 * ---------------------------------------------------------------------------------------- */

/* One scan of xyz answers up to CAPTRA_MAX_RADII ball queries that share the centroids
 * (the MSG loop of pointnet_utils.py:228-233).  idx_r has nsample_r columns. */
#define CAPTRA_MAX_RADII 4
int captra_ball_query_multi(int b, int n, int m, int nradii, const float *radii_host,
                            const int *nsamples_host, const float *new_xyz, const float *xyz,
                            int *const *idx_host_ptrs, captra_stream_t stream);

/* QueryAndGroup (pointnet2_utils.py:271-303) as one call: idx [B,M,K] = ball_query(radius, K, xyz, new_xyz) (idx
 * must be zeroed by the caller, as for the launcher above) and grouped [B,C,M,K] = group_points(points [B,C,N], idx).
 * For C <= 8 on a binned cloud the query warp writes its grouped rows itself (one launch set, idx never re-read);
 * wider rows run the query and the shared-memory-staged gather back to back. */
int captra_ball_query_group(int b, int n, int m, int c, float radius, int nsample, const float *new_xyz,
                            const float *xyz, const float *points, int *idx, float *grouped,
                            captra_stream_t stream);

/* FPS + the gather that always follows it (pointnet_utils.py:225-226): additionally writes
 * new_xyz [B,M,3] = dataset[b, idxs[b,:], :].  new_xyz may be NULL.  temp may be NULL for n <= 32768 (the running
 * distances then start at 1e10, as pointnet2_utils.py:27 fills them, and live in registers only: one CTA per cloud
 * up to 8192 points, a thread-block cluster of 4 / 8 CTAs exchanging through distributed shared memory above). */
int captra_fps_gather(int b, int n, int m, const float *dataset, float *temp, int *idxs,
                      float *new_xyz, captra_stream_t stream);

/* The sampling + grouping-index step of an MSG set-abstraction layer (pointnet_utils.py:225-233) as a PIPELINE: FPS
 * (-> fps_idx [B,M], new_xyz [B,M,3]) and the multi-radius ball query around the sampled centroids run CONCURRENTLY.
 * FPS is a latency chain that keeps one SM per cloud busy; it publishes every 16 rounds how many centroids are
 * final (progress [B] ints, zeroed by the caller) and lets its stream successor start early (programmatic dependent
 * launch); the ball query's blocks, laid out in the order FPS completes them, wait on that counter.  Results are
 * identical to captra_fps_gather followed by captra_ball_query_multi.  n <= 8192; idx buffers zeroed by the caller. */
int captra_fps_ball_query(int b, int n, int m, const float *xyz, int *fps_idx, float *new_xyz, int nradii,
                          const float *radii_host, const int *nsamples_host, int *const *idx_host_ptrs,
                          int *progress, captra_stream_t stream);

/* three_nn + inverse-distance weights (pointnet_utils.py:284-287) + three_interpolate:
 * out[., i] = sum_j w_j * points[., idx_j].  idx and weight [B,n,3] are required scratch /
 * outputs; dist (the sqrt'ed distance ThreeNN returns, pointnet2_utils.py:134) may be NULL.
 * point_major == 0: reference layout, points [B,C,m] -> out [B,C,n].
 * point_major == 1: points [B,m,C] -> out row (b*n+i) at out + row*out_row_stride +
 *                   out_col_offset (lets the caller write straight into a concat buffer).
 * unknown == NULL skips the search: idx and weight are then inputs (the two networks of a rigid
 * object see the same canonicalised cloud and share one search). */
int captra_three_nn_interpolate(int b, int c, int n, int m, const float *unknown,
                                const float *known, const float *points, float *out,
                                float *dist, int *idx, float *weight, int point_major,
                                int64_t out_row_stride, int out_col_offset,
                                captra_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * 3. Shared per-point MLPs (pointnet_utils.py:242-246, :296-298, :337-341; backbones.py:68).
 *    Weights are passed pre-folded (conv bias + eval-mode BatchNorm -> W', b'), row-major
 *    W'[cout][cin], fp32.
 * ---------------------------------------------------------------------------------------- */
#define CAPTRA_MAX_MLP_LAYERS 4

typedef struct {
    int nlayers;                          /* 1..CAPTRA_MAX_MLP_LAYERS */
    int cin;                              /* input channels of layer 0 */
    int cout[CAPTRA_MAX_MLP_LAYERS];      /* output channels per layer */
    const float *w[CAPTRA_MAX_MLP_LAYERS];/* device, [cout_l][cin_l] */
    const float *bias[CAPTRA_MAX_MLP_LAYERS]; /* device, [cout_l] */
    int relu_last;                        /* apply ReLU after the last layer */
} captra_mlp_desc;

/* Weights are re-laid-out once per model load into a caller-owned device buffer ("packed",
 * 16-byte aligned, captra_mlp_pack_bytes() bytes); the execution calls take the same desc (only
 * nlayers/cin/cout/relu_last are read there) plus that buffer.  impl: 0 = exact-fp32 CUDA-core
 * kernel, 1 = tcgen05 3xTF32 kernel (packs are impl-specific). */
int64_t captra_mlp_pack_bytes(const captra_mlp_desc *mlp, int impl);
int captra_mlp_pack(const captra_mlp_desc *mlp, int impl, void *packed, captra_stream_t stream);

/* Fused set-abstraction scale (pointnet_utils.py:233-246): for every centroid s and sample k
 *   row = [ feats[b, idx[b,s,k], :] (cfeat ch) , xyz[b, idx[b,s,k], :] - new_xyz[b,s,:] (3 ch) ]
 * -> MLP (ReLU after every layer) -> max over k -> out[(b*S+s)*ldo + col_off + c].
 * feats is POINT-major [B,N,cfeat] (NULL when cfeat==0), xyz [B,N,3], new_xyz [B,S,3],
 * idx [B,S,K] int32; out is point-major with row stride ldo, so the scales of one MSG layer
 * write side by side into one [B,S,sum(cout)] tensor (the torch.cat of pointnet_utils.py:249). */
int captra_sa_mlp_max(int b, int n, int s, int k, int cfeat, const float *xyz,
                      const float *new_xyz, const float *feats, const int *idx,
                      const captra_mlp_desc *mlp, const void *packed, float *out, int64_t ldo,
                      int col_off, int impl, captra_stream_t stream);

/* The same set-abstraction scale with layer 0 PROJECTED per point.  Layer 0 is linear in
 * [features | xyz - centroid] (pointnet_utils.py:239-246: conv -> BN -> ReLU on the concatenation), and its
 * feature part depends on the point only, while every point sits in many balls.  The caller computes
 *   pre[b*N + j, :] = W0[:, :cfeat] * feats[b, j, :]            (one captra_point_mlp launch, no bias / ReLU)
 * once per cloud, and this entry forms
 *   h0 = relu(pre[idx] + wxyz_bias[0]*dx + wxyz_bias[1]*dy + wxyz_bias[2]*dz + wxyz_bias[3])
 * on the fly (wxyz_bias is [4][cpre]: the three coordinate columns of the BN-folded W0, then its bias) and
 * runs layers 1.. (`mlp`, cin == cpre) and the max as above.  pre has row stride ldpre, so the scales of
 * one MSG layer read column blocks of one projected buffer.  tcgen05 paths (impl 1, 2) only. */
int captra_sa_mlp_max_pre(int b, int n, int s, int k, int cpre, const float *xyz, const float *new_xyz,
                          const float *pre, int64_t ldpre, const float *wxyz_bias, const int *idx,
                          const captra_mlp_desc *mlp, const void *packed, float *out, int64_t ldo,
                          int col_off, int impl, captra_stream_t stream);

/* Pointwise MLP on rows.  Row r of the input is the concatenation
 *   [ segA[r*ldA : +ca] , segB[(bcast_rows ? r / bcast_rows : r)*ldB : +cb] ]
 * (the torch.cat of pointnet_utils.py:185,292 and the `repeat` of :281-282 without
 * materialising them).  If group > 0 a max over each `group` consecutive rows follows the last
 * layer (group-all SA, pointnet_utils.py:337-343).  Output row r (or group g) is written at
 * y + r*ldy + col_off. */
int captra_point_mlp(int64_t rows, const float *segA, int64_t ldA, int ca, const float *segB,
                     int64_t ldB, int cb, int bcast_rows, const captra_mlp_desc *mlp,
                     const void *packed, float *y, int64_t ldy, int col_off, int group, int impl,
                     captra_stream_t stream);

/* impl 2 (fp16x3) keeps fp32-class accuracy only while |activation| < 65504; the kernels record any
 * larger value.  Synchronises the device; returns 0 (clean), -1 (an operand saturated: re-run the
 * model with impl 1) or a positive error code.  reset != 0 clears the record. */
int captra_f16_overflow_flag(int reset);

/* GroupNorm folded into the consumer layer (RotationRegressor heads, blocks.py:146-193: conv1d ->
 * GroupNorm(C/2 groups) -> ReLU).  captra_group_norm_affine turns the statistics of a pre-norm
 * activation y [clouds*npts, C] (point-major) into per-(cloud, channel) scale/shift; the next
 * layer then reads relu(y * scale[cloud] + shift[cloud]) through captra_point_mlp_affine (tcgen05
 * path only), so the normalised tensor is never written. */
int captra_group_norm_affine(int clouds, int npts, int c, int channels_per_group, const float *y,
                             int64_t ldy, const float *gamma, const float *beta, float eps,
                             float *scale, float *shift, captra_stream_t stream);
int captra_point_mlp_affine(int64_t rows, const float *x, int64_t ldx, int cin, const float *in_scale,
                            const float *in_shift, int rows_per_cloud, const captra_mlp_desc *mlp,
                            const void *packed, float *y, int64_t ldy, int col_off, int impl,
                            captra_stream_t stream);

/* y[r, :] <- relu(y[r, :] * scale[r / rows_per_cloud, :] + shift[...]) in place: the unfused form of the
 * normalise-on-load above, for cloud sizes that are not a multiple of the 128-row tile. */
int captra_group_norm_relu_rows(int64_t rows, int c, int rows_per_cloud, float *y, int64_t ldy, const float *scale,
                                const float *shift, captra_stream_t stream);

/* The same launch with the GroupNorm statistics of its OUTPUT fused into the epilogue (in_scale / in_shift may be
 * NULL: plain input).  stats [ceil(rows/128)*4][2][cout] receives, per 32-row block, the column sums and sums of
 * squares of y; captra_group_norm_finalize turns the blocks of each cloud into the per-(cloud, channel) affine of
 * the following GroupNorm -- the activation is never re-read for its statistics.  Deterministic (no atomics). */
int captra_point_mlp_gnstats(int64_t rows, const float *x, int64_t ldx, int cin, const float *in_scale,
                             const float *in_shift, int rows_per_cloud, const captra_mlp_desc *mlp,
                             const void *packed, float *y, int64_t ldy, int col_off, float *stats, int impl,
                             captra_stream_t stream);
int captra_group_norm_finalize(int clouds, int npts, int c, int channels_per_group, const float *stats,
                               const float *gamma, const float *beta, float eps, float *scale, float *shift,
                               captra_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * 4. Pose fit (pose_utils/procrustes.py, pose_utils/pose_fit.py) -- Python in the reference,
 *    with torch.svd on the CPU (procrustes.py:27-30,170-174); here on the device.
 * ---------------------------------------------------------------------------------------- */

/* R = U diag(1,1,det(U V^T)) V^T for count 3x3 matrices M (row-major) -- procrustes.py:30-54 */
int captra_procrustes_rot3(int64_t count, const float *M, float *R, captra_stream_t stream);
/* 2x2 variant with the orthogonality check / identity fallback -- procrustes.py:174-204 */
int captra_procrustes_rot2(int64_t count, const float *M, float *R, captra_stream_t stream);

/* Fused part_fit_st_no_ransac (pose_fit.py:38-53 -> procrustes.py:132-164):
 *   mask[b,p,i] = (labels[b,i] == p) (labels int64 [B,N]), or -- when labels is NULL -- the
 *   binary float mask [B,P,N] given in `mask` (exactly one of the two must be non-NULL);
 *   valid = sum(mask) > 3 and outputs finite
 *   (optional sym refinement R <- R*Ry via the 2-D fit, procrustes.py:147-151,213-228)
 *   scale       = sum w (R s_c).t_c / (sum w |R s_c|^2 + 1e-6)     (or given_scale)
 *   translation = sum w (t - scale R s) / max(sum w, 1)
 * source/target are addressed through element strides (in floats) so the reference's
 * transposed views ([B,P,3,N].transpose(-1,-2), networks.py:227) need no copy:
 *   src(b,p,i,c) = source[b*ssb + p*ssp + i*ssn + c*ssc].
 * rotation [B,P,3,3] row-major or NULL (-> 3x3 Procrustes, procrustes.py:142-145).
 * Outputs: scale [B,P], translation [B,P,3], valid [B,P] (uint8), rot_out [B,P,3,3] (the
 * rotation actually used for s,t -- may be NULL). */
int captra_part_fit_st(int b, int p, int n, const int64_t *labels, const float *mask,
                       const float *source, int64_t ssb, int64_t ssp, int64_t ssn, int64_t ssc,
                       const float *target, int64_t tsb, int64_t tsp, int64_t tsn, int64_t tsc,
                       const float *rotation, const float *given_scale, int sym, float *scale,
                       float *translation, uint8_t *valid, float *rot_out,
                       captra_stream_t stream);

/* The tracker's call of the fit (networks.py:218-232) as one launch: source = pred_nocs [B,P,3,N], target =
 * points [B,3,N] + points_mean [B,3] for every part, labels [B,N] int64, rotation [B,P,3,3] (given), then
 *   scale <- valid * scale + (1 - valid) * prev_scale,  translation likewise (fp32, as written there). */
int captra_part_fit_track(int b, int p, int n, const int64_t *labels, const float *nocs, const float *points,
                          const float *points_mean, const float *rotation, int sym, const float *prev_scale,
                          const float *prev_translation, float *scale, float *translation, uint8_t *valid,
                          captra_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * 4b. Per-frame glue between the networks and the fit (Python / ~150 small torch kernels per frame in the
 *     reference), one launch each.
 * ---------------------------------------------------------------------------------------- */

/* networks.py:38-41 and :170-187: copy p of cloud b canonicalised by part p's pose,
 *   out = R^T ((points + mean) - t) / s,   points [B,3,N], points_mean [B,3], rotation [B*P,3,3],
 *   translation [B*P,3], scale [B*P].  Any of the three outputs may be NULL: out_pm [B*P,N,3] (point-major, what
 *   the sampling / grouping kernels read), out_cm [B*P,3,N] (the reference's layout, pred_dict['points']),
 *   out_dup [B*P,N,6] = [xyz, xyz] (the l0 skip of backbones.py:67 with use_xyz_feat). */
int captra_canonicalize(int b, int p, int n, const float *points, const float *points_mean,
                        const float *rotation, const float *translation, const float *scale,
                        float *out_pm, float *out_cm, float *out_dup, captra_stream_t stream);

/* networks.py:44-46 + model.py:458 on the heads' raw point-major outputs seg_raw [B*N, ld_seg] (nseg <= 8 classes)
 * and nocs_raw [B*N, ld_nocs] (nnocs = 3P channels): labels [B,N] int64 = argmax of softmax (first maximum),
 * nocs [B,nnocs,N] = sigmoid - 0.5, seg [B,nseg,N] = softmax (may be NULL). */
int captra_coord_head_post(int b, int n, int nseg, int nnocs, const float *seg_raw, int64_t ld_seg,
                           const float *nocs_raw, int64_t ld_nocs, int64_t *labels, float *nocs, float *seg,
                           captra_stream_t stream);

/* blocks.py:181-193 + networks.py:127-141 (diagonal of :200-203) + part_dof_utils.py:124-141: per part p the raw
 * output of head p on copy p, raw_host_ptrs[p] -> [B*N, ld] point-major (3 channels sym / 6), is turned per point into
 * a unit vector / 3x3 matrix, averaged over the points with labels == p (default (0,1,0) / identity when there are
 * none), converted to a rotation (y-axis frame / Gram-Schmidt) and composed: rotation[b,p] = rot_prev[b,p] . dR.
 * rtvec [B,P,3 or 9] (the averaged prediction) may be NULL. */
int captra_rot_head_post(int b, int p, int n, int sym, const float *const *raw_host_ptrs, int64_t ld,
                         const int64_t *labels, const float *rot_prev, float *rotation, float *rtvec,
                         captra_stream_t stream);

/* Per-frame eval / loss reductions of the tracking loop (model.py:511-561), as sums so that ranks can be
 * all-reduced:  sums[5p + q], q = 0..4: sum over clouds of sdiff, tdiff, rdiff (degrees; y axis only when sym),
 * 5deg5cm, 10deg10cm of part p (part_dof_utils.py:40-67, metrics.py:5-47); then sums[5P + 0..4] = number of clouds,
 * sum over (cloud, class) of mIoU and their number (loss.py:122-134), sum of the l2 NOCS errors over the points the
 * loss counts and their number (loss.py:42-70).  gt_labels / gt_nocs [B,N] int64 / [B,3,N] may be NULL (those
 * sums are then 0); seg [B,nseg,N], nocs [B,3P,N], pred_labels [B,N] are the CoordNet predictions.
 * per_instance [B,P,5] may be NULL; scratch: B*3 floats.  accumulate != 0 adds this frame's sums to what `sums`
 * holds (the per-batch accumulation of model.py:511-582 over the frames of a trajectory batch). */
int captra_track_eval(int b, int p, int n, int nseg, int sym, const float *gt_rotation, const float *gt_translation,
                      const float *gt_scale, const float *rotation, const float *translation, const float *scale,
                      const float *seg, const float *nocs, const int64_t *pred_labels, const int64_t *gt_labels,
                      const float *gt_nocs, float *scratch, float *per_instance, float *sums, int accumulate,
                      captra_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * 4c. Data-side crop of a frame on the device (datasets/nocs_data/nocs_data_process.py:92-109,148-164,
 *     nocs_utils.py:5-33, data_utils.py:138-158): numpy on the host in the reference.
 * ---------------------------------------------------------------------------------------- */

/* Back-project the depth window [r0..r1] x [c0..c1] (window_host = {r0, c0, r1, c1}, inclusive; depth [H,W] fp32 in
 * millimetres, 0 = no measurement; mask [H,W] int32 or NULL), test every point against the ball around center_host
 * (3 doubles) and select, in row-major pixel order, the points of the first radius max(radius, .05) * 1.10^i, i < 10,
 * that holds at least ten of them (grow != 0; only i = 0 otherwise) -- or every back-projected point when even the
 * last radius is empty.  kinv_host: the 3x3 inverse intrinsics (row-major doubles).
 * Outputs: meta[0] = n, meta[1] = selected level threshold, meta[2] = index of the radius used; pts [n,3] fp64,
 * pmask [n], pix [n] (row * W + col).  Scratch: level [window pixels] bytes, hist [rows * 11], row_off [rows].
 * pts / pmask / pix must hold one entry per window pixel (n is only known on the device). */
int captra_crop_select(int h, int w, const float *depth, const int *mask, const int *window_host,
                       const double *kinv_host, const double *center_host, double radius, int grow,
                       uint8_t *level, int *hist, int *row_off, int *meta, double *pts, int *pmask, int *pix,
                       captra_stream_t stream);
/* cloud[j] = float(pts[(sel ? sel[j] : j) % n]), j < count: the tiled (nocs_data_process.py:105-106) and, with
 * sel = a permutation prefix, randomly thinned (data_utils.py:147-149) list FPS then runs on. */
int captra_crop_subset(int n, int count, const int64_t *sel, const double *pts, float *cloud, captra_stream_t stream);
/* out[j] = selected entry (sel ? sel[q] : q) % n with q = fps_idx ? fps_idx[j] : j -- points (fp64), mask values,
 * index into the selected list, pixel index. */
int captra_crop_gather(int n, int count, const int *fps_idx, const int64_t *sel, const double *pts, const int *pmask,
                       const int *pix, double *out_pts, int *out_mask, int64_t *out_idx, int *out_pix,
                       captra_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * 5. Unit-test doorway for the tcgen05 primitives: D[128,n] = A[128,k] * W[n,k]^T on one CTA
 *    (terms = 1: single-pass TF32, 3: 3xTF32), and the clock64 phase stamps the fused kernel
 *    records when CAPTRA_TC_DBG has bit 32/128 set.  Not used by the product path.
 * ---------------------------------------------------------------------------------------- */
int captra_debug_umma_gemm(int k, int n, const float *A, const float *W, float *D, int terms,
                           captra_stream_t stream);
int captra_debug_tc_timestamps(long long *out_host, int max_pairs);

#ifdef __cplusplus
}
#endif
#endif /* CAPTRA_OPS_H_ */
