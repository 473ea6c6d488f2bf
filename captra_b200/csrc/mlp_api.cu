// mlp_api.cu -- C ABI of the fused per-point MLP kernels; dispatches on `impl`
// (0: fp32 CUDA cores, mlp_simt.cu; 1: tcgen05 3xTF32, mlp_tc.cu).
#include "mlp_common.cuh"

namespace captra {
int simt_pack(const captra_mlp_desc *d, void *packed, cudaStream_t stream);
int simt_sa_mlp_max(int b, int n, int s, int k, int cfeat, const float *xyz, const float *new_xyz,
                    const float *feats, const int *idx, const captra_mlp_desc *d, const void *packed,
                    float *out, int64_t ldo, int col_off, cudaStream_t stream);
int simt_point_mlp(int64_t rows, const float *segA, int64_t ldA, int ca, const float *segB, int64_t ldB,
                   int cb, int bcast, const captra_mlp_desc *d, const void *packed, float *y, int64_t ldy,
                   int col_off, int group, cudaStream_t stream);
int64_t tc_pack_bytes(const captra_mlp_desc *d, bool f16);
int tc_pack(const captra_mlp_desc *d, void *packed, bool f16, cudaStream_t stream);
int tc_sa_mlp_max(int b, int n, int s, int k, int cfeat, const float *xyz, const float *new_xyz,
                  const float *feats, const int *idx, const captra_mlp_desc *d, const void *packed,
                  float *out, int64_t ldo, int col_off, bool f16, cudaStream_t stream);
int tc_sa_mlp_max_pre(int b, int n, int s, int k, int cpre, const float *xyz, const float *new_xyz, const float *pre,
                      int64_t ldpre, const float *tab, const int *idx, const captra_mlp_desc *d, const void *packed,
                      float *out, int64_t ldo, int col_off, bool f16, cudaStream_t stream);
int tc_point_mlp(int64_t rows, const float *segA, int64_t ldA, int ca, const float *segB, int64_t ldB, int cb,
                 int bcast, const captra_mlp_desc *d, const void *packed, float *y, int64_t ldy, int col_off,
                 int group, bool f16, cudaStream_t stream);
int tc_point_mlp_affine(int64_t rows, const float *x, int64_t ldx, int cin, const float *in_scale, const float *in_shift,
                        int rows_per_cloud, const captra_mlp_desc *d, const void *packed, float *y, int64_t ldy,
                        int col_off, bool f16, cudaStream_t stream);
int tc_point_mlp_gnstats(int64_t rows, const float *x, int64_t ldx, int cin, const float *in_scale, const float *in_shift,
                         int rows_per_cloud, const captra_mlp_desc *d, const void *packed, float *y, int64_t ldy,
                         int col_off, float *stats, bool f16, cudaStream_t stream);
}  // namespace captra

using namespace captra;

extern "C" int64_t captra_mlp_pack_bytes(const captra_mlp_desc *d, int impl) {
    if (check_desc(d, "mlp_pack_bytes") != CAPTRA_OK) return -1;
    if (impl == 0) return (int64_t)(simt_layout(*d).total_floats * sizeof(float));
    if (impl == 1 || impl == 2) {
        const int64_t n = tc_pack_bytes(d, impl == 2);
        if (n < 0) set_error("mlp_pack_bytes: these layer widths are not covered by the tcgen05 path (use impl 0)");
        return n;
    }
    set_error("mlp_pack_bytes: impl %d not available", impl);
    return -1;
}

extern "C" int captra_mlp_pack(const captra_mlp_desc *d, int impl, void *packed, captra_stream_t stream) {
    int rc = check_desc(d, "mlp_pack");
    if (rc) return rc;
    CAPTRA_REQUIRE(packed && (reinterpret_cast<uintptr_t>(packed) & 15) == 0, "mlp_pack: packed buffer must be 16-byte aligned");
    for (int l = 0; l < d->nlayers; ++l) CAPTRA_REQUIRE(d->w[l], "mlp_pack: null weights for layer %d", l);
    if (impl == 0) return simt_pack(d, packed, as_stream(stream));
    if (impl == 1 || impl == 2) return tc_pack(d, packed, impl == 2, as_stream(stream));
    set_error("mlp_pack: impl %d not available", impl);
    return CAPTRA_ERR_UNSUPPORTED;
}

extern "C" int captra_sa_mlp_max(int b, int n, int s, int k, int cfeat, const float *xyz,
                                 const float *new_xyz, const float *feats, const int *idx,
                                 const captra_mlp_desc *d, const void *packed, float *out,
                                 int64_t ldo, int col_off, int impl, captra_stream_t stream) {
    int rc = check_desc(d, "sa_mlp_max");
    if (rc) return rc;
    CAPTRA_REQUIRE(b >= 0 && n >= 1 && s >= 0 && k >= 1 && cfeat >= 0, "sa_mlp_max: bad sizes");
    CAPTRA_REQUIRE(d->cin == cfeat + 3, "sa_mlp_max: mlp cin=%d but cfeat+3=%d", d->cin, cfeat + 3);
    if (b == 0 || s == 0) return CAPTRA_OK;
    CAPTRA_REQUIRE(xyz && new_xyz && idx && packed && out && (feats || cfeat == 0), "sa_mlp_max: null pointer");
    CAPTRA_REQUIRE((int64_t)b * n < (1LL << 31), "sa_mlp_max: b*n overflows int");
    if (impl == 0)
        return simt_sa_mlp_max(b, n, s, k, cfeat, xyz, new_xyz, feats, idx, d, packed, out, ldo, col_off, as_stream(stream));
    if (impl == 1 || impl == 2)
        return tc_sa_mlp_max(b, n, s, k, cfeat, xyz, new_xyz, feats, idx, d, packed, out, ldo, col_off, impl == 2, as_stream(stream));
    set_error("sa_mlp_max: impl %d not available", impl);
    return CAPTRA_ERR_UNSUPPORTED;
}

extern "C" int captra_sa_mlp_max_pre(int b, int n, int s, int k, int cpre, const float *xyz, const float *new_xyz,
                                     const float *pre, int64_t ldpre, const float *wxyz_bias, const int *idx,
                                     const captra_mlp_desc *d, const void *packed, float *out, int64_t ldo,
                                     int col_off, int impl, captra_stream_t stream) {
    int rc = check_desc(d, "sa_mlp_max_pre");
    if (rc) return rc;
    CAPTRA_REQUIRE(b >= 0 && n >= 1 && s >= 0 && k >= 1 && cpre >= 1 && ldpre >= cpre, "sa_mlp_max_pre: bad sizes");
    CAPTRA_REQUIRE(d->cin == cpre, "sa_mlp_max_pre: tail mlp cin=%d but the projected rows have %d channels", d->cin, cpre);
    CAPTRA_REQUIRE(impl == 1 || impl == 2, "sa_mlp_max_pre: only the tcgen05 paths (impl 1, 2) implement the projected layer 0");
    if (b == 0 || s == 0) return CAPTRA_OK;
    CAPTRA_REQUIRE(xyz && new_xyz && idx && packed && out && pre && wxyz_bias, "sa_mlp_max_pre: null pointer");
    return tc_sa_mlp_max_pre(b, n, s, k, cpre, xyz, new_xyz, pre, ldpre, wxyz_bias, idx, d, packed, out, ldo, col_off, impl == 2,
                             as_stream(stream));
}

extern "C" int captra_point_mlp(int64_t rows, const float *segA, int64_t ldA, int ca, const float *segB,
                                int64_t ldB, int cb, int bcast_rows, const captra_mlp_desc *d,
                                const void *packed, float *y, int64_t ldy, int col_off, int group,
                                int impl, captra_stream_t stream) {
    int rc = check_desc(d, "point_mlp");
    if (rc) return rc;
    CAPTRA_REQUIRE(rows >= 0 && ca >= 0 && cb >= 0 && group >= 0 && bcast_rows >= 0, "point_mlp: bad sizes");
    CAPTRA_REQUIRE(d->cin == ca + cb, "point_mlp: mlp cin=%d but ca+cb=%d", d->cin, ca + cb);
    CAPTRA_REQUIRE(group == 0 || rows % group == 0, "point_mlp: rows %lld not a multiple of group %d", (long long)rows, group);
    if (rows == 0) return CAPTRA_OK;
    CAPTRA_REQUIRE(packed && y && (segA || ca == 0) && (segB || cb == 0), "point_mlp: null pointer");
    if (impl == 0)
        return simt_point_mlp(rows, segA, ldA, ca, segB, ldB, cb, bcast_rows, d, packed, y, ldy, col_off, group, as_stream(stream));
    if (impl == 1 || impl == 2)
        return tc_point_mlp(rows, segA, ldA, ca, segB, ldB, cb, bcast_rows, d, packed, y, ldy, col_off, group, impl == 2, as_stream(stream));
    set_error("point_mlp: impl %d not available", impl);
    return CAPTRA_ERR_UNSUPPORTED;
}

extern "C" int captra_point_mlp_affine(int64_t rows, const float *x, int64_t ldx, int cin, const float *in_scale,
                                       const float *in_shift, int rows_per_cloud, const captra_mlp_desc *d,
                                       const void *packed, float *y, int64_t ldy, int col_off, int impl,
                                       captra_stream_t stream) {
    int rc = check_desc(d, "point_mlp_affine");
    if (rc) return rc;
    CAPTRA_REQUIRE(rows >= 0 && cin >= 1 && rows_per_cloud >= 1, "point_mlp_affine: bad sizes");
    CAPTRA_REQUIRE(d->cin == cin, "point_mlp_affine: mlp cin=%d but input has %d channels", d->cin, cin);
    CAPTRA_REQUIRE(impl == 1 || impl == 2, "point_mlp_affine: only the tcgen05 paths (impl 1, 2) implement normalise-on-load");
    if (rows == 0) return CAPTRA_OK;
    CAPTRA_REQUIRE(x && in_scale && in_shift && packed && y, "point_mlp_affine: null pointer");
    return tc_point_mlp_affine(rows, x, ldx, cin, in_scale, in_shift, rows_per_cloud, d, packed, y, ldy, col_off, impl == 2, as_stream(stream));
}

extern "C" int captra_point_mlp_gnstats(int64_t rows, const float *x, int64_t ldx, int cin, const float *in_scale,
                                        const float *in_shift, int rows_per_cloud, const captra_mlp_desc *d,
                                        const void *packed, float *y, int64_t ldy, int col_off, float *stats, int impl,
                                        captra_stream_t stream) {
    int rc = check_desc(d, "point_mlp_gnstats");
    if (rc) return rc;
    CAPTRA_REQUIRE(rows >= 0 && cin >= 1 && rows_per_cloud >= 1, "point_mlp_gnstats: bad sizes");
    CAPTRA_REQUIRE(d->cin == cin, "point_mlp_gnstats: mlp cin=%d but input has %d channels", d->cin, cin);
    CAPTRA_REQUIRE(impl == 1 || impl == 2, "point_mlp_gnstats: only the tcgen05 paths (impl 1, 2) fuse the statistics");
    CAPTRA_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), "point_mlp_gnstats: scale and shift come together");
    CAPTRA_REQUIRE(rows_per_cloud % 128 == 0, "point_mlp_gnstats: rows_per_cloud must be a multiple of 128 (got %d)", rows_per_cloud);
    if (rows == 0) return CAPTRA_OK;
    CAPTRA_REQUIRE(x && packed && y && stats, "point_mlp_gnstats: null pointer");
    return tc_point_mlp_gnstats(rows, x, ldx, cin, in_scale, in_shift, rows_per_cloud, d, packed, y, ldy, col_off, stats, impl == 2,
                                as_stream(stream));
}
