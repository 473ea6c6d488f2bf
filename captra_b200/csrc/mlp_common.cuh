// mlp_common.cuh -- shared host-side description of the fused per-point MLP kernels.
#pragma once
#include "common.cuh"

namespace captra {

constexpr int MLP_KC = 16;   // K-chunk (input channels per shared-memory stage), SIMT path
constexpr int MLP_TR = 64;   // rows per tile, SIMT path

inline int round_up(int a, int b) { return (a + b - 1) / b * b; }

// Layout of the packed-weights buffer for impl 0 (fp32 CUDA cores):
//   per layer l: Wt_l [cin_pad_l][coutp_l] (K-major rows, zero padded), then bias_l [coutp_l]
struct SimtLayout {
    int nlayers;
    int cin[CAPTRA_MAX_MLP_LAYERS], cin_pad[CAPTRA_MAX_MLP_LAYERS];
    int cout[CAPTRA_MAX_MLP_LAYERS], coutp[CAPTRA_MAX_MLP_LAYERS], cw[CAPTRA_MAX_MLP_LAYERS];
    size_t off_w[CAPTRA_MAX_MLP_LAYERS], off_b[CAPTRA_MAX_MLP_LAYERS];  // in floats
    size_t total_floats;
};

inline SimtLayout simt_layout(const captra_mlp_desc &d) {
    SimtLayout L{};
    L.nlayers = d.nlayers;
    size_t off = 0;
    int cin = d.cin;
    for (int l = 0; l < d.nlayers; ++l) {
        L.cin[l] = cin;
        L.cin_pad[l] = round_up(cin, MLP_KC);
        L.cout[l] = d.cout[l];
        L.cw[l] = d.cout[l] > 64 ? 128 : 64;          // column-pass width
        L.coutp[l] = round_up(d.cout[l], L.cw[l]);
        L.off_w[l] = off; off += (size_t)L.cin_pad[l] * L.coutp[l];
        L.off_b[l] = off; off += L.coutp[l];
        cin = d.cout[l];
    }
    L.total_floats = off;
    return L;
}

inline int check_desc(const captra_mlp_desc *d, const char *who) {
    CAPTRA_REQUIRE(d, "%s: null mlp desc", who);
    CAPTRA_REQUIRE(d->nlayers >= 1 && d->nlayers <= CAPTRA_MAX_MLP_LAYERS, "%s: nlayers=%d", who, d->nlayers);
    CAPTRA_REQUIRE(d->cin >= 1, "%s: cin=%d", who, d->cin);
    for (int l = 0; l < d->nlayers; ++l) CAPTRA_REQUIRE(d->cout[l] >= 1, "%s: cout[%d]=%d", who, l, d->cout[l]);
    return CAPTRA_OK;
}

}  // namespace captra
