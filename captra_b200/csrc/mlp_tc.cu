// mlp_tc.cu -- fused shared per-point MLP on the 5th-gen tensor cores (impl 1, 2): tcgen05.mma with
// TMEM accumulators, operands staged in shared memory (weights by bulk-copy TMA), 3-term split
// arithmetic (3xTF32 or fp16x3) so the result matches true fp32 to ~1e-6 (the reference's convs are true fp32;
// single-pass TF32 misses the 1e-4 pose tolerance, SURVEY section 7 "hard parts").
#include "mlp_common.cuh"
#include "tc_common.cuh"

#include <math.h>
#include <stdlib.h>
#include <type_traits>

namespace captra {
using namespace tc;

constexpr int TC_ROWS = 128;                 // UMMA M (cta_group::1): one TMEM lane per row
constexpr int TC_CHUNK_BYTES = TC_ROWS * 16; // one 16-byte K chunk of all 128 rows (A operand)

// ------------------------------------------------------------------------------------------------
// Debug / unit-test entry: D[128,N] = A[128,K] * W[N,K]^T with 3xTF32, one CTA.  Exercises the
// descriptor encodings, the chunk-major no-swizzle operand layout, TMEM alloc/ld and the
// commit/mbarrier handshake in isolation (tests/test_tc_gpu.py).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) umma_debug_gemm_kernel(int K, int N, const float *__restrict__ A,
                                                               const float *__restrict__ W, float *__restrict__ D,
                                                               int terms) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t mma_done;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nchunk = K / 4;
    float *a_hi = reinterpret_cast<float *>(smem_raw);
    float *a_lo = a_hi + (size_t)nchunk * TC_ROWS * 4;
    float *b_hi = a_lo + (size_t)nchunk * TC_ROWS * 4;
    float *b_lo = b_hi + (size_t)nchunk * N * 4;

    if (warp == 0) tmem_alloc<256>(&tmem_base_s);
    if (tid == 0) {
        mbar_init(&mma_done, 1);
        fence_mbar_init();
    }
    // operands -> smem, chunk-major: element (r,k) at chunk (k/4): [chunk][row][4]
    for (int i = tid; i < TC_ROWS * K; i += 128) {
        const int r = i / K, k = i - r * K;
        float hi, lo;
        split_tf32(A[i], hi, lo);
        const size_t o = ((size_t)(k >> 2) * TC_ROWS + r) * 4 + (k & 3);
        a_hi[o] = hi; a_lo[o] = lo;
    }
    for (int i = tid; i < N * K; i += 128) {
        const int n = i / K, k = i - n * K;
        float hi, lo;
        split_tf32(W[i], hi, lo);
        const size_t o = ((size_t)(k >> 2) * N + n) * 4 + (k & 3);
        b_hi[o] = hi; b_lo[o] = lo;
    }
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_d = tmem_base_s;

    if (tid == 0) {
        const uint32_t idesc = make_idesc(2, TC_ROWS, N);
        const uint32_t a_lbo = TC_CHUNK_BYTES, b_lbo = (uint32_t)N * 16, sbo = 128;
        uint32_t acc = 0;
        for (int s = 0; s < K / 8; ++s) {
            const uint32_t ao = (uint32_t)(2 * s) * a_lbo, bo = (uint32_t)(2 * s) * b_lbo;
            const uint64_t ah = smem_desc_kmajor_noswz(smem_u32(a_hi) + ao, a_lbo, sbo);
            const uint64_t al = smem_desc_kmajor_noswz(smem_u32(a_lo) + ao, a_lbo, sbo);
            const uint64_t bh = smem_desc_kmajor_noswz(smem_u32(b_hi) + bo, b_lbo, sbo);
            const uint64_t bl = smem_desc_kmajor_noswz(smem_u32(b_lo) + bo, b_lbo, sbo);
            if (terms >= 3) {  // small terms first
                umma_tf32(tmem_d, al, bh, idesc, acc); acc = 1;
                umma_tf32(tmem_d, ah, bl, idesc, acc);
            }
            umma_tf32(tmem_d, ah, bh, idesc, acc); acc = 1;
        }
        umma_commit(&mma_done);
    }
    mbar_wait(&mma_done, 0);
    tcgen05_fence_after();
    const int row = warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 16) {
        uint32_t v[16];
        tmem_ld_32x16(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem_d);
}


// ------------------------------------------------------------------------------------------------
// The fused kernel.
//
// Persistent CTAs walk 128-row tiles.  A K slab is KC = 16*G input channels; for layers > 0 producer
// group g (128 threads: thread r <-> tile row r <-> TMEM lane r) owns the 16-channel PIECE g of EVERY
// slab, so all 4G producer warps work on the same slab at once: the per-slab latency is one piece, the
// pipeline restarts quickly at a layer boundary, and 16 (LARGE) / 2x8 (SMALL, two CTAs per SM) warps
// give the SM enough independent instruction streams to hide the TMEM / L2 / conversion latencies.
//
//   layer l, K slab s, stage st = slab counter mod NST:
//     producers   : the A slab -> (hi, lo) planes of a_stage[st]
//         layer 0 : rows dealt out per WARP for coalesced reads (8 rows x 128 contiguous bytes per LDG.256
//                   instruction), gathered through the metadata ring (SA) or row pointers (dense), one slab
//                   ahead in registers (two at 168 registers); optional on-load transforms: GroupNorm
//                   affine + ReLU of the producer layer (heads), or the projected layer 0 of an SA scale
//                   (relu(P[j] + W_x (x_j - c) + b), tc_sa_mlp_max_pre)
//         layer>0 : read straight out of TMEM -- 16 accumulator columns of layer l-1 ARE piece g of
//                   slab s of layer l -- + bias, ReLU, split.  The epilogue of layer l-1 and the MMAs
//                   of layer l overlap slab by slab; accumulators ping-pong between two TMEM regions
//                   and no activation ever touches shared memory in fp32.
//     TMA warp    : W slab (hi plane | lo plane, pre-split, chunk-major in global) -> w_stage[st]; for SA
//                   launches it also resolves the NEXT tile's gather metadata (index -> point ->
//                   coordinates) into a two-tile shared-memory ring while it waits for free stages
//     MMA warp    : k-steps x 3 terms (lo*hi, hi*lo, hi*hi) of tcgen05.mma, tcgen05.commit ->
//                   empty[st]; after a layer's last slab also -> d_ready.  K-steps (and pieces) that
//                   only carry zero padding are skipped in every layer.
//   last epilogue : grouped (GRP) -> the last layer is computed TRANSPOSED (channels in TMEM lanes), so the
//                                max over a group's rows is a per-thread reduction of the RAW accumulators
//                                (bias and ReLU commute with the max), one coalesced row write per group
//                   rows        -> the warp's 32 x 16 block is transposed through the idle A stages and
//                                written as 8 rows x 64 contiguous bytes per store; optionally the column
//                                sums / sums of squares of the block (GroupNorm statistics of the output)
//
// The template parameters select exactly the code a launch runs (loader, operand type, CTA shape, stages,
// ring-only layer 0, epilogue kind, GroupNorm-head extras): at 96 registers per thread every dead branch
// costs spills, and with the L1 carved out for shared memory a spill is an L2 round trip.
//
// Shared-memory operand layout (K-major, no swizzle): element (row, k) of a slab lives at
//   plane + (k / EPC) * ROWS*16 + row*16 + (k % EPC) * sizeof  (EPC = 8 halfs or 4 tf32 per 16-byte
//   chunk)  -> LBO = ROWS*16, SBO = 128 in the descriptors, so thread r writes 16-byte pieces at
//   r*16 (conflict-free) and never touches another row.
//
// fp16x3 conversion cost matters (it is the producers' inner loop): two elements per F2FP via
// cvt.*.f16x2.f32.  For layers > 0 the ReLU is folded into the conversion: hi = cvt.rz.relu (round
// toward zero, so the remainder x - hi is >= 0 for x >= 0 and the .relu on lo = cvt.rn.relu(x - hi)
// both keeps it and zeroes the x < 0 case, where hi = 0 and x - hi = x < 0).
// ------------------------------------------------------------------------------------------------
constexpr int TC_BIAS_PAD = 16;                   // bias region = npad + 16 floats; [npad] = 1 / weight scale
constexpr int TC_MAX_STAGES = 4;
// channels per K slab: 8 chunks of 16 bytes per row (LARGE) or 4 (SMALL); a chunk is 8 halfs / 4 tf32
__host__ __device__ constexpr int tc_kc(bool f16, bool small) { return (f16 ? 64 : 32) / (small ? 2 : 1); }
__host__ __device__ constexpr int tc_groups(bool f16, bool small) { return tc_kc(f16, small) / 16; }
__host__ __device__ constexpr int tc_threads(bool f16, bool small) { return 128 * tc_groups(f16, small) + 64; }

// Registers: the 64 K registers of an SM are split over its four sub-partitions (16 K each) and a CTA's warps are
// dealt out round-robin, so the per-thread budget follows the sub-partition with the most warps: 18 warps (LARGE
// fp16x3) or 2 x 10 (SMALL fp16x3) -> 5 warps -> 96 registers; 10 or 2 x 6 warps (3xTF32) -> 3 warps -> 168.
// __launch_bounds__ makes ptxas apply exactly that (a larger __maxnreg__ fails at launch).
struct TcArgs {
    int nlayers, relu_last, cout_last;
    int kpad[CAPTRA_MAX_MLP_LAYERS], npad[CAPTRA_MAX_MLP_LAYERS];
    int kreal[CAPTRA_MAX_MLP_LAYERS];           // input channels that carry data (cin, then npad[l-1]); the rest of kpad is skipped
    const float *wpk[CAPTRA_MAX_MLP_LAYERS];   // [nslab][2 planes][NCHUNK][npad][16 bytes]
    const float *bias[CAPTRA_MAX_MLP_LAYERS];  // [npad + 16]
    int64_t rows, ntiles;
    int group;                                  // max over each `group` rows (32|64|128) or 0
    float *out; int64_t ldo; int col_off;
    int n, s, cfeat; const float *xyz, *new_xyz, *feats; const int *idx;
    int64_t ldf;                                // SA: row stride of feats (cfeat, or the width of the projected buffer)
    const float *pre_tab;                       // SA, projected layer 0 (see tc_sa_mlp_max_pre): [4][cfeat] = wx, wy, wz, bias
    int pre_pad;                                // > 0: that table is staged in shared memory (4 * pre_pad floats)
    const float *segA; int64_t ldA; int ca; const float *segB; int64_t ldB; int cb; int bcast;
    const float *in_scale, *in_shift; int rows_per_cloud;   // dense: x <- relu(x * scale[cloud] + shift[cloud]) on load (GroupNorm + ReLU of the producer layer)
    int aff_pad;                                // > 0: the tile's scale/shift rows are staged in shared memory (2 * aff_pad floats)
    float *stats;                               // dense rows: per 32-row block column sums and sums of squares [blocks][2][cout_total] (GroupNorm statistics of the OUTPUT), or null
    int wstage_bytes, bias_floats, region_cols, tmem_cols, nst_log2;
    int nsplit, last_npad, cout_total;          // single wide layer split into 256-column chunks over grid.y
    int f16, small;
    int red2;                                   // the grouped-max buffer is double-buffered by tile parity (no barrier after its read-out)
    int cin0;                                   // true layer-0 width
    int tpose;                                  // grouped max: the last layer is computed transposed (channels in TMEM lanes)
    int dbg;                                    // timing probes (CAPTRA_TC_DBG): 1 no A production, 2 no W copies, 4 no MMAs, 8 no proxy fence, 64 no layer-0 gather, 128 no TMEM reads in the producers, 32 phase stamps
};

__device__ __forceinline__ void tmem_alloc_dyn(uint32_t *smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_dyn(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// fp16 operands saturate at +-65504.  A value that large is recorded in g_f16_overflow so the host can
// detect the (never observed) case and re-run on the 3xTF32 path (captra_f16_overflow_flag).
__device__ int g_f16_overflow;
// x -> (hi, lo) as two fp16 numbers (hi = rn(x), lo = rn(x - hi)): 22 mantissa bits together (weights)
__device__ __forceinline__ void split_f16(float x, unsigned short &hi, unsigned short &lo) {
    unsigned short h, l;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
    float hf;
    asm("cvt.f32.f16 %0, %1;" : "=f"(hf) : "h"(h));
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(l) : "f"(x - hf));
    hi = h; lo = l;
}
// Two activations -> packed (hi, lo) f16x2 words (element 0 in the low half).  RELU = false: hi = rn(x),
// lo = rn(x - hi) (signed inputs of layer 0).  RELU = true: the split of max(x, 0) with the ReLU folded
// into the conversions (see the header comment).  x - hi is exact in fp32 either way.
template <bool RELU>
__device__ __forceinline__ void split2_f16(float x0, float x1, uint32_t &hi2, uint32_t &lo2) {
    if (RELU) asm("cvt.rz.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi2) : "f"(x1), "f"(x0));
    else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi2) : "f"(x1), "f"(x0));
    float h0, h1;
    asm("{\n\t.reg .f16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(h0), "=f"(h1) : "r"(hi2));
    const float d0 = __fsub_rn(x0, h0), d1 = __fsub_rn(x1, h1);
    if (RELU) asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo2) : "f"(d1), "f"(d0));
    else asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo2) : "f"(d1), "f"(d0));
}
// Max over the 32 lanes of a warp for 16 values per lane in 16 shuffles (recursive halving): after
// the four halving rounds lane l holds column ((l&1)<<3 | (l&2)<<1 | (l&4)>>1 | (l&8)>>3); a final
// xor-16 round merges the two half-warps.  (redux.sync on the bit patterns is one instruction per
// value but serialises at ~60 cycles each through the uniform datapath.)
__device__ __forceinline__ float warp_colmax16(const float (&x)[16], int lane, int &col) {
    float y[8], z[4], w[2];
    const bool b0 = lane & 1, b1 = lane & 2, b2 = lane & 4, b3 = lane & 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float keep = b0 ? x[j + 8] : x[j], send = b0 ? x[j] : x[j + 8];
        y[j] = fmaxf(keep, __shfl_xor_sync(kFull, send, 1));
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float keep = b1 ? y[j + 4] : y[j], send = b1 ? y[j] : y[j + 4];
        z[j] = fmaxf(keep, __shfl_xor_sync(kFull, send, 2));
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const float keep = b2 ? z[j + 2] : z[j], send = b2 ? z[j] : z[j + 2];
        w[j] = fmaxf(keep, __shfl_xor_sync(kFull, send, 4));
    }
    const float keep = b3 ? w[1] : w[0], send = b3 ? w[0] : w[1];
    float v = fmaxf(keep, __shfl_xor_sync(kFull, send, 8));
    v = fmaxf(v, __shfl_xor_sync(kFull, v, 16));
    col = (b0 ? 8 : 0) | (b1 ? 4 : 0) | (b2 ? 2 : 0) | (b3 ? 1 : 0);
    return v;
}

// The timing probes (knock-out knobs and phase stamps, CAPTRA_TC_DBG) are compiled in only with -DCAPTRA_TC_PROBES:
//   CAPTRA_EXTRA_NVCC_FLAGS=-DCAPTRA_TC_PROBES CAPTRA_LIB_OUT=captra_b200/libcaptra_ops_probes.so python -m captra_b200.build
//   CAPTRA_LIB_PATH=$PWD/captra_b200/libcaptra_ops_probes.so python scripts/tc_probe.py
// In the product library they are constant-folded away: the run-time tests sat in the per-slab loops of launches
// that are instruction-issue bound.
#ifdef CAPTRA_TC_PROBES
#define TC_DBG (a.dbg)
#else
#define TC_DBG 0
#endif

// timing probe (CAPTRA_TC_DBG bit 32): producer thread 0 of CTA 0 stamps clock64() at phase
// boundaries of its first tiles; read back with captra_debug_tc_timestamps
__device__ long long g_tc_ts[512];
__device__ int g_tc_ts_n;
#define TC_STAMP(code)                                                              \
    do {                                                                            \
        if ((TC_DBG & 32) && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) {       \
            const int i__ = g_tc_ts_n;                                              \
            if (i__ < 255) { g_tc_ts[2 * i__] = clock64(); g_tc_ts[2 * i__ + 1] = (code); g_tc_ts_n = i__ + 1; } \
        }                                                                           \
    } while (0)

// one 16-channel piece of one row, split into the two operand planes
template <bool F16>
struct TcPiece {
    uint32_t hi[F16 ? 8 : 16], lo[F16 ? 8 : 16];
};

// Two shapes of the same kernel.  LARGE (a layer wider than 128 columns): 8-chunk slabs, G = 4 (fp16x3)
// or 2 (3xTF32) producer groups, one CTA per SM.  SMALL (every layer <= 128 columns: the sa1 scales,
// fp1): 4-chunk slabs, G = 2 / 1, 64 KB and 256 TMEM columns per CTA so TWO CTAs share an SM and one
// tile's epilogue overlaps the other's MMAs.
template <int MODE, bool F16, bool SMALL, int NSTL2, bool RING, bool GRP, bool GN>  // GRP: grouped-max epilogue (always for SA), else row output; GN: GroupNorm head launch (affine on load and / or output statistics); MODE 0: SA gather loader, 1: dense-row loader; 2^NSTL2 pipeline stages; RING: SA layer 0 (<= 8 channels) comes entirely from the metadata ring; the last epilogue is chosen by a.group
__global__ void __launch_bounds__(tc_threads(F16, SMALL), SMALL ? 2 : 1) mlp_tc_kernel(const TcArgs a) {
    constexpr int G = tc_groups(F16, SMALL);
    constexpr int PROD = 128 * G;                  // producer / epilogue threads (warps 0 .. 4G-1)
    constexpr int NTHREADS = PROD + 64;            // + MMA warp (4G) + TMA warp (4G+1)
    constexpr int NCHUNK = SMALL ? 4 : 8;
    constexpr int A_PLANE = NCHUNK * TC_CHUNK_BYTES;   // bytes of one (hi or lo) A slab plane
    constexpr int A_STAGE = 2 * A_PLANE;
    constexpr int KC = tc_kc(F16, SMALL);
    constexpr int KMMA = F16 ? 16 : 8;             // K of one tcgen05.mma (32 bytes per row)
    constexpr int KSTEPS = KC / KMMA;              // == NCHUNK / 2
    static_assert(KC == 16 * G && KSTEPS == NCHUNK / 2, "one 16-channel piece per producer group");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t full[TC_MAX_STAGES], empty[TC_MAX_STAGES], d_ready;
    __shared__ uint64_t meta_full[2], meta_empty[2];   // SA: per-tile row metadata ring (written by the TMA warp)
    __shared__ uint32_t tmem_base_s;

    constexpr uint32_t nst_log2 = NSTL2, NST = 1u << NSTL2;   // compile-time: stage addresses and parities fold into immediates
    uint8_t *a_stage = smem_raw;                                          // NST x A_STAGE
    uint8_t *w_stage = a_stage + (size_t)NST * A_STAGE;                   // NST x wstage_bytes
    float *bias_s = reinterpret_cast<float *>(w_stage + (size_t)NST * a.wstage_bytes);
    float *aff_s = bias_s + a.bias_floats;                                // [2][aff_pad] when a.aff_pad > 0
    float4 *meta_s = reinterpret_cast<float4 *>(aff_s + 2 * a.aff_pad + 4 * a.pre_pad);   // [2 tiles][128 rows][2]: {point, dx, dy, dz} or the assembled 8-channel row
    float *red = reinterpret_cast<float *>(meta_s + 2 * TC_ROWS * 2);     // [2 tiles][4][256] partial maxima of the grouped epilogue, double-buffered by tile parity (also the slack the
                                                                          // transposed last layer's 128-row weight reads may run into)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // a single wide layer is split over grid.y: this CTA owns output columns [256*y, 256*y + npad)
    int npad_y = 0, col_off = a.col_off, cout_last = a.cout_last;
    const float *wpk0 = a.wpk[0], *bias0 = a.bias[0];
    if (a.nsplit > 1) {
        const int y = blockIdx.y;
        npad_y = (y < a.nsplit - 1) ? 256 : a.last_npad;
        wpk0 += (size_t)y * (a.kpad[0] / KC) * 2 * NCHUNK * 256 * 4;
        bias0 += y * (256 + TC_BIAS_PAD);
        col_off += y * 256;
        cout_last = min(256, a.cout_total - y * 256);
    }
    auto NPAD = [&](int l) { return a.nsplit > 1 ? npad_y : a.npad[l]; };
    auto WPK = [&](int l) { return l == 0 ? wpk0 : a.wpk[l]; };
    auto BIAS = [&](int l) { return l == 0 ? bias0 : a.bias[l]; };

    if (warp == PROD / 32) tmem_alloc_dyn(&tmem_base_s, (uint32_t)a.tmem_cols);
    if (tid == 0) {
        for (uint32_t i = 0; i < NST; ++i) { mbar_init(&full[i], 4 * G + 1); mbar_init(&empty[i], 1); }
        mbar_init(&d_ready, 1);
        for (int i = 0; i < 2; ++i) { mbar_init(&meta_full[i], 1); mbar_init(&meta_empty[i], 4 * G); }
        fence_mbar_init();
    }
    {   // biases of all layers -> smem
        int off = 0;
        for (int l = 0; l < a.nlayers; ++l) {
            for (int i = tid; i < NPAD(l) + TC_BIAS_PAD; i += NTHREADS) bias_s[off + i] = __ldg(BIAS(l) + i);
            off += NPAD(l) + TC_BIAS_PAD;
        }
    }
    if (a.pre_pad) {   // [wx | wy | wz | bias] of the projected layer 0, zero padded
        for (int i = tid; i < 4 * a.pre_pad; i += NTHREADS) {
            const int t = i / a.pre_pad, c = i - t * a.pre_pad;
            aff_s[i] = c < a.cfeat ? __ldg(a.pre_tab + (size_t)t * a.cfeat + c) : 0.f;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    // Role dispatch on a provably warp-uniform warp index: the MMA and TMA warps run their loops
    // converged and elect one lane only around the issue itself, so descriptors and counters stay in
    // uniform registers.  (With `if (tid == X)` the compiler wrapped every tcgen05.mma in an
    // ELECT / 5x R2UR / BRA.U.ANY waterfall: ~80 cycles per MMA against 64 tensor cycles at N = 128.)
    const int warp_u = __shfl_sync(kFull, warp, 0);

    if (warp_u == PROD / 32) {
        // ===================== MMA issuer =====================
        uint32_t it = 0;
        const uint64_t a_desc0 = smem_desc_kmajor_noswz(smem_u32(a_stage), TC_CHUNK_BYTES, 128);
        for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
            for (int l = 0; l < a.nlayers; ++l) {
                const int nslab = a.kpad[l] / KC;
                const uint32_t npad = (uint32_t)NPAD(l);
                const uint32_t idesc = make_idesc(F16 ? 0 : 2, TC_ROWS, (int)npad);
                const uint32_t idesc_t = make_idesc(F16 ? 0 : 2, TC_ROWS, TC_ROWS);
                const bool tpose_l = a.tpose && l == a.nlayers - 1;
                const uint32_t nblk = (npad + 127u) / 128u;
                const uint32_t b_lbo = npad * 16;
                const uint64_t b_desc0 = smem_desc_kmajor_noswz(smem_u32(w_stage), b_lbo, 128);
                const uint32_t b_lo_off = (NCHUNK * b_lbo) >> 4, b_step = (2 * b_lbo) >> 4;   // in 16-byte units
                const uint32_t tmem_d = tmem_base + (uint32_t)((l & 1) * a.region_cols);
                int krem = a.kreal[l];                                  // channels with data left in this layer
                for (int s = 0; s < nslab; ++s, ++it, krem -= KC) {
                    const uint32_t st = it & (NST - 1), ph = (it >> nst_log2) & 1;
                    mbar_wait_warp(&full[st], ph);
                    tcgen05_fence_after();
                    // descriptors of this stage: only the 14-bit start-address field (16-byte units) moves
                    uint64_t ah = a_desc0 + (uint64_t)((st * A_STAGE) >> 4);
                    uint64_t al = ah + (A_PLANE >> 4);
                    uint64_t bh = b_desc0 + (uint64_t)((st * (uint32_t)a.wstage_bytes) >> 4);
                    uint64_t bl = bh + b_lo_off;
                    const int nks = min(KSTEPS, (krem + KMMA - 1) / KMMA);   // zero padding needs no MMAs
                    if (elect_one()) {
                        if (tpose_l) {
                            // D^T[channel, point] += W_blk[128 x K] * A[128 points x K]^T: the weights take the A
                            // role (one 128-row block of output channels per MMA, 2048 bytes apart inside a
                            // chunk), the activation slab the B role (N = 128 points).  Lanes = channels, so
                            // the max over a centroid's points is a per-thread reduction in the epilogue.
    #pragma unroll
                            for (int j = 0; j < KSTEPS; ++j) {
                                if (j >= nks || (TC_DBG & 4)) break;
                                for (uint32_t cb = 0; cb < nblk; ++cb) {
                                    const uint32_t dT = tmem_d + cb * 128u;
                                    const uint64_t wo = (uint64_t)(cb * 128u);        // 128 rows * 16 bytes, in 16-byte units
                                    if (F16) {
                                        umma_f16(dT, bh + wo, al, idesc_t, (s | j) ? 1u : 0u);
                                        umma_f16(dT, bl + wo, ah, idesc_t, 1u);
                                        umma_f16(dT, bh + wo, ah, idesc_t, 1u);
                                    } else {
                                        umma_tf32(dT, bh + wo, al, idesc_t, (s | j) ? 1u : 0u);
                                        umma_tf32(dT, bl + wo, ah, idesc_t, 1u);
                                        umma_tf32(dT, bh + wo, ah, idesc_t, 1u);
                                    }
                                }
                                ah += (2 * TC_CHUNK_BYTES) >> 4; al += (2 * TC_CHUNK_BYTES) >> 4;
                                bh += b_step; bl += b_step;
                            }
                        } else {
    #pragma unroll
                            for (int j = 0; j < KSTEPS; ++j) {
                                if (j >= nks || (TC_DBG & 4)) break;
                                if (F16) {
                                    umma_f16(tmem_d, al, bh, idesc, (s | j) ? 1u : 0u);
                                    umma_f16(tmem_d, ah, bl, idesc, 1u);
                                    umma_f16(tmem_d, ah, bh, idesc, 1u);
                                } else {
                                    umma_tf32(tmem_d, al, bh, idesc, (s | j) ? 1u : 0u);
                                    umma_tf32(tmem_d, ah, bl, idesc, 1u);
                                    umma_tf32(tmem_d, ah, bh, idesc, 1u);
                                }
                                ah += (2 * TC_CHUNK_BYTES) >> 4; al += (2 * TC_CHUNK_BYTES) >> 4;
                                bh += b_step; bl += b_step;
                            }
                        }
                        umma_commit(&empty[st]);
                        if (s == nslab - 1) umma_commit(&d_ready);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp_u == PROD / 32 + 1) {
        // ===================== weight producer (bulk-copy TMA) + SA row metadata =====================
        // The weight copies need one elected lane and are never on the critical path, so this warp also
        // resolves the SA gather's dependent loads for the NEXT tile -- neighbour index -> point ->
        // coordinates (and, for layer-0 inputs of <= 8 channels, the features themselves) -- spread over
        // the first three slabs of the current tile, and leaves {point, dx, dy, dz | f0..f3} per row in a
        // two-tile shared-memory ring.  The producers then start a tile with shared-memory reads instead
        // of two L2 round trips (1300 + ~2000 cycles per tile before; sa1 has 1-slab layer 0s).
        uint32_t it = 0, tcnt = 0;
        const int glog2 = a.group > 0 ? __ffs(a.group) - 1 : 0;
        // (the LARGE shape keeps this a run-time flag: its register allocation is better with the branch in place)
        const bool l0_small = SMALL ? RING : (MODE == 0 && a.cin0 <= 8 && a.cfeat <= 4);
        int m_idx[4];
        int m_pnt[4];
        float m_xyz[4][3], m_cen[4][3], m_f[4][4];
        auto meta_phase1 = [&](int64_t tile) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int64_t lg = tile * TC_ROWS + lane + 32 * k;
                m_idx[k] = lg < a.rows ? __ldg(a.idx + lg) : 0;
            }
        };
        auto meta_phase2 = [&](int64_t tile) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int64_t lg = tile * TC_ROWS + lane + 32 * k;
                const bool ok = lg < a.rows;
                const uint32_t cen = (uint32_t)(lg >> glog2);                  // rows / group < 2^31 (host check)
                const int pnt = (int)(cen / (uint32_t)a.s) * a.n + m_idx[k];   // b * n + idx < 2^31 (host check)
                m_pnt[k] = ok ? pnt : -1;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    m_xyz[k][j] = ok ? __ldg(a.xyz + (int64_t)pnt * 3 + j) : 0.f;
                    m_cen[k][j] = ok ? __ldg(a.new_xyz + (int64_t)cen * 3 + j) : 0.f;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) m_f[k][j] = (ok && l0_small && j < a.cfeat) ? __ldg(a.feats + (int64_t)pnt * a.cfeat + j) : 0.f;
            }
        };
        auto meta_phase3 = [&](uint32_t tc) {            // tc = this CTA's running tile count of the tile described
            const uint32_t buf = tc & 1, use = tc >> 1;
            mbar_wait_warp_relaxed(&meta_empty[buf], (use & 1) ^ 1);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float4 *dst = meta_s + ((size_t)buf * TC_ROWS + lane + 32 * k) * 2;
                const float dx = __fsub_rn(m_xyz[k][0], m_cen[k][0]), dy = __fsub_rn(m_xyz[k][1], m_cen[k][1]),
                            dz = __fsub_rn(m_xyz[k][2], m_cen[k][2]);
                if (l0_small) {
                    // the whole layer-0 row, assembled here once: [f0 .. f(cfeat-1), dx, dy, dz, 0 ..] (zeros past the end)
                    float in[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int t = j - a.cfeat;
                        const float v = t < 0 ? (j < 4 ? m_f[k][j < 4 ? j : 0] : 0.f) : (t == 0 ? dx : (t == 1 ? dy : (t == 2 ? dz : 0.f)));
                        in[j] = m_pnt[k] >= 0 ? v : 0.f;
                    }
                    dst[0] = make_float4(in[0], in[1], in[2], in[3]);
                    dst[1] = make_float4(in[4], in[5], in[6], in[7]);
                } else {
                    dst[0] = make_float4(__int_as_float(m_pnt[k]), dx, dy, dz);
                    dst[1] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            __syncwarp();
            if (elect_one()) mbar_arrive(&meta_full[buf]);
            __syncwarp();
        };
        if (MODE == 0 && blockIdx.x < a.ntiles) {
            meta_phase1(blockIdx.x);
            meta_phase2(blockIdx.x);
            meta_phase3(0);
        }
        for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++tcnt) {
            const int64_t tile_next = tile + gridDim.x;
            const bool has_next = MODE == 0 && tile_next < a.ntiles;
            int nissued = 0;
            for (int l = 0; l < a.nlayers; ++l) {
                const int nslab = a.kpad[l] / KC;
                const uint32_t bytes = 2u * NCHUNK * (uint32_t)NPAD(l) * 16u;
                const uint8_t *src = reinterpret_cast<const uint8_t *>(WPK(l));
                for (int s = 0; s < nslab; ++s, ++it) {
                    const uint32_t st = it & (NST - 1), ph = (it >> nst_log2) & 1;
                    mbar_wait_warp_relaxed(&empty[st], ph ^ 1);
                    if (elect_one()) {
                        if (TC_DBG & 2) {
                            mbar_arrive(&full[st]);
                        } else {
                            mbar_arrive_expect_tx(&full[st], bytes);
                            bulk_g2s(w_stage + (size_t)st * a.wstage_bytes, src + (size_t)s * bytes, bytes, &full[st]);
                        }
                    }
                    __syncwarp();
                    if (has_next) {     // one metadata phase after each of the first three copies
                        if (nissued == 0) meta_phase1(tile_next);
                        else if (nissued == 1) meta_phase2(tile_next);
                        else if (nissued == 2) meta_phase3(tcnt + 1);
                    }
                    ++nissued;
                }
            }
            if (has_next) {             // tiles with fewer than three slabs
                if (nissued < 1) meta_phase1(tile_next);
                if (nissued < 2) meta_phase2(tile_next);
                if (nissued < 3) meta_phase3(tcnt + 1);
            }
        }
    } else {
        // ===================== producer / epilogue threads =====================
        const int r = tid & 127;              // tile row == TMEM lane (layers > 0 and the epilogue)
        const int g = tid >> 7;               // producer group == piece of every slab (layers > 0)
        const int cg = 16 * g;                // first channel of this thread's piece within a slab
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        // this thread's 16 bytes of chunk 0 of its piece in stage 0 (hi plane)
        uint8_t *const my_a = a_stage + (size_t)(F16 ? 2 * g : 4 * g) * TC_CHUNK_BYTES + r * 16;
        uint32_t it = 0;                      // global slab counter (the same sequence in every role)
        uint32_t dl = 0;
        float amax = 0.f;                     // largest |operand| this thread converted to fp16
        int aff_cloud = -1;                   // cloud whose scale/shift rows sit in aff_s

        // ---- layer-0 loader geometry.  Layer 0 has no tie to the TMEM lanes, so its rows are dealt out
        // for COALESCED global reads: a warp owns ROWS_W consecutive tile rows; one load instruction
        // covers RPI rows x LPR lanes x 32 bytes (8 fp32 channels = one "unit"), i.e. 128 contiguous
        // bytes per row (LDG.256: 8 cache lines per instruction instead of the 32 a thread-per-row
        // gather touches), and a lane keeps two units per slab.  A unit is exactly one 16-byte fp16
        // operand chunk (two tf32 chunks), and the 8 lanes of a quarter warp hold 8 consecutive rows of
        // the same chunk, so the operand stores stay 16-byte wide and bank-conflict free.
        constexpr int UPR = KC / 8;                     // units per row per slab: 8 | 4 | 4 | 2
        constexpr int LPR = UPR < 4 ? UPR : 4;          // lanes sharing a row in one instruction
        constexpr int RPI = 32 / LPR;                   // rows per instruction
        constexpr int ROWS_W = TC_ROWS / (4 * G);       // rows owned by a warp
        constexpr int RS = ROWS_W / RPI;                // row sets per warp
        constexpr int US = UPR / LPR;                   // unit sets per row
        static_assert(RS * US == 2 && (RS == 1 || US == 1), "two 32-byte units per lane per slab");
        const int lrow = lane % RPI, lcq = lane / RPI;
        auto unit_row = [&](int u) { return warp * ROWS_W + (RS == 2 ? u * RPI : 0) + lrow; };   // tile row of unit u
        auto unit_idx = [&](int u) { return lcq + (US == 2 ? u * LPR : 0); };                     // unit within the slab
        const int glog2 = a.group > 0 ? __ffs(a.group) - 1 : 0;

        struct RowMeta {                      // dense mode: one tile row of the layer-0 loader (SA rows come from the metadata ring)
            bool valid;
            const float *pa, *pb;             // rows of the two input segments
        };
        const bool l0_small = SMALL ? RING : (MODE == 0 && a.cin0 <= 8 && a.cfeat <= 4);   // SA layer 0 entirely inside the metadata ring (compile-time for the SMALL shape: the gather code is not built into the sa1 launches)
        uint32_t tcnt = 0;                    // this CTA's running tile count (metadata ring slot and parity)

        auto warp_wait = [&](uint64_t *bar, uint32_t parity) {   // one lane polls, the warp follows
            if (lane == 0) mbar_wait(bar, parity);
            __syncwarp();
        };
        auto acquire = [&](uint32_t i) {      // wait until the MMAs that last read this slab's stage are done
            warp_wait(&empty[i & (NST - 1)], ((i >> nst_log2) & 1) ^ 1);
        };
        auto release = [&](uint32_t i) {      // hand the filled stage to the MMA thread
            // every writer fences its own generic-proxy stores towards the async proxy; one elected
            // lane per warp then arrives
            if (!(TC_DBG & 8)) fence_proxy_async_smem();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[i & (NST - 1)]);
        };
        // 16 fp32 values -> operand planes.  RELU: the values are pre-activation (layers > 0).
        auto convert = [&](const float (&x)[16], auto relu_tag, TcPiece<F16> &p) {
            constexpr bool RELU = decltype(relu_tag)::value;
            if (F16) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    split2_f16<RELU>(x[2 * j], x[2 * j + 1], p.hi[j], p.lo[j]);
                    amax = RELU ? fmaxf(fmaxf(amax, x[2 * j]), x[2 * j + 1])      // saturation watch, checked once per kernel
                                : fmaxf(fmaxf(amax, fabsf(x[2 * j])), fabsf(x[2 * j + 1]));
                }
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float hi, lo;
                    split_tf32(RELU ? fmaxf(x[j], 0.f) : x[j], hi, lo);
                    p.hi[j] = __float_as_uint(hi); p.lo[j] = __float_as_uint(lo);
                }
            }
        };
        auto store_piece = [&](uint32_t i, const TcPiece<F16> &p) {
            uint8_t *hi = my_a + (size_t)(i & (NST - 1)) * A_STAGE, *lo = hi + A_PLANE;
            constexpr int NQ = F16 ? 2 : 4;   // 16-byte chunks per piece
#pragma unroll
            for (int q = 0; q < NQ; ++q) {
                *reinterpret_cast<uint4 *>(hi + q * TC_CHUNK_BYTES) = make_uint4(p.hi[4 * q], p.hi[4 * q + 1], p.hi[4 * q + 2], p.hi[4 * q + 3]);
                *reinterpret_cast<uint4 *>(lo + q * TC_CHUNK_BYTES) = make_uint4(p.lo[4 * q], p.lo[4 * q + 1], p.lo[4 * q + 2], p.lo[4 * q + 3]);
            }
        };
        // one layer-0 unit (8 signed fp32 values) -> operand words: 4 + 4 packed fp16 pairs, or 8 + 8 tf32
        auto convert_unit = [&](const float (&x)[8], uint32_t (&hi)[F16 ? 4 : 8], uint32_t (&lo)[F16 ? 4 : 8]) {
            if (F16) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    split2_f16<false>(x[2 * j], x[2 * j + 1], hi[j], lo[j]);
                    amax = fmaxf(fmaxf(amax, fabsf(x[2 * j])), fabsf(x[2 * j + 1]));
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float h, l;
                    split_tf32(x[j], h, l);
                    hi[j] = __float_as_uint(h); lo[j] = __float_as_uint(l);
                }
            }
        };
        auto store_unit = [&](uint32_t i, int row, int uq, const uint32_t (&hi)[F16 ? 4 : 8], const uint32_t (&lo)[F16 ? 4 : 8]) {
            uint8_t *ph = a_stage + (size_t)(i & (NST - 1)) * A_STAGE + (size_t)(F16 ? uq : 2 * uq) * TC_CHUNK_BYTES + row * 16;
            uint8_t *pl = ph + A_PLANE;
#pragma unroll
            for (int q = 0; q < (F16 ? 1 : 2); ++q) {
                *reinterpret_cast<uint4 *>(ph + q * TC_CHUNK_BYTES) = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
                *reinterpret_cast<uint4 *>(pl + q * TC_CHUNK_BYTES) = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
            }
        };

        for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
            const int64_t grow = tile * TC_ROWS + r;     // this thread's row in layers > 0 and the epilogue
            const bool valid = grow < a.rows;
            TC_STAMP(0);
            // ---- layer-0 row metadata
            RowMeta meta[RS];
#pragma unroll
            for (int ri = 0; ri < RS; ++ri) {
                const int64_t lg = tile * TC_ROWS + warp * ROWS_W + ri * RPI + lrow;
                meta[ri].valid = MODE == 1 && lg < a.rows;
                meta[ri].pa = meta[ri].pb = nullptr;
                if (meta[ri].valid) {
                    meta[ri].pa = a.segA ? a.segA + lg * a.ldA : nullptr;
                    meta[ri].pb = a.segB ? a.segB + (a.bcast ? lg / a.bcast : lg) * a.ldB : nullptr;
                }
            }
            // SA: this tile's slot of the metadata ring (filled by the TMA warp during the previous tile)
            const float4 *mrow = meta_s + ((size_t)(tcnt & 1) * TC_ROWS + warp * ROWS_W + lrow) * 2;
            if (MODE == 0) warp_wait(&meta_full[tcnt & 1], (tcnt >> 1) & 1);
            if (MODE == 1 && GN && a.in_scale) {
                // a tile lies inside one cloud (rows_per_cloud % 128 == 0, checked on the host): stage the
                // cloud's scale/shift rows (GroupNorm + ReLU of the producer layer) in shared memory
                const int cloud = (int)((tile * TC_ROWS) / a.rows_per_cloud);
                if (cloud != aff_cloud) {
                    asm volatile("bar.sync 1, %0;" ::"n"(PROD) : "memory");   // everyone is done with the previous rows
                    for (int i = tid; i < a.aff_pad; i += PROD) {
                        aff_s[i] = i < a.ca ? __ldg(a.in_scale + (size_t)cloud * a.ca + i) : 0.f;
                        aff_s[a.aff_pad + i] = i < a.ca ? __ldg(a.in_shift + (size_t)cloud * a.ca + i) : 0.f;
                    }
                    asm volatile("bar.sync 1, %0;" ::"n"(PROD) : "memory");
                    aff_cloud = cloud;
                }
            }
            // 8 consecutive layer-0 channels from c0 of one row (row set ri of this lane)
            auto load_unit = [&](int ri, int c0, float (&x)[8]) {
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = 0.f;
                if (c0 >= a.cin0 || (TC_DBG & 65)) return;             // padding channels (probes: no gather)
                const float *pa = nullptr, *pb = nullptr;
                float dx = 0.f, dy = 0.f, dz = 0.f;
                if (MODE == 0) {
                    const float4 m0 = mrow[ri * RPI * 2];
                    if (l0_small) {                                       // the assembled row itself: no global access at all
                        const float4 m1 = mrow[ri * RPI * 2 + 1];
                        x[0] = m0.x; x[1] = m0.y; x[2] = m0.z; x[3] = m0.w;
                        x[4] = m1.x; x[5] = m1.y; x[6] = m1.z; x[7] = m1.w;
                        return;
                    }
                    const int pnt = __float_as_int(m0.x);
                    if (pnt < 0) return;                                  // row past the end
                    dx = m0.y; dy = m0.z; dz = m0.w;
                    pa = a.feats ? a.feats + (int64_t)pnt * a.ldf : nullptr;
                } else {
                    if (!meta[ri].valid) return;                          // row past the end
                    pa = meta[ri].pa; pb = meta[ri].pb;
                }
                const int ca = MODE == 0 ? a.cfeat : a.ca;            // width of the leading segment
                const int na = ca - c0;                               // its channels left from c0
                const float *vp = nullptr;                            // 8 contiguous floats?
                if (na >= 8) vp = pa + c0;
                else if (MODE == 1 && na <= 0 && a.cb + na >= 8) vp = pb - na;
                const uintptr_t al = reinterpret_cast<uintptr_t>(vp);
                if (vp && (al & 31) == 0) {
                    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                                 : "=f"(x[0]), "=f"(x[1]), "=f"(x[2]), "=f"(x[3]), "=f"(x[4]), "=f"(x[5]), "=f"(x[6]), "=f"(x[7])
                                 : "l"(vp));
                } else if (vp && (al & 15) == 0) {
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const float4 t = __ldg(reinterpret_cast<const float4 *>(vp) + q);
                        x[4 * q] = t.x; x[4 * q + 1] = t.y; x[4 * q + 2] = t.z; x[4 * q + 3] = t.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int t = j - na;                         // channel index within the trailing segment
                        if (t < 0) x[j] = __ldg(pa + c0 + j);
                        else if (MODE == 0) x[j] = a.pre_pad ? 0.f : (t == 0 ? dx : (t == 1 ? dy : (t == 2 ? dz : 0.f)));
                        else if (t < a.cb) x[j] = __ldg(pb + t);
                    }
                }
                if (MODE == 0 && a.pre_pad) {
                    // projected layer 0: the gathered row is W_f * features of the point; add the coordinate
                    // part and the bias, apply the ReLU -- this IS the first tensor-core layer's operand
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const float4 wx = *reinterpret_cast<const float4 *>(aff_s + c0 + 4 * q);
                        const float4 wy = *reinterpret_cast<const float4 *>(aff_s + a.pre_pad + c0 + 4 * q);
                        const float4 wz = *reinterpret_cast<const float4 *>(aff_s + 2 * a.pre_pad + c0 + 4 * q);
                        const float4 bb = *reinterpret_cast<const float4 *>(aff_s + 3 * a.pre_pad + c0 + 4 * q);
                        x[4 * q] = fmaxf(fmaf(wz.x, dz, fmaf(wy.x, dy, fmaf(wx.x, dx, x[4 * q] + bb.x))), 0.f);
                        x[4 * q + 1] = fmaxf(fmaf(wz.y, dz, fmaf(wy.y, dy, fmaf(wx.y, dx, x[4 * q + 1] + bb.y))), 0.f);
                        x[4 * q + 2] = fmaxf(fmaf(wz.z, dz, fmaf(wy.z, dy, fmaf(wx.z, dx, x[4 * q + 2] + bb.z))), 0.f);
                        x[4 * q + 3] = fmaxf(fmaf(wz.w, dz, fmaf(wy.w, dy, fmaf(wx.w, dx, x[4 * q + 3] + bb.w))), 0.f);
                    }
                }
            };
            // GroupNorm + ReLU of the producer layer, applied to a loaded unit (dense mode)
            auto affine_unit = [&](const RowMeta &m, int c0, float (&x)[8]) {
                if (MODE != 1 || !GN || !a.in_scale || !m.valid || c0 >= a.cin0) return;
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const float4 sc = *reinterpret_cast<const float4 *>(aff_s + c0 + 4 * q);
                    const float4 sh = *reinterpret_cast<const float4 *>(aff_s + a.aff_pad + c0 + 4 * q);
                    x[4 * q] = fmaxf(fmaf(x[4 * q], sc.x, sh.x), 0.f);
                    x[4 * q + 1] = fmaxf(fmaf(x[4 * q + 1], sc.y, sh.y), 0.f);
                    x[4 * q + 2] = fmaxf(fmaf(x[4 * q + 2], sc.z, sh.z), 0.f);
                    x[4 * q + 3] = fmaxf(fmaf(x[4 * q + 3], sc.w, sh.w), 0.f);
                }
            };
            TC_STAMP(1);
            // ---------------- layer 0 ----------------
            if constexpr (RING) {
                // <= 8 input channels: one slab, the row comes assembled from the metadata ring (straight-line code:
                // no prefetch buffers, no loops -- these launches are instruction-issue bound)
                const int kstore = (a.kreal[0] + KMMA - 1) / KMMA * KMMA;
                // no acquire: the previous tile's last epilogue waited for d_ready, i.e. for EVERY MMA issued so far, so
                // all stages are free at a tile boundary (the same holds for the first NST slabs of any layer, below)
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int c0 = 8 * unit_idx(u);
                    if (c0 < kstore && !(TC_DBG & 1)) {
                        float x[8];
                        load_unit(RS == 2 ? u : 0, c0, x);
                        uint32_t hi[F16 ? 4 : 8], lo[F16 ? 4 : 8];
                        convert_unit(x, hi, lo);
                        store_unit(it, unit_row(u), unit_idx(u), hi, lo);
                    }
                }
                release(it);
                ++it;
            } else {
                // coalesced gather, one or two slabs ahead
                const int nslab = a.kpad[0] / KC;
                const int kstore = (a.kreal[0] + KMMA - 1) / KMMA * KMMA;   // channels the MMA k-steps read
                // prefetch distance: two slabs where the register budget allows (3xTF32 LARGE: 168), else one (96)
#ifdef CAPTRA_TC_PF2      // A/B build knob: CAPTRA_EXTRA_NVCC_FLAGS=-DCAPTRA_TC_PF2 CAPTRA_LIB_OUT=... python -m captra_b200.build
                constexpr int PF = SMALL ? 1 : 2;
#else
                constexpr int PF = (SMALL || F16) ? 1 : 2;
#endif
                float b0[2][8], b1[PF == 2 ? 2 : 1][8];
                auto load_slab = [&](int s, float (&buf)[2][8]) {
#pragma unroll
                    for (int u = 0; u < 2; ++u) load_unit(RS == 2 ? u : 0, s * KC + 8 * unit_idx(u), buf[u]);
                };
                auto step = [&](int s, float (&buf)[2][8]) {   // buf holds slab s and is refilled with slab s + PF
                    if (s >= (int)NST) acquire(it);            // the first NST slabs of a layer find their stages free (see above)
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int c0 = s * KC + 8 * unit_idx(u);
                        if (c0 < kstore && !(TC_DBG & 1)) {        // else: pure padding, the MMA thread skips these k-steps
                            uint32_t hi[F16 ? 4 : 8], lo[F16 ? 4 : 8];
                            affine_unit(meta[RS == 2 ? u : 0], c0, buf[u]);
                            convert_unit(buf[u], hi, lo);
                            store_unit(it, unit_row(u), unit_idx(u), hi, lo);
                        }
                    }
                    release(it);
                    // refill only now: the proxy fence in release() waits for the thread's outstanding
                    // loads, which would put the L2 latency of the prefetch on every slab's critical path
                    if (s + PF < nslab) load_slab(s + PF, buf);
                    ++it;
                };
                load_slab(0, b0);
                if constexpr (PF == 2) {
                    if (nslab > 1) load_slab(1, b1);
                    for (int s = 0; s < nslab; s += 2) {
                        step(s, b0);
                        if (s + 1 < nslab) step(s + 1, b1);
                    }
                } else {
                    for (int s = 0; s < nslab; ++s) step(s, b0);
                }
            }
            if (MODE == 0) {              // layer 0 has read this tile's metadata: the slot may be refilled
                __syncwarp();
                if (lane == 0) mbar_arrive(&meta_empty[tcnt & 1]);
            }
            ++tcnt;
            TC_STAMP(2);
            // ---------------- layers 1..L-1: previous accumulator -> next operand ----------------
            int bias_off = 0;
            for (int l = 1; l < a.nlayers; ++l) {
                const int nslab = a.kpad[l] / KC, kreal = a.kreal[l];     // kreal == NPAD(l-1)
                const uint32_t tsrc = tmem_base + (uint32_t)(((l - 1) & 1) * a.region_cols) + lane_addr + (uint32_t)cg;
                const float *bz = bias_s + bias_off + cg;
                const float inv = bias_s[bias_off + NPAD(l - 1)];         // 1 / weight scale of layer l-1
                warp_wait(&d_ready, dl & 1);
                ++dl;
                tcgen05_fence_after();
                TC_STAMP(10 + l);
                uint32_t v[16];
                if (cg < kreal && !(TC_DBG & 129)) tmem_ld_32x16(tsrc, v);
                for (int s = 0; s < nslab; ++s, ++it) {
                    const int c0 = s * KC + cg;
                    const bool active = c0 < kreal && !(TC_DBG & 1);
                    TcPiece<F16> p;
                    if (active) {
                        tmem_ld_wait();
                        float x[16];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 b4 = *reinterpret_cast<const float4 *>(bz + s * KC + 4 * q);
                            x[4 * q + 0] = fmaf(__uint_as_float(v[4 * q + 0]), inv, b4.x);
                            x[4 * q + 1] = fmaf(__uint_as_float(v[4 * q + 1]), inv, b4.y);
                            x[4 * q + 2] = fmaf(__uint_as_float(v[4 * q + 2]), inv, b4.z);
                            x[4 * q + 3] = fmaf(__uint_as_float(v[4 * q + 3]), inv, b4.w);
                        }
                        // overlap the next TMEM read with the conversion (the empty asm pins the fmas
                        // before it, so the compiler does not copy v to keep the old values alive)
#pragma unroll
                        for (int j = 0; j < 16; ++j) asm volatile("" : "+f"(x[j]));
                        if (s + 1 < nslab && c0 + KC < kreal && !(TC_DBG & 128)) tmem_ld_32x16(tsrc + (uint32_t)((s + 1) * KC), v);
                        convert(x, std::true_type{}, p);
                    }
                    if (s >= (int)NST) acquire(it);            // d_ready of the previous layer already covers the first NST slabs
                    if (active) store_piece(it, p);
                    release(it);
                }
                bias_off += NPAD(l - 1) + TC_BIAS_PAD;
                TC_STAMP(20 + l);
            }
            // ---------------- last epilogue ----------------
            {
                const int l = a.nlayers - 1;
                const int npad = NPAD(l);
                const uint32_t tbase = tmem_base + (uint32_t)((l & 1) * a.region_cols) + lane_addr;
                warp_wait(&d_ready, dl & 1);
                ++dl;
                tcgen05_fence_after();
                TC_STAMP(30);
                const float inv = bias_s[bias_off + npad];   // 1 / weight scale of the last layer
                const float *bl = bias_s + bias_off;
                if (GRP) {    // compile-time: each launch carries only the epilogue it runs
                    // Transposed accumulator: TMEM lane = output channel (block cb, lane r), column = tile
                    // row (point).  Group g reduces the columns [g*SEG, (g+1)*SEG) of every channel block,
                    // 32 columns (one smallest centroid group) at a time: a per-thread max of the raw
                    // accumulators -- inv > 0, so bias and ReLU are applied to the maxima only.
                    constexpr int SEG = TC_ROWS / G;              // 32 | 64 | 128 columns per producer group
                    const int nblk = (npad + 127) / 128;
                    // this tile's half of the buffer: the next tile writes the other half, so no barrier is needed after
                    // the read-out below (a half is rewritten two tiles later, after everybody passed the next tile's barrier)
                    const bool red2 = SMALL ? a.red2 != 0 : true;     // (LARGE always has room: compile-time, no extra live flag at 96 registers)
                    float *redt = red + (red2 ? (tcnt & 1) * 1024 : 0);
                    for (int cb = 0; cb < nblk; ++cb) {
                        const uint32_t tsrc = tbase + (uint32_t)(cb * 128 + g * SEG);
                        uint32_t v0[16], v1[16];
                        tmem_ld_32x16(tsrc, v0);
#pragma unroll
                        for (int q = 0; q < SEG / 32; ++q) {
                            tmem_ld_32x16(tsrc + (uint32_t)(32 * q + 16), v1);
                            tmem_ld_wait();                       // waits for both; v0 is consumed first
                            float m = __uint_as_float(v0[0]);
#pragma unroll
                            for (int j = 1; j < 16; ++j) m = fmaxf(m, __uint_as_float(v0[j]));
#pragma unroll
                            for (int j = 0; j < 16; ++j) m = fmaxf(m, __uint_as_float(v1[j]));
                            asm volatile("" : "+f"(m));
                            if (q + 1 < SEG / 32) tmem_ld_32x16(tsrc + (uint32_t)(32 * (q + 1)), v0);
                            redt[(g * (SEG / 32) + q) * 256 + cb * 128 + r] = m;
                        }
                    }
                    // combine the 32-column maxima that belong to one centroid and write them out
                    const int wpg = a.group / 32;                 // 32-column segments per centroid: 1, 2 or 4
                    tcgen05_fence_before();
                    asm volatile("bar.sync 1, %0;" ::"n"(PROD) : "memory");
                    const int ngroups = 4 / wpg;
                    for (int o = tid; o < ngroups * npad; o += PROD) {
                        const int gi = o / npad, c = o - gi * npad;
                        const int64_t cen = (tile * TC_ROWS) / a.group + gi;
                        if (c >= cout_last || cen * a.group >= a.rows) continue;
                        float m = redt[(gi * wpg) * 256 + c];
                        for (int w = 1; w < wpg; ++w) m = fmaxf(m, redt[(gi * wpg + w) * 256 + c]);
                        a.out[cen * a.ldo + col_off + c] = fmaxf(fmaf(m, inv, bl[c]), 0.f);
                    }
                    if (!red2) asm volatile("bar.sync 1, %0;" ::"n"(PROD) : "memory");   // single buffer: reused by the next tile
                } else {
                    // row-major accumulator: the groups alternate 16-column chunks of the thread's row.  A thread
                    // owns a ROW of the accumulator, so storing it directly makes every store instruction touch 32
                    // rows (32 cache lines, half a sector each).  The warp's 32 x 16 block is instead transposed
                    // through shared memory so that a store instruction writes 8 rows x 64 contiguous bytes.
                    // Staging lives in the A-operand stages, which are idle here (every MMA of the tile has
                    // completed): each warp uses only the bytes IT writes in layer 0 (rows warp*ROWS_W.. of every
                    // chunk), so a warp that runs ahead into the next tile's layer 0 cannot clobber it, and nobody
                    // reaches layer 1 of the next tile before all warps have left this epilogue.
                    constexpr int SEG = ROWS_W * 16;                  // contiguous bytes this warp owns per operand chunk
                    static_assert(2048 / SEG <= NCHUNK * 2 * (int)NST, "staging does not fit the warp's share of the A stages");
                    auto stg = [&](int o) -> uint8_t * {               // byte o of the warp's 2 KB staging tile
                        const int seg = o / SEG;
                        return a_stage + (size_t)(seg % NCHUNK) * TC_CHUNK_BYTES + (size_t)(seg / NCHUNK) * A_PLANE + warp * SEG + (o % SEG);
                    };
                    const int64_t row0 = tile * TC_ROWS + (warp & 3) * 32;        // first row of this warp's TMEM lane quarter
                    const bool vec_ok = ((reinterpret_cast<uintptr_t>(a.out + col_off) & 15) == 0) && ((a.ldo & 3) == 0);
                    const uint32_t tsrc = tbase;
                    uint32_t v[16];
                    int c0 = cg;
                    if (c0 < npad) tmem_ld_32x16(tsrc + (uint32_t)c0, v);
                    for (; c0 < npad; c0 += KC) {
                        tmem_ld_wait();
                        float x[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) x[j] = fmaf(__uint_as_float(v[j]), inv, bl[c0 + j]);
#pragma unroll
                        for (int j = 0; j < 16; ++j) asm volatile("" : "+f"(x[j]));
                        if (c0 + KC < npad) tmem_ld_32x16(tsrc + (uint32_t)(c0 + KC), v);
                        if (a.relu_last) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) x[j] = fmaxf(x[j], 0.f);
                        }
                        // row `lane`, 16-byte quad q -> slot (q ^ (lane >> 1)) & 3 of the row's 64 bytes: conflict-free both ways
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            *reinterpret_cast<float4 *>(stg(lane * 64 + ((q ^ (lane >> 1)) & 3) * 16)) = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
                        __syncwarp();
                        float4 ssum = make_float4(0.f, 0.f, 0.f, 0.f), ssq = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int R = (lane >> 2) + 8 * i, q = lane & 3;
                            const float4 t = *reinterpret_cast<const float4 *>(stg(R * 64 + ((q ^ (R >> 1)) & 3) * 16));
                            const int64_t orow = row0 + R;
                            const int c = c0 + 4 * q;
                            if (GN && a.stats && orow < a.rows) {
                                ssum.x += t.x; ssum.y += t.y; ssum.z += t.z; ssum.w += t.w;
                                ssq.x = fmaf(t.x, t.x, ssq.x); ssq.y = fmaf(t.y, t.y, ssq.y);
                                ssq.z = fmaf(t.z, t.z, ssq.z); ssq.w = fmaf(t.w, t.w, ssq.w);
                            }
                            if (orow < a.rows && c < cout_last) {
                                float *dst = a.out + orow * a.ldo + col_off + c;
                                if (vec_ok && c + 4 <= cout_last) {
                                    *reinterpret_cast<float4 *>(dst) = t;
                                } else {
                                    dst[0] = t.x;
                                    if (c + 1 < cout_last) dst[1] = t.y;
                                    if (c + 2 < cout_last) dst[2] = t.z;
                                    if (c + 3 < cout_last) dst[3] = t.w;
                                }
                            }
                        }
                        if (GN && a.stats) {
                            // GroupNorm statistics of this layer's output, fused into its epilogue: column sums over
                            // the warp's 32 rows (lanes that share q hold the same 4 columns), one plain store per
                            // (32-row block, column) -- deterministic, no atomics; captra_group_norm_finalize
                            // combines the blocks of a cloud in fp64
#pragma unroll
                            for (int o = 4; o < 32; o <<= 1) {
                                ssum.x += __shfl_xor_sync(kFull, ssum.x, o); ssum.y += __shfl_xor_sync(kFull, ssum.y, o);
                                ssum.z += __shfl_xor_sync(kFull, ssum.z, o); ssum.w += __shfl_xor_sync(kFull, ssum.w, o);
                                ssq.x += __shfl_xor_sync(kFull, ssq.x, o); ssq.y += __shfl_xor_sync(kFull, ssq.y, o);
                                ssq.z += __shfl_xor_sync(kFull, ssq.z, o); ssq.w += __shfl_xor_sync(kFull, ssq.w, o);
                            }
                            const int c = c0 + 4 * lane;             // lanes 0..3 <-> q
                            if (lane < 4 && c < cout_last) {
                                const int ccol = col_off - a.col_off + c;                      // column within the layer's output
                                float *sp = a.stats + (size_t)(row0 >> 5) * 2 * a.cout_total + ccol;
                                const float vs[4] = {ssum.x, ssum.y, ssum.z, ssum.w}, vq[4] = {ssq.x, ssq.y, ssq.z, ssq.w};
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    if (c + j < cout_last) { sp[j] = vs[j]; sp[a.cout_total + j] = vq[j]; }
                            }
                        }
                        __syncwarp();                                 // the next chunk reuses the staging tile
                    }
                }
                tcgen05_fence_before();
                TC_STAMP(31);
            }
        }
        if (F16 && !(amax < 65000.f)) g_f16_overflow = 1;   // also catches NaN
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == PROD / 32) tmem_dealloc_dyn(tmem_base, (uint32_t)a.tmem_cols);
}

// bias -> [npad + 16]: bias, then 1/scale at [npad] and scale at [npad+1].  The weight scale is a power of
// two that brings max|W| of the layer into [2^12, 2^13) for fp16 operands (1 for TF32).
__global__ void tc_scale_kernel(int cin, int cout, int npad, int f16, const float *__restrict__ w,
                                const float *__restrict__ bias, float *__restrict__ bp) {
    __shared__ float red[256];
    float m = 0.f;
    for (int i = threadIdx.x; i < cin * cout; i += blockDim.x) m = fmaxf(m, fabsf(w[i]));
    red[threadIdx.x] = m;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + o]);
        __syncthreads();
    }
    for (int c = threadIdx.x; c < npad + TC_BIAS_PAD; c += blockDim.x) bp[c] = (c < cout && bias) ? bias[c] : 0.f;
    __syncthreads();
    if (threadIdx.x == 0) {
        float scale = 1.f;
        if (f16 && red[0] > 0.f && isfinite(red[0])) {
            int e;
            frexpf(red[0], &e);                 // red[0] = f * 2^e, f in [0.5, 1)
            scale = ldexpf(1.f, 13 - e);        // max|W| * scale in [2^12, 2^13)
        }
        bp[npad] = 1.f / scale;
        bp[npad + 1] = scale;
    }
}

// weights -> [slab][plane hi|lo][8 chunks][npad][16 bytes], pre-split, zero padded
template <bool F16>
__global__ void pack_tc_kernel(int cin, int cout, int kpad, int npad, int nchunk, const float *__restrict__ w,
                               const float *__restrict__ bp, float *__restrict__ wpk) {
    constexpr int EPC = F16 ? 8 : 4;                           // elements per 16-byte chunk
    const int TC_NCHUNK = nchunk, KC = nchunk * EPC;           // chunks / channels per slab
    const int nslab = kpad / KC;
    const int plane_bytes = TC_NCHUNK * npad * 16;
    const float scale = bp[npad + 1];
    const int total = nslab * TC_NCHUNK * npad * EPC;          // one thread per element -> writes hi and lo
    uint8_t *out = reinterpret_cast<uint8_t *>(wpk);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int e = i % EPC;
        const int n = (i / EPC) % npad;
        const int q = (i / EPC / npad) % TC_NCHUNK;
        const int s = i / EPC / npad / TC_NCHUNK;
        const int k = s * KC + q * EPC + e;
        const float v = (k < cin && n < cout) ? w[(size_t)n * cin + k] * scale : 0.f;
        const size_t o = (size_t)s * 2 * plane_bytes + ((size_t)q * npad + n) * 16;
        if (F16) {
            unsigned short hi, lo;
            split_f16(v, hi, lo);
            reinterpret_cast<unsigned short *>(out + o)[e] = hi;
            reinterpret_cast<unsigned short *>(out + o + plane_bytes)[e] = lo;
        } else {
            float hi, lo;
            split_tf32(v, hi, lo);
            reinterpret_cast<float *>(out + o)[e] = hi;
            reinterpret_cast<float *>(out + o + plane_bytes)[e] = lo;
        }
    }
}

struct TcLayout {
    int nlayers, kpad[CAPTRA_MAX_MLP_LAYERS], npad[CAPTRA_MAX_MLP_LAYERS], kreal[CAPTRA_MAX_MLP_LAYERS];
    size_t off_w[CAPTRA_MAX_MLP_LAYERS], off_b[CAPTRA_MAX_MLP_LAYERS], total_floats;
    int wstage_bytes, bias_floats, nsplit, last_npad;
    int nstages, region_cols, tmem_cols, target_occ, red2;
    size_t smem_bytes;
    bool supported, small;
    int nchunk;
};

static int pow2_at_least(int v, int lo) {
    int p = lo;
    while (p < v) p <<= 1;
    return p;
}

static TcLayout tc_layout_v(const captra_mlp_desc &d, bool f16, bool small) {
    const int TC_KC = tc_kc(f16, small);
    const int TC_NCHUNK = small ? 4 : 8;
    const int TC_A_STAGE = 2 * TC_NCHUNK * TC_CHUNK_BYTES;
    TcLayout L{};
    L.small = small; L.nchunk = TC_NCHUNK;
    L.nlayers = d.nlayers;
    L.supported = true;
    L.nsplit = 1;
    size_t off = 0;
    int cin = d.cin, npmax = 32;
    for (int l = 0; l < d.nlayers; ++l) {
        // input channels that carry data: the true width of layer 0, then the previous layer's
        // accumulator columns (its outputs rounded up to the MMA N granularity; the columns past cout
        // are exact zeros).  The slab padding beyond kreal is never produced nor multiplied.
        L.kreal[l] = cin;
        L.kpad[l] = round_up(cin, TC_KC);
        L.npad[l] = round_up(d.cout[l], 16);
        if (L.npad[l] > 256) {
            if (d.nlayers == 1) {   // one wide layer: 256-column chunks over grid.y
                L.nsplit = ceil_div(L.npad[l], 256);
                L.last_npad = L.npad[l] - 256 * (L.nsplit - 1);
            } else {
                L.supported = false;
            }
        }
        npmax = max(npmax, min(L.npad[l], 256));
        // packed weights of a layer: nslab * 2 planes * NCHUNK chunks * npad * 16 bytes (in floats: /4)
        L.off_w[l] = off; off += (size_t)(L.kpad[l] / TC_KC) * 2 * TC_NCHUNK * L.npad[l] * 4;
        const int nb = L.nsplit > 1 ? L.nsplit * (256 + TC_BIAS_PAD) : L.npad[l] + TC_BIAS_PAD;
        L.off_b[l] = off; off += nb;
        L.bias_floats += nb;
        cin = L.npad[l];           // the next layer sees the padded width (zero weights on the pad)
    }
    L.total_floats = off;
    L.wstage_bytes = 2 * TC_NCHUNK * npmax * 16;
    // TMEM: accumulators of consecutive layers ping-pong between two regions
    L.region_cols = pow2_at_least(npmax, 32);
    L.tmem_cols = d.nlayers > 1 ? 2 * L.region_cols : L.region_cols;
    L.target_occ = 1;
    // LARGE: 4 pipeline stages if they fit in 227 KB, else 2.  SMALL: 2 stages, two CTAs per SM.
    const size_t fixed = (size_t)L.bias_floats * 4 + 256 + 4096 + 8192;   // biases, slack, the grouped epilogue's [4][256] buffer, the SA metadata ring
    const size_t per_stage = (size_t)TC_A_STAGE + L.wstage_bytes;
    const size_t budget = small ? (size_t)(227 * 1024) / 2 - 2048 : (size_t)225 * 1024;
    L.nstages = (!small && fixed + 4 * per_stage <= budget) ? 4 : 2;
    L.smem_bytes = fixed + (size_t)L.nstages * per_stage;
    if (L.smem_bytes > budget) L.supported = false;
    // a second [4][256] grouped-max buffer (double-buffered by tile parity: one barrier less per tile) where it fits
    L.red2 = (!small || L.smem_bytes + 4096 + 4096 <= budget) ? 1 : 0;   // SMALL: only with room to spare (+ the projected-layer-0 table some launches add); LARGE: always
    if (L.red2) L.smem_bytes += 4096;
    if (L.smem_bytes > budget) L.supported = false;
    // SMALL keeps two CTAs on an SM, so a CTA owns 256 TMEM columns: fused chains ping-pong between two regions
    // (layers <= 128 columns); a SINGLE layer needs one region and may be up to 256 columns wide (its 512-wide
    // variants split over grid.y as in the LARGE shape).  Single-layer launches -- the RotationRegressor heads, sa3,
    // fp3 -- are latency-bound (ncu: tensor pipe 30 %, issue slots 39 %, 28 % of the warps active): a tile's
    // load -> convert -> MMA -> epilogue chain is serial inside a CTA, so two half-size CTAs per SM overlap one
    // tile's epilogue and global loads with the other's MMAs.  CAPTRA_TC_SMALL_WIDE=0 restores one LARGE CTA per SM.
    static const bool small_wide = [] { const char *e = getenv("CAPTRA_TC_SMALL_WIDE"); return !e || atoi(e) != 0; }();
    if (small && d.nlayers >= 2 && npmax > 128) L.supported = false;
    if (small && d.nlayers < 2 && !small_wide) L.supported = false;
    return L;
}

// fused chains whose layers are all <= 128 columns run the SMALL kernel shape
static TcLayout tc_layout(const captra_mlp_desc &d, bool f16) {
    const TcLayout S = tc_layout_v(d, f16, true);
    return S.supported ? S : tc_layout_v(d, f16, false);
}

int64_t tc_pack_bytes(const captra_mlp_desc *d, bool f16) {
    const TcLayout L = tc_layout(*d, f16);
    return L.supported ? (int64_t)(L.total_floats * sizeof(float)) : -1;
}

int tc_pack(const captra_mlp_desc *d, void *packed, bool f16, cudaStream_t stream) {
    const TcLayout L = tc_layout(*d, f16);
    const int KC = tc_kc(f16, L.small), TC_NCHUNK = L.nchunk;
    CAPTRA_REQUIRE(L.supported, "mlp_pack(tc): layer widths not supported by the tcgen05 path");
    int cin = d->cin;      // true input width of layer l (the packed K is zero padded to L.kpad[l])
    for (int l = 0; l < d->nlayers; ++l) {
        float *base = reinterpret_cast<float *>(packed);
        for (int y = 0; y < L.nsplit; ++y) {   // nsplit > 1 only for a single wide layer
            const int npad_y = L.nsplit == 1 ? L.npad[l] : (y < L.nsplit - 1 ? 256 : L.last_npad);
            const int cout_y = L.nsplit == 1 ? d->cout[l] : min(256, d->cout[l] - 256 * y);
            const float *w_y = d->w[l] + (size_t)y * 256 * cin;
            float *bp = base + L.off_b[l] + (size_t)y * (256 + TC_BIAS_PAD);
            float *wp = base + L.off_w[l] + (size_t)y * (L.kpad[l] / KC) * 2 * TC_NCHUNK * 256 * 4;
            tc_scale_kernel<<<1, 256, 0, stream>>>(cin, cout_y, npad_y, f16 ? 1 : 0, w_y, d->bias[l] ? d->bias[l] + y * 256 : nullptr, bp);
            CAPTRA_CHECK_LAUNCH("mlp_pack(tc scale)");
            if (f16) pack_tc_kernel<true><<<64, 256, 0, stream>>>(cin, cout_y, L.kpad[l], npad_y, TC_NCHUNK, w_y, bp, wp);
            else pack_tc_kernel<false><<<64, 256, 0, stream>>>(cin, cout_y, L.kpad[l], npad_y, TC_NCHUNK, w_y, bp, wp);
            CAPTRA_CHECK_LAUNCH("mlp_pack(tc)");
        }
        cin = d->cout[l];
    }
    return CAPTRA_OK;
}

static int tc_fill(TcArgs &a, const captra_mlp_desc *d, const void *packed, bool f16, size_t *smem) {
    const TcLayout L = tc_layout(*d, f16);
    a.f16 = f16 ? 1 : 0; a.small = L.small ? 1 : 0;
    CAPTRA_REQUIRE(L.supported, "mlp(tc): layer widths not supported by the tcgen05 path");
    a.nlayers = d->nlayers; a.relu_last = d->relu_last; a.cout_last = d->cout[d->nlayers - 1];
    for (int l = 0; l < d->nlayers; ++l) {
        a.kpad[l] = L.kpad[l]; a.npad[l] = L.npad[l];
        a.wpk[l] = reinterpret_cast<const float *>(packed) + L.off_w[l];
        a.bias[l] = reinterpret_cast<const float *>(packed) + L.off_b[l];
    }
    a.wstage_bytes = L.wstage_bytes; a.bias_floats = L.bias_floats;
    a.region_cols = L.region_cols; a.tmem_cols = L.tmem_cols; a.nst_log2 = L.nstages == 4 ? 2 : 1;
    a.red2 = L.red2;
    a.nsplit = L.nsplit; a.last_npad = L.last_npad; a.cout_total = d->cout[d->nlayers - 1];
    for (int l = 0; l < d->nlayers; ++l) a.kreal[l] = L.kreal[l];
    a.cin0 = d->cin;
    { const char *e = getenv("CAPTRA_TC_DBG"); a.dbg = e ? atoi(e) : 0; }
    *smem = L.smem_bytes;
    return CAPTRA_OK;
}

template <int MODE, bool F16, bool SMALL, int NSTL2, bool RING = false, bool GRP = (MODE == 0), bool GN = false>
static int tc_launch_t(TcArgs &a, size_t smem, cudaStream_t stream) {
    if (MODE == 0 && SMALL && !RING && a.cin0 <= 8 && a.cfeat <= 4 && !a.pre_pad) return tc_launch_t<MODE, F16, SMALL, NSTL2, MODE == 0 && SMALL, true>(a, smem, stream);
    if (MODE == 1 && !GRP && a.group > 0) return tc_launch_t<MODE, F16, SMALL, NSTL2, false, true>(a, smem, stream);
    if (MODE == 1 && !GRP && !GN && (a.in_scale || a.stats)) return tc_launch_t<MODE, F16, SMALL, NSTL2, false, false, MODE == 1>(a, smem, stream);
    auto kern = mlp_tc_kernel<MODE, F16, SMALL, NSTL2, RING, GRP, GN>;
    CAPTRA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CAPTRA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    a.ntiles = ceil_div<int64_t>(a.rows, TC_ROWS);
    a.tpose = a.group > 0 ? 1 : 0;
    if (a.tpose) {
        // the transposed last layer accumulates 128-channel blocks of 128 columns (points) each
        // (a wide single layer split over grid.y: every CTA owns at most 256 channels = two blocks)
        const int need = a.nsplit > 1 ? 256 : ceil_div(a.npad[a.nlayers - 1], 128) * 128;
        if (a.region_cols < need) {
            a.region_cols = need;     // 128 or 256: already a power of two
            a.tmem_cols = a.nlayers > 1 ? 2 * need : need;
        }
        CAPTRA_REQUIRE(a.tmem_cols <= (SMALL ? 256 : 512), "mlp(tc): grouped max does not fit the tensor memory");
    }
    const int64_t slots = (int64_t)max(1, sm_count() / a.nsplit) * (SMALL ? 2 : 1);
    const int gx = (int)(a.ntiles < slots ? a.ntiles : slots);
    kern<<<dim3(gx, a.nsplit), tc_threads(F16, SMALL), smem, stream>>>(a);
    CAPTRA_CHECK_LAUNCH("mlp_tc");
    return CAPTRA_OK;
}
template <int MODE>
static int tc_launch(TcArgs &a, size_t smem, cudaStream_t stream) {
    if (a.small) return a.f16 ? tc_launch_t<MODE, true, true, 1>(a, smem, stream) : tc_launch_t<MODE, false, true, 1>(a, smem, stream);
    if (a.nst_log2 == 2) return a.f16 ? tc_launch_t<MODE, true, false, 2>(a, smem, stream) : tc_launch_t<MODE, false, false, 2>(a, smem, stream);
    return a.f16 ? tc_launch_t<MODE, true, false, 1>(a, smem, stream) : tc_launch_t<MODE, false, false, 1>(a, smem, stream);
}

int tc_sa_mlp_max(int b, int n, int s, int k, int cfeat, const float *xyz, const float *new_xyz,
                  const float *feats, const int *idx, const captra_mlp_desc *d, const void *packed,
                  float *out, int64_t ldo, int col_off, bool f16, cudaStream_t stream) {
    CAPTRA_REQUIRE(k == 32 || k == 64 || k == 128, "sa_mlp_max(tc): nsample must be 32, 64 or 128 (got %d)", k);
    CAPTRA_REQUIRE(d->relu_last, "sa_mlp_max(tc): the max epilogue needs a ReLU after the last layer");
    CAPTRA_REQUIRE((int64_t)b * s < 2147483647LL && (int64_t)b * n < 2147483647LL,
                   "sa_mlp_max(tc): too many centroids or points (b * s = %lld, b * n = %lld)", (long long)b * s, (long long)b * n);
    TcArgs a{};
    size_t smem;
    int rc = tc_fill(a, d, packed, f16, &smem);
    if (rc) return rc;
    a.rows = (int64_t)b * s * k; a.group = k;
    a.out = out; a.ldo = ldo; a.col_off = col_off;
    a.n = n; a.s = s; a.cfeat = cfeat; a.xyz = xyz; a.new_xyz = new_xyz; a.feats = feats; a.idx = idx;
    a.ldf = cfeat;
    return tc_launch<0>(a, smem, stream);
}

// SA scale with a PROJECTED layer 0.  Layer 0 of an SA scale is linear in [features | x_j - c]; its feature part
// W_f f_j depends on the point only, while a point sits in nsample * S / N balls (32 for sa2 K=128).  The caller
// therefore computes P = F W_f^T once per cloud (one dense launch for all scales) and this entry gathers rows of
// P instead of rows of F: h0 = relu(P[j] + W_x (x_j - c) + b0) is formed by the loader and is the first
// tensor-core layer's operand.  For sa2 that removes 35 % of the scale's MACs, 60 % of its gather bytes and the
// slowest stage of the tile pipeline.  `d` / `packed` describe layers 1.. of the scale (cin = cpre).
int tc_sa_mlp_max_pre(int b, int n, int s, int k, int cpre, const float *xyz, const float *new_xyz, const float *pre,
                      int64_t ldpre, const float *tab, const int *idx, const captra_mlp_desc *d, const void *packed,
                      float *out, int64_t ldo, int col_off, bool f16, cudaStream_t stream) {
    CAPTRA_REQUIRE(k == 32 || k == 64 || k == 128, "sa_mlp_max_pre(tc): nsample must be 32, 64 or 128 (got %d)", k);
    CAPTRA_REQUIRE(d->relu_last, "sa_mlp_max_pre(tc): the max epilogue needs a ReLU after the last layer");
    CAPTRA_REQUIRE((int64_t)b * s < 2147483647LL && (int64_t)b * n < 2147483647LL, "sa_mlp_max_pre(tc): too many centroids or points");
    CAPTRA_REQUIRE(cpre > 8, "sa_mlp_max_pre(tc): projected width must exceed 8 channels (got %d)", cpre);
    TcArgs a{};
    size_t smem;
    int rc = tc_fill(a, d, packed, f16, &smem);
    if (rc) return rc;
    a.rows = (int64_t)b * s * k; a.group = k;
    a.out = out; a.ldo = ldo; a.col_off = col_off;
    a.n = n; a.s = s; a.cfeat = cpre; a.xyz = xyz; a.new_xyz = new_xyz; a.feats = pre; a.idx = idx;
    a.ldf = ldpre; a.pre_tab = tab; a.pre_pad = round_up(cpre, 16);
    const size_t extra = (size_t)4 * a.pre_pad * sizeof(float);
    const size_t budget = a.small ? (size_t)(227 * 1024) / 2 - 2048 : (size_t)225 * 1024;
    CAPTRA_REQUIRE(smem + extra <= budget, "sa_mlp_max_pre(tc): %d projected channels do not fit the shared-memory budget", cpre);
    smem += extra;
    return tc_launch<0>(a, smem, stream);
}

static int tc_point_mlp_ex(int64_t rows, const float *segA, int64_t ldA, int ca, const float *segB, int64_t ldB, int cb,
                           int bcast, const captra_mlp_desc *d, const void *packed, float *y, int64_t ldy, int col_off,
                           int group, bool f16, cudaStream_t stream, const float *in_scale, const float *in_shift,
                           int rows_per_cloud, float *stats = nullptr) {
    CAPTRA_REQUIRE(group == 0 || ((group == 32 || group == 64 || group == 128) && d->relu_last),
                   "point_mlp(tc): grouped max needs group in {32,64,128} and a final ReLU");
    TcArgs a{};
    size_t smem;
    int rc = tc_fill(a, d, packed, f16, &smem);
    if (rc) return rc;
    a.rows = rows; a.group = group;
    a.out = y; a.ldo = ldy; a.col_off = col_off;
    a.segA = segA; a.ldA = ldA; a.ca = ca; a.segB = segB; a.ldB = ldB; a.cb = cb; a.bcast = bcast;
    a.in_scale = in_scale; a.in_shift = in_shift; a.rows_per_cloud = rows_per_cloud;
    CAPTRA_REQUIRE(stats == nullptr || group == 0, "point_mlp: output statistics need row output (group 0)");
    a.stats = stats;
    if (a.small && group == 0) smem -= 4096 + 8192 + (a.red2 ? 4096 : 0);      // dense rows never touch the grouped-max buffer or the SA metadata ring (both sit at the end)
    if (in_scale) {
        // a tile must not straddle two clouds: the kernel stages one cloud's scale/shift rows per tile
        CAPTRA_REQUIRE(rows_per_cloud % TC_ROWS == 0, "point_mlp_affine: rows_per_cloud must be a multiple of %d (got %d)", TC_ROWS, rows_per_cloud);
        CAPTRA_REQUIRE(cb == 0 && segB == nullptr, "point_mlp_affine: single input segment only");
        const int pad = round_up(ca, 16);
        const size_t extra = (size_t)2 * pad * sizeof(float);
        const size_t budget = a.small ? (size_t)(227 * 1024) / 2 - 2048 : (size_t)225 * 1024;
        CAPTRA_REQUIRE(smem + extra <= budget, "point_mlp_affine: %d input channels do not fit the shared-memory budget", ca);
        a.aff_pad = pad; smem += extra;
    }
    return tc_launch<1>(a, smem, stream);
}

int tc_point_mlp(int64_t rows, const float *segA, int64_t ldA, int ca, const float *segB, int64_t ldB, int cb,
                 int bcast, const captra_mlp_desc *d, const void *packed, float *y, int64_t ldy, int col_off,
                 int group, bool f16, cudaStream_t stream) {
    return tc_point_mlp_ex(rows, segA, ldA, ca, segB, ldB, cb, bcast, d, packed, y, ldy, col_off, group, f16, stream, nullptr, nullptr, 0);
}

// ------------------------------------------------------------------------------------------------
// GroupNorm statistics -> per-(cloud, channel) affine.  y is point-major [clouds*npts, C]; a group is
// `cpg` adjacent channels over all npts points of one cloud (blocks.py:73 GroupNorm(C/2, C)).
//   scale[b,c] = gamma[c] * rstd(b,g),  shift[b,c] = beta[c] - mean(b,g) * scale[b,c]
// so the consumer layer applies GroupNorm + ReLU as relu(y*scale + shift) while loading its input.
// One CTA per (cloud, 32-channel block): a warp reads 4 rows x 128 contiguous bytes per instruction, 8 loads in
// flight per thread (the op is a pure HBM stream: B*npts*C*4 bytes in, 2*B*C*4 out), fp32 partials per thread,
// fp64 combine.  512 CTAs for the RotationRegressor shapes, all resident at once.
// ------------------------------------------------------------------------------------------------
constexpr int GN_CB = 32;       // channels per CTA
constexpr int GN_TY = 32;       // rows in flight per pass (256 threads = 8 x 32)
constexpr int GN_UNROLL = 8;
__global__ void __launch_bounds__(256) group_norm_affine_kernel(int npts, int C, int cpg, const float *__restrict__ y, int64_t ld,
                                                                const float *__restrict__ gamma, const float *__restrict__ beta,
                                                                float eps, float *__restrict__ scale, float *__restrict__ shift) {
    __shared__ double s_sum[GN_TY][GN_CB], s_sq[GN_TY][GN_CB];
    const int b = blockIdx.y, cb = blockIdx.x * GN_CB;
    const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
    const int c0 = cb + tx * 4;
    float sm[4] = {0, 0, 0, 0}, sq[4] = {0, 0, 0, 0};
    if (c0 < C) {
        const float *base = y + ((size_t)b * npts) * ld + c0;
        const bool vec = (c0 + 4 <= C) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0) && ((ld & 3) == 0);
        if (vec) {
            int r = ty;
            for (; r + (GN_UNROLL - 1) * GN_TY < npts; r += GN_UNROLL * GN_TY) {
                float4 t[GN_UNROLL];
#pragma unroll
                for (int u = 0; u < GN_UNROLL; ++u) t[u] = __ldg(reinterpret_cast<const float4 *>(base + (size_t)(r + u * GN_TY) * ld));
#pragma unroll
                for (int u = 0; u < GN_UNROLL; ++u) {
                    sm[0] += t[u].x; sm[1] += t[u].y; sm[2] += t[u].z; sm[3] += t[u].w;
                    sq[0] = fmaf(t[u].x, t[u].x, sq[0]); sq[1] = fmaf(t[u].y, t[u].y, sq[1]);
                    sq[2] = fmaf(t[u].z, t[u].z, sq[2]); sq[3] = fmaf(t[u].w, t[u].w, sq[3]);
                }
            }
            for (; r < npts; r += GN_TY) {
                const float4 t = __ldg(reinterpret_cast<const float4 *>(base + (size_t)r * ld));
                sm[0] += t.x; sm[1] += t.y; sm[2] += t.z; sm[3] += t.w;
                sq[0] = fmaf(t.x, t.x, sq[0]); sq[1] = fmaf(t.y, t.y, sq[1]); sq[2] = fmaf(t.z, t.z, sq[2]); sq[3] = fmaf(t.w, t.w, sq[3]);
            }
        } else {
            for (int r = ty; r < npts; r += GN_TY) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float v = (c0 + j < C) ? __ldg(base + (size_t)r * ld + j) : 0.f;
                    sm[j] += v; sq[j] = fmaf(v, v, sq[j]);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { s_sum[ty][tx * 4 + j] = sm[j]; s_sq[ty][tx * 4 + j] = sq[j]; }
    __syncthreads();
    if (threadIdx.x < GN_CB) {
        double a = 0, q = 0;
        for (int i = 0; i < GN_TY; ++i) { a += s_sum[i][threadIdx.x]; q += s_sq[i][threadIdx.x]; }
        s_sum[0][threadIdx.x] = a; s_sq[0][threadIdx.x] = q;
    }
    __syncthreads();
    if (threadIdx.x < GN_CB) {
        const int c = cb + threadIdx.x;
        if (c < C) {
            const int g0 = (threadIdx.x / cpg) * cpg;   // GN_CB % cpg == 0 is checked on the host
            double a = 0, q = 0;
            for (int j = 0; j < cpg; ++j) { a += s_sum[0][g0 + j]; q += s_sq[0][g0 + j]; }
            const double n = (double)npts * cpg;
            const double mean = a / n;
            const double var = fmax(q / n - mean * mean, 0.0);
            const double rstd = 1.0 / sqrt(var + (double)eps);
            const double sc = (gamma ? (double)gamma[c] : 1.0) * rstd;
            scale[(size_t)b * C + c] = (float)sc;
            shift[(size_t)b * C + c] = (float)((beta ? (double)beta[c] : 0.0) - mean * sc);
        }
    }
}

// Second half of the fused GroupNorm statistics: stats [clouds * npts / 32][2][C] holds, per 32-row block, the column
// sums and sums of squares the producing GEMM's epilogue wrote (TcArgs::stats).  One CTA per (cloud, 32 channels)
// combines the npts / 32 blocks of its cloud in fp64 and emits the same per-(cloud, channel) affine as
// group_norm_affine_kernel -- 16 MB of partials instead of a second pass over the 268 MB activation.
__global__ void __launch_bounds__(256) group_norm_finalize_kernel(int nblk, int npts, int C, int cpg, const float *__restrict__ stats,
                                                                  const float *__restrict__ gamma, const float *__restrict__ beta,
                                                                  float eps, float *__restrict__ scale, float *__restrict__ shift) {
    __shared__ double s_sum[8][GN_CB], s_sq[8][GN_CB];
    const int b = blockIdx.y, cb = blockIdx.x * GN_CB;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = cb + tx;
    double a = 0, q = 0;
    if (c < C) {
        const float *base = stats + (size_t)b * nblk * 2 * C + c;
        int r = ty;
        for (; r + 24 < nblk; r += 32) {          // 8 loads in flight per thread
            float va[4], vq[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                va[u] = __ldg(base + (size_t)(r + 8 * u) * 2 * C);
                vq[u] = __ldg(base + (size_t)(r + 8 * u) * 2 * C + C);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) { a += (double)va[u]; q += (double)vq[u]; }
        }
        for (; r < nblk; r += 8) {
            a += (double)__ldg(base + (size_t)r * 2 * C);
            q += (double)__ldg(base + (size_t)r * 2 * C + C);
        }
    }
    s_sum[ty][tx] = a; s_sq[ty][tx] = q;
    __syncthreads();
    if (threadIdx.x < GN_CB) {
        double a2 = 0, q2 = 0;
        for (int i = 0; i < 8; ++i) { a2 += s_sum[i][threadIdx.x]; q2 += s_sq[i][threadIdx.x]; }
        s_sum[0][threadIdx.x] = a2; s_sq[0][threadIdx.x] = q2;
    }
    __syncthreads();
    if (threadIdx.x < GN_CB && c < C) {
        const int g0 = (threadIdx.x / cpg) * cpg;
        double a2 = 0, q2 = 0;
        for (int j = 0; j < cpg; ++j) { a2 += s_sum[0][g0 + j]; q2 += s_sq[0][g0 + j]; }
        const double n = (double)npts * cpg;
        const double mean = a2 / n;
        const double var = fmax(q2 / n - mean * mean, 0.0);
        const double rstd = 1.0 / sqrt(var + (double)eps);
        const double sc = (gamma ? (double)gamma[c] : 1.0) * rstd;
        scale[(size_t)b * C + c] = (float)sc;
        shift[(size_t)b * C + c] = (float)((beta ? (double)beta[c] : 0.0) - mean * sc);
    }
}

}  // namespace captra

using namespace captra;

extern "C" int captra_group_norm_affine(int clouds, int npts, int c, int channels_per_group, const float *y,
                                        int64_t ldy, const float *gamma, const float *beta, float eps,
                                        float *scale, float *shift, captra_stream_t stream) {
    CAPTRA_REQUIRE(clouds >= 0 && npts >= 1 && c >= 1, "group_norm_affine: bad sizes");
    CAPTRA_REQUIRE(channels_per_group >= 1 && GN_CB % channels_per_group == 0 && c % channels_per_group == 0,
                   "group_norm_affine: channels_per_group must divide %d and C", GN_CB);
    if (clouds == 0) return CAPTRA_OK;
    CAPTRA_REQUIRE(y && scale && shift, "group_norm_affine: null pointer");
    CAPTRA_REQUIRE(clouds <= 65535, "group_norm_affine: too many clouds");
    group_norm_affine_kernel<<<dim3(ceil_div(c, GN_CB), clouds), 256, 0, as_stream(stream)>>>(npts, c, channels_per_group, y, ldy, gamma,
                                                                                          beta, eps, scale, shift);
    CAPTRA_CHECK_LAUNCH("group_norm_affine");
    return CAPTRA_OK;
}

extern "C" int captra_group_norm_finalize(int clouds, int npts, int c, int channels_per_group, const float *stats,
                                          const float *gamma, const float *beta, float eps, float *scale, float *shift,
                                          captra_stream_t stream) {
    CAPTRA_REQUIRE(clouds >= 0 && npts >= 32 && npts % 32 == 0 && c >= 1, "group_norm_finalize: bad sizes (npts must be a multiple of 32)");
    CAPTRA_REQUIRE(channels_per_group >= 1 && GN_CB % channels_per_group == 0 && c % channels_per_group == 0,
                   "group_norm_finalize: channels_per_group must divide %d and C", GN_CB);
    if (clouds == 0) return CAPTRA_OK;
    CAPTRA_REQUIRE(stats && scale && shift, "group_norm_finalize: null pointer");
    CAPTRA_REQUIRE(clouds <= 65535, "group_norm_finalize: too many clouds");
    group_norm_finalize_kernel<<<dim3(ceil_div(c, GN_CB), clouds), 256, 0, as_stream(stream)>>>(npts / 32, npts, c, channels_per_group, stats,
                                                                                           gamma, beta, eps, scale, shift);
    CAPTRA_CHECK_LAUNCH("group_norm_finalize");
    return CAPTRA_OK;
}

namespace captra {
int tc_point_mlp_affine(int64_t rows, const float *x, int64_t ldx, int cin, const float *in_scale, const float *in_shift,
                        int rows_per_cloud, const captra_mlp_desc *d, const void *packed, float *y, int64_t ldy,
                        int col_off, bool f16, cudaStream_t stream) {
    return tc_point_mlp_ex(rows, x, ldx, cin, nullptr, 0, 0, 0, d, packed, y, ldy, col_off, 0, f16, stream, in_scale, in_shift, rows_per_cloud);
}
int tc_point_mlp_gnstats(int64_t rows, const float *x, int64_t ldx, int cin, const float *in_scale, const float *in_shift,
                         int rows_per_cloud, const captra_mlp_desc *d, const void *packed, float *y, int64_t ldy,
                         int col_off, float *stats, bool f16, cudaStream_t stream) {
    return tc_point_mlp_ex(rows, x, ldx, cin, nullptr, 0, 0, 0, d, packed, y, ldy, col_off, 0, f16, stream, in_scale, in_shift, rows_per_cloud, stats);
}
}  // namespace captra

extern "C" int captra_f16_overflow_flag(int reset) {
    int v = 0;
    CAPTRA_CUDA(cudaDeviceSynchronize());
    CAPTRA_CUDA(cudaMemcpyFromSymbol(&v, g_f16_overflow, sizeof(int)));
    if (reset) {
        int zero = 0;
        CAPTRA_CUDA(cudaMemcpyToSymbol(g_f16_overflow, &zero, sizeof(int)));
    }
    return v ? -1 : 0;
}

extern "C" int captra_debug_tc_timestamps(long long *out_host, int max_pairs) {
    int n = 0;
    CAPTRA_CUDA(cudaDeviceSynchronize());
    CAPTRA_CUDA(cudaMemcpyFromSymbol(&n, g_tc_ts_n, sizeof(int)));
    if (n > max_pairs) n = max_pairs;
    CAPTRA_CUDA(cudaMemcpyFromSymbol(out_host, g_tc_ts, sizeof(long long) * 2 * n));
    int zero = 0;
    CAPTRA_CUDA(cudaMemcpyToSymbol(g_tc_ts_n, &zero, sizeof(int)));
    return n;
}

extern "C" int captra_debug_umma_gemm(int k, int n, const float *A, const float *W, float *D, int terms,
                                      captra_stream_t stream) {
    CAPTRA_REQUIRE(k >= 8 && k % 8 == 0 && n >= 16 && n <= 256 && n % 16 == 0, "debug_umma_gemm: need K%%8==0, 16<=N<=256, N%%16==0");
    const size_t smem = (size_t)2 * (k / 4) * (TC_ROWS + n) * 16;
    CAPTRA_REQUIRE(smem <= 220 * 1024, "debug_umma_gemm: operands need %zu B of shared memory", smem);
    CAPTRA_CUDA(cudaFuncSetAttribute(umma_debug_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem + 1024));
    umma_debug_gemm_kernel<<<1, 128, smem, as_stream(stream)>>>(k, n, A, W, D, terms);
    CAPTRA_CHECK_LAUNCH("debug_umma_gemm");
    return CAPTRA_OK;
}
