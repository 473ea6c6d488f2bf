// mlp_tc.cu -- fused shared per-point MLP on the 5th-gen tensor cores (impl 1): tcgen05.mma with
// TMEM accumulators, operands staged in shared memory (weights by bulk-copy TMA), 3xTF32 split
// arithmetic so the result matches true fp32 to ~1e-6 (the reference's convs are true fp32;
// single-pass TF32 misses the 1e-4 pose tolerance, SURVEY section 7 "hard parts").
#include "mlp_common.cuh"
#include "tc_common.cuh"

#include <math.h>
#include <stdlib.h>

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
// Persistent CTAs walk 128-row tiles.  Warps 0-3 ("row threads", thread r <-> tile row r <-> TMEM
// lane r) produce the A operand and run the epilogues; warp 4 (one elected thread) issues
// tcgen05.mma; warp 5 (one elected thread) streams the pre-split weights with bulk-copy TMA.
//
//   layer l, K slab s (16 input channels), stage st = slab counter mod NST:
//     row threads : A slab -> (hi, lo) planes of a_stage[st]
//         layer 0 : gathered from global through the index list (SA) or row pointers (dense), with a
//                   3-slab register prefetch so the L2 gather latency hides behind the MMAs
//         layer>0 : read straight out of TMEM -- 16 accumulator columns of layer l-1 ARE slab s of
//                   layer l -- + bias, ReLU, 3xTF32 split.  The epilogue of layer l-1 and the MMAs
//                   of layer l therefore overlap slab by slab; accumulators ping-pong between two
//                   TMEM regions and no activation ever touches shared memory in fp32.
//     TMA thread  : W slab (hi plane | lo plane, pre-split, chunk-major in global) -> w_stage[st]
//     MMA thread  : 2 k-steps x 3 terms (lo*hi, hi*lo, hi*hi) of tcgen05.mma kind::tf32,
//                   tcgen05.commit -> empty[st]; after a layer's last slab also -> d_ready
//   last epilogue : group > 0 -> max over the rows of each group by redux.sync on the (non-negative)
//                                float bit patterns, one coalesced row write per group
//                   group = 0 -> the thread's row straight to global (point-major)
//
// Shared-memory operand layout (K-major, no swizzle): element (row, k) of a 16-channel slab lives at
//   plane + (k/4) * ROWS*16 + row*16 + (k%4)*4      -> LBO = ROWS*16, SBO = 128 in the descriptors,
// so thread r writes 16-byte pieces at r*16 (conflict-free) and never touches another row.
// ------------------------------------------------------------------------------------------------
constexpr int TC_GROUPS = 2;                      // producer groups; group g owns every 2nd slab
constexpr int TC_PROD = 128 * TC_GROUPS;          // producer / epilogue threads (warps 0-7)
constexpr int TC_THREADS = TC_PROD + 64;          // + MMA warp (8) + TMA warp (9)
// A slab is always 8 chunks of 16 bytes per row = 4 k-steps; it spans 32 channels as TF32 (impl 1,
// 3xTF32) or 64 channels as fp16 (impl 2, "fp16x3": the same hi/lo split with 11-bit mantissas, half
// the operand bytes and K=16 per MMA; weights are pre-scaled by a power of two per layer so their lo
// parts stay normal, activations saturate at +-65504).
constexpr int TC_NCHUNK = 8;
constexpr int TC_BIAS_PAD = 16;                   // bias region = npad + 16 floats; [npad] = 1 / weight scale
__host__ __device__ constexpr int tc_kc(bool f16, bool small) { return (f16 ? 64 : 32) / (small ? 2 : 1); }
constexpr int TC_A_PLANE = TC_NCHUNK * TC_CHUNK_BYTES;   // bytes of one (hi or lo) A slab plane: 16 KB
constexpr int TC_A_STAGE = 2 * TC_A_PLANE;        // hi + lo
constexpr int TC_MAX_STAGES = 4;

struct TcArgs {
    int nlayers, relu_last, cout_last;
    int kpad[CAPTRA_MAX_MLP_LAYERS], npad[CAPTRA_MAX_MLP_LAYERS];
    const float *wpk[CAPTRA_MAX_MLP_LAYERS];   // [nslab][2 planes][8 chunks][npad][4]
    const float *bias[CAPTRA_MAX_MLP_LAYERS];  // [npad]
    int64_t rows, ntiles;
    int group;                                  // max over each `group` rows (32|64|128) or 0
    float *out; int64_t ldo; int col_off;
    int n, s, cfeat; const float *xyz, *new_xyz, *feats; const int *idx;
    const float *segA; int64_t ldA; int ca; const float *segB; int64_t ldB; int cb; int bcast;
    const float *in_scale, *in_shift; int rows_per_cloud;   // dense: x <- relu(x * scale[cloud] + shift[cloud]) on load (GroupNorm + ReLU of the producer layer)
    int wstage_bytes, bias_floats, region_cols, tmem_cols, nst_log2;
    int nsplit, last_npad, cout_total;          // single wide layer split into 256-column chunks over grid.y
    int f16, small;
    int cin0, klast0;                           // true layer-0 width; k-steps that carry data in layer 0's last slab
    int dbg;                                    // timing probes (CAPTRA_TC_DBG): 1 no A stores, 2 no W copies, 4 no MMAs, 8 no last epilogue, 32/128 stamps, 64 no gather
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
// x -> (hi, lo) as two fp16 numbers (hi = rn(x), lo = rn(x - hi)): 22 mantissa bits together
__device__ __forceinline__ void split_f16(float x, unsigned short &hi, unsigned short &lo) {
    unsigned short h, l;
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(h) : "f"(x));
    float hf;
    asm("cvt.f32.f16 %0, %1;" : "=f"(hf) : "h"(h));
    asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(l) : "f"(x - hf));
    hi = h; lo = l;
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
__device__ __forceinline__ uint32_t pack2(unsigned short a, unsigned short b) { return (uint32_t)a | ((uint32_t)b << 16); }

// timing probe: producer thread 0 of CTA 0 stamps clock64() at phase boundaries of its first tiles
__device__ long long g_tc_ts[512];
__device__ int g_tc_ts_n;
#define TC_STAMP_T(code, T)                                                         \
    do {                                                                            \
        if ((a.dbg & (T)) && blockIdx.x == 0 && blockIdx.y == 0) {                  \
            const int i__ = g_tc_ts_n;                                              \
            if (i__ < 255) { g_tc_ts[2 * i__] = clock64(); g_tc_ts[2 * i__ + 1] = (code); g_tc_ts_n = i__ + 1; } \
        }                                                                           \
    } while (0)
#define TC_STAMP(code)                                                              \
    do {                                                                            \
        if ((a.dbg & 32) && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0) {       \
            const int i__ = g_tc_ts_n;                                              \
            if (i__ < 255) { g_tc_ts[2 * i__] = clock64(); g_tc_ts[2 * i__ + 1] = (code); g_tc_ts_n = i__ + 1; } \
        }                                                                           \
    } while (0)

// What bounds this kernel (clock64 probes, profiles/r01_tc_probe.txt): the single MMA-issuing thread.
// One tcgen05.mma costs it ~80 cycles to issue and one barrier round (two try_waits + commit) ~450,
// while a kind::tf32 MMA of N=128 is only 64 tensor-pipe cycles.  Hence: 32-channel slabs (12 MMAs
// per barrier round), ONE `full` barrier per stage shared by the A producers and the weight TMA
// (4 warp arrivals + 1 arrive.expect_tx), descriptors advanced by adding to their low word, and
// two producer groups that alternate slabs so their per-slab latency chains overlap.
//
// Two shapes of the same kernel.  LARGE (wide layers): 320 threads, two producer groups alternating
// 8-chunk slabs, one CTA per SM.  SMALL (every layer <= 128 columns: the sa1 scales, fp1): the tile
// time is a chain of latencies (gather, MMA drain, TMEM read, epilogue), not throughput, so the CTA
// shrinks to one producer group and 4-chunk slabs (192 threads, 64 KB, 256 TMEM columns) and TWO CTAs
// share an SM, each on its own tile.
template <int MODE, bool F16, bool SMALL>  // MODE 0: SA gather loader, 1: dense-row loader; the last epilogue is chosen by a.group
__global__ void __launch_bounds__(SMALL ? 192 : 320, SMALL ? 2 : 1) mlp_tc_kernel(const TcArgs a) {
    constexpr int TC_GROUPS = SMALL ? 1 : 2;
    constexpr int TC_PROD = 128 * TC_GROUPS;
    constexpr int TC_THREADS = TC_PROD + 64;
    constexpr int TC_NCHUNK = SMALL ? 4 : 8;
    constexpr int TC_A_PLANE = TC_NCHUNK * TC_CHUNK_BYTES;
    constexpr int TC_A_STAGE = 2 * TC_A_PLANE;
    constexpr int TC_KC = tc_kc(F16, SMALL);
    constexpr int PIECES = TC_KC / 16;             // 16-channel pieces per slab
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t full[TC_MAX_STAGES], empty[TC_MAX_STAGES], d_ready;
    __shared__ uint32_t tmem_base_s;

    const uint32_t nst_log2 = (uint32_t)a.nst_log2, NST = 1u << nst_log2;
    uint8_t *a_stage = smem_raw;                                          // NST x TC_A_STAGE
    uint8_t *w_stage = a_stage + (size_t)NST * TC_A_STAGE;                // NST x wstage_bytes
    float *bias_s = reinterpret_cast<float *>(w_stage + (size_t)NST * a.wstage_bytes);
    float *red = reinterpret_cast<float *>(a_stage);                      // [4][256], aliases stage 0 (idle in the last epilogue)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // a single wide layer is split over grid.y: this CTA owns output columns [256*y, 256*y + npad)
    int npad_y = 0, col_off = a.col_off, cout_last = a.cout_last;
    const float *wpk0 = a.wpk[0], *bias0 = a.bias[0];
    if (a.nsplit > 1) {
        const int y = blockIdx.y;
        npad_y = (y < a.nsplit - 1) ? 256 : a.last_npad;
        wpk0 += (size_t)y * (a.kpad[0] / TC_KC) * 2 * TC_NCHUNK * 256 * 4;
        bias0 += y * (256 + TC_BIAS_PAD);
        col_off += y * 256;
        cout_last = min(256, a.cout_total - y * 256);
    }
    auto NPAD = [&](int l) { return a.nsplit > 1 ? npad_y : a.npad[l]; };
    auto WPK = [&](int l) { return l == 0 ? wpk0 : a.wpk[l]; };
    auto BIAS = [&](int l) { return l == 0 ? bias0 : a.bias[l]; };

    if (warp == TC_PROD / 32) tmem_alloc_dyn(&tmem_base_s, (uint32_t)a.tmem_cols);
    if (tid == 0) {
        for (uint32_t i = 0; i < NST; ++i) { mbar_init(&full[i], 4 + 1); mbar_init(&empty[i], 1); }
        mbar_init(&d_ready, 1);
        fence_mbar_init();
    }
    {   // biases of all layers -> smem
        int off = 0;
        for (int l = 0; l < a.nlayers; ++l) {
            for (int i = tid; i < NPAD(l) + TC_BIAS_PAD; i += TC_THREADS) bias_s[off + i] = __ldg(BIAS(l) + i);
            off += NPAD(l) + TC_BIAS_PAD;
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (tid == TC_PROD) {
        // ===================== MMA issuer =====================
        uint32_t it = 0;
        const uint64_t a_desc0 = smem_desc_kmajor_noswz(smem_u32(a_stage), TC_CHUNK_BYTES, 128);
        for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
            for (int l = 0; l < a.nlayers; ++l) {
                const int nslab = a.kpad[l] / TC_KC;
                const uint32_t npad = (uint32_t)NPAD(l);
                const uint32_t idesc = make_idesc(F16 ? 0 : 2, TC_ROWS, (int)npad);
                const uint32_t b_lbo = npad * 16;
                const uint64_t b_desc0 = smem_desc_kmajor_noswz(smem_u32(w_stage), b_lbo, 128);
                const uint32_t b_lo_off = (TC_NCHUNK * b_lbo) >> 4, b_step = (2 * b_lbo) >> 4;   // in 16-byte units
                const uint32_t tmem_d = tmem_base + (uint32_t)((l & 1) * a.region_cols);
                for (int s = 0; s < nslab; ++s, ++it) {
                    const uint32_t st = it & (NST - 1), ph = (it >> nst_log2) & 1;
                    TC_STAMP_T(40, 128);
                    mbar_wait(&full[st], ph);
                    tcgen05_fence_after();
                    TC_STAMP_T(42, 128);
                    // descriptors of this stage: only the 14-bit start-address field (16-byte units) moves
                    uint64_t ah = a_desc0 + (uint64_t)((st * TC_A_STAGE) >> 4);
                    uint64_t al = ah + (TC_A_PLANE >> 4);
                    uint64_t bh = b_desc0 + (uint64_t)((st * (uint32_t)a.wstage_bytes) >> 4);
                    uint64_t bl = bh + b_lo_off;
                    const int nks = (l == 0 && s == nslab - 1) ? a.klast0 : TC_NCHUNK / 2;   // zero padding needs no MMAs
#pragma unroll
                    for (int j = 0; j < TC_NCHUNK / 2; ++j) {
                        if ((a.dbg & 4) || j >= nks) break;
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
                    TC_STAMP_T(43, 128);
                    umma_commit(&empty[st]);
                    if (s == nslab - 1) umma_commit(&d_ready);
                }
            }
        }
    } else if (tid == TC_PROD + 32) {
        // ===================== weight producer (bulk-copy TMA) =====================
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
            for (int l = 0; l < a.nlayers; ++l) {
                const int nslab = a.kpad[l] / TC_KC;
                const uint32_t bytes = 2u * TC_NCHUNK * (uint32_t)NPAD(l) * 16u;
                const uint8_t *src = reinterpret_cast<const uint8_t *>(WPK(l));
                for (int s = 0; s < nslab; ++s, ++it) {
                    const uint32_t st = it & (NST - 1), ph = (it >> nst_log2) & 1;
                    mbar_wait(&empty[st], ph ^ 1);
                    if (a.dbg & 2) {
                        mbar_arrive(&full[st]);
                    } else {
                        mbar_arrive_expect_tx(&full[st], bytes);
                        bulk_g2s(w_stage + (size_t)st * a.wstage_bytes, src + (size_t)s * bytes, bytes, &full[st]);
                    }
                }
            }
        }
    } else if (tid < TC_PROD) {
        // ===================== producer / epilogue threads =====================
        const int r = tid & 127;              // tile row == TMEM lane
        const int g = tid >> 7;               // producer group
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        uint32_t base = 0;                    // global index of the current layer's slab 0
        uint32_t dl = 0;
        float amax = 0.f;                     // largest |operand| this thread converted to fp16

        auto warp_wait = [&](uint64_t *bar, uint32_t parity) {   // one lane polls, the warp follows
            if (lane == 0) mbar_wait(bar, parity);
            __syncwarp();
        };
        // store 16 channels (piece `pc` of the slab) of global slab `it` as hi/lo operand planes
        auto store16 = [&](uint32_t it, int pc, const float4 (&v)[4]) {
            if (a.dbg & 1) return;
            const uint32_t st = it & (NST - 1);
            float *plane_hi = reinterpret_cast<float *>(a_stage + st * TC_A_STAGE) + r * 4;
            float *plane_lo = plane_hi + TC_A_PLANE / 4;
            if (F16) {        // 2 chunks of 8 halfs
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const float e[8] = {v[2 * q].x, v[2 * q].y, v[2 * q].z, v[2 * q].w, v[2 * q + 1].x, v[2 * q + 1].y, v[2 * q + 1].z, v[2 * q + 1].w};
                    unsigned short hh[8], ll[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        split_f16(e[j], hh[j], ll[j]);
                        amax = fmaxf(amax, fabsf(e[j]));      // saturation watch, checked once per tile
                    }
                    const int chunk = 2 * pc + q;
                    *reinterpret_cast<uint4 *>(plane_hi + chunk * (TC_ROWS * 4)) =
                        make_uint4(pack2(hh[0], hh[1]), pack2(hh[2], hh[3]), pack2(hh[4], hh[5]), pack2(hh[6], hh[7]));
                    *reinterpret_cast<uint4 *>(plane_lo + chunk * (TC_ROWS * 4)) =
                        make_uint4(pack2(ll[0], ll[1]), pack2(ll[2], ll[3]), pack2(ll[4], ll[5]), pack2(ll[6], ll[7]));
                }
            } else {          // 4 chunks of 4 floats
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float4 hh, ll;
                    split_tf32(v[q].x, hh.x, ll.x); split_tf32(v[q].y, hh.y, ll.y);
                    split_tf32(v[q].z, hh.z, ll.z); split_tf32(v[q].w, hh.w, ll.w);
                    const int chunk = 4 * pc + q;
                    *reinterpret_cast<float4 *>(plane_hi + chunk * (TC_ROWS * 4)) = hh;
                    *reinterpret_cast<float4 *>(plane_lo + chunk * (TC_ROWS * 4)) = ll;
                }
            }
        };
        auto acquire = [&](uint32_t it) {     // wait until the MMAs that last read this slab's stage are done
            warp_wait(&empty[it & (NST - 1)], ((it >> nst_log2) & 1) ^ 1);
        };
        auto release = [&](uint32_t it) {     // hand the filled stage to the MMA thread
            // every writer fences its own generic-proxy stores towards the async proxy; one elected
            // lane per warp then arrives
            fence_proxy_async_smem();
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[it & (NST - 1)]);
        };

        for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
            const int64_t grow = tile * TC_ROWS + r;
            const bool valid = grow < a.rows;
            TC_STAMP(0);
            // ---- per-row metadata
            const float *frow = nullptr;         // SA: feature row of the gathered point
            float px = 0.f, py = 0.f, pz = 0.f;  // SA: point - centroid
            const float *arow = nullptr, *brow = nullptr;
            if (valid) {
                if (MODE == 0) {
                    const int64_t cen = grow / a.group;
                    const int64_t b = cen / a.s;
                    const int64_t p = b * a.n + __ldg(a.idx + grow);
                    frow = a.feats ? a.feats + p * a.cfeat : nullptr;
                    px = __fsub_rn(__ldg(a.xyz + p * 3 + 0), __ldg(a.new_xyz + cen * 3 + 0));
                    py = __fsub_rn(__ldg(a.xyz + p * 3 + 1), __ldg(a.new_xyz + cen * 3 + 1));
                    pz = __fsub_rn(__ldg(a.xyz + p * 3 + 2), __ldg(a.new_xyz + cen * 3 + 2));
                } else {
                    arow = a.segA ? a.segA + grow * a.ldA : nullptr;
                    brow = a.segB ? a.segB + (a.bcast ? grow / a.bcast : grow) * a.ldB : nullptr;
                }
            }
            const float *nsc = nullptr, *nsh = nullptr;   // per-cloud affine of the input (dense mode)
            if (MODE == 1 && a.in_scale && valid) {
                const int64_t cloud = grow / a.rows_per_cloud;
                nsc = a.in_scale + cloud * a.ca;
                nsh = a.in_shift + cloud * a.ca;
            }
            auto in0 = [&](int c) -> float {     // layer-0 input element c of this row
                if (!valid) return 0.f;
                if (MODE == 0) {
                    if (c < a.cfeat) return __ldg(frow + c);
                    const int e = c - a.cfeat;
                    return e == 0 ? px : (e == 1 ? py : (e == 2 ? pz : 0.f));
                } else {
                    if (c < a.ca) return __ldg(arow + c);
                    if (c < a.ca + a.cb) return __ldg(brow + (c - a.ca));
                    return 0.f;
                }
            };
            const int nvec = MODE == 0 ? a.cfeat : a.ca;           // leading segment length
            const float *vbase = MODE == 0 ? frow : arow;
            const bool vec_row = valid && vbase && ((reinterpret_cast<uintptr_t>(vbase) & 15) == 0);
            auto load16 = [&](int c0, float4 (&v)[4]) {             // 16 consecutive layer-0 channels
                if (c0 >= a.cin0) {   // pure padding
#pragma unroll
                    for (int q = 0; q < 4; ++q) v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                    return;
                }
                if (a.dbg & 64) {   // probe: no global gather
#pragma unroll
                    for (int q = 0; q < 4; ++q) v[q] = make_float4(px, py, pz, 1.f);
                    return;
                }
                if (vec_row && c0 + 16 <= nvec) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) v[q] = __ldg(reinterpret_cast<const float4 *>(vbase + c0) + q);
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        v[q] = make_float4(in0(c0 + 4 * q), in0(c0 + 4 * q + 1), in0(c0 + 4 * q + 2), in0(c0 + 4 * q + 3));
                }
                if (MODE == 1 && nsc) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float *e = reinterpret_cast<float *>(&v[q]);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int c = c0 + 4 * q + j;
                            e[j] = c < a.ca ? fmaxf(fmaf(e[j], __ldg(nsc + c), __ldg(nsh + c)), 0.f) : 0.f;
                        }
                    }
                }
            };
            TC_STAMP(1);
            // ---------------- layer 0: this group's slabs; each piece's registers are refilled with the
            // same piece of the group's NEXT slab right after it is stored (one-slab-ahead prefetch) ----
            {
                const int nslab = a.kpad[0] / TC_KC;
                int s = (int)((g - base) & (TC_GROUPS - 1));      // first slab of this layer owned by group g
                float4 cur[PIECES][4];
                if (s < nslab) {
#pragma unroll
                    for (int pc = 0; pc < PIECES; ++pc) load16(s * TC_KC + 16 * pc, cur[pc]);
                }
                for (; s < nslab; s += TC_GROUPS) {
                    const bool more = s + TC_GROUPS < nslab;
                    acquire(base + s);
#pragma unroll
                    for (int pc = 0; pc < PIECES; ++pc) {
                        store16(base + s, pc, cur[pc]);
                        if (more) load16((s + TC_GROUPS) * TC_KC + 16 * pc, cur[pc]);
                    }
                    release(base + s);
                }
                base += nslab;
            }
            TC_STAMP(2);
            // ---------------- layers 1..L-1: previous accumulator -> next operand ----------------
            int bias_off = 0;
            for (int l = 1; l < a.nlayers; ++l) {
                const int nslab = a.kpad[l] / TC_KC;          // == NPAD(l-1) / TC_KC
                const uint32_t tsrc = tmem_base + (uint32_t)(((l - 1) & 1) * a.region_cols) + lane_addr;
                const float *bz = bias_s + bias_off;
                const float inv = bz[NPAD(l - 1)];            // 1 / weight scale of layer l-1
                int s = (int)((g - base) & (TC_GROUPS - 1));
                warp_wait(&d_ready, dl & 1);
                ++dl;
                tcgen05_fence_after();
                TC_STAMP(10 + l);
                uint32_t v[16];
                if (s < nslab) tmem_ld_32x16(tsrc + (uint32_t)(TC_KC * s), v);
                for (; s < nslab; s += TC_GROUPS) {
                    acquire(base + s);
#pragma unroll
                    for (int pc = 0; pc < PIECES; ++pc) {
                        tmem_ld_wait();
                        float4 x[4];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 b4 = *reinterpret_cast<const float4 *>(bz + s * TC_KC + 16 * pc + 4 * q);
                            x[q].x = fmaxf(fmaf(__uint_as_float(v[4 * q + 0]), inv, b4.x), 0.f);
                            x[q].y = fmaxf(fmaf(__uint_as_float(v[4 * q + 1]), inv, b4.y), 0.f);
                            x[q].z = fmaxf(fmaf(__uint_as_float(v[4 * q + 2]), inv, b4.z), 0.f);
                            x[q].w = fmaxf(fmaf(__uint_as_float(v[4 * q + 3]), inv, b4.w), 0.f);
                        }
                        // overlap the next TMEM read: next piece of this slab, else the group's next slab
                        if (pc + 1 < PIECES) tmem_ld_32x16(tsrc + (uint32_t)(TC_KC * s + 16 * (pc + 1)), v);
                        else if (s + TC_GROUPS < nslab) tmem_ld_32x16(tsrc + (uint32_t)(TC_KC * (s + TC_GROUPS)), v);
                        store16(base + s, pc, x);
                    }
                    release(base + s);
                }
                base += nslab;
                bias_off += NPAD(l - 1) + TC_BIAS_PAD;
                TC_STAMP(20 + l);
            }
            // ---------------- last epilogue: the groups alternate 16-column chunks ----------------
            {
                const int l = a.nlayers - 1;
                const int npad = NPAD(l);
                const uint32_t tsrc = tmem_base + (uint32_t)((l & 1) * a.region_cols) + lane_addr;
                warp_wait(&d_ready, dl & 1);
                ++dl;
                tcgen05_fence_after();
                TC_STAMP(30);
                const float inv = bias_s[bias_off + npad];   // 1 / weight scale of the last layer
                uint32_t v[16];
                int c0 = 16 * g;
                if (c0 < npad) tmem_ld_32x16(tsrc + (uint32_t)c0, v);
                for (; c0 < npad; c0 += 16 * TC_GROUPS) {
                    tmem_ld_wait();
                    if (a.dbg & 8) break;
                    float x[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float t = fmaf(__uint_as_float(v[j]), inv, bias_s[bias_off + c0 + j]);
                        x[j] = a.relu_last ? fmaxf(t, 0.f) : t;
                    }
                    if (c0 + 16 * TC_GROUPS < npad) tmem_ld_32x16(tsrc + (uint32_t)(c0 + 16 * TC_GROUPS), v);
                    if (a.group > 0) {
                        // max over the 32 rows of this warp for each of the 16 columns
                        if (!valid) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) x[j] = -INFINITY;
                        }
                        int col;
                        const float m = warp_colmax16(x, lane, col);
                        if (lane < 16) red[(warp & 3) * 256 + c0 + col] = m;
                    } else if (valid) {
                        float *dst = a.out + grow * a.ldo + col_off + c0;
                        if (c0 + 16 <= cout_last && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
                            for (int q = 0; q < 4; ++q)
                                *reinterpret_cast<float4 *>(dst + 4 * q) = make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (c0 + j < cout_last) dst[j] = x[j];
                        }
                    }
                }
                if (a.group > 0) {
                    // combine the per-warp maxima of the warps that share a group and write them out
                    const int wpg = a.group / 32;                 // row-warps per group: 1, 2 or 4
                    asm volatile("bar.sync 1, %0;" ::"n"(TC_PROD) : "memory");
                    const int ngroups = 4 / wpg;
                    for (int o = tid; o < ngroups * npad; o += TC_PROD) {
                        const int gi = o / npad, c = o - gi * npad;
                        const int64_t cen = (tile * TC_ROWS) / a.group + gi;
                        if (c >= cout_last || cen * a.group >= a.rows) continue;
                        float m = red[(gi * wpg) * 256 + c];
                        for (int w = 1; w < wpg; ++w) m = fmaxf(m, red[(gi * wpg + w) * 256 + c]);
                        a.out[cen * a.ldo + col_off + c] = m;
                    }
                    asm volatile("bar.sync 1, %0;" ::"n"(TC_PROD) : "memory");   // red aliases a_stage[0]
                }
                tcgen05_fence_before();
                TC_STAMP(31);
            }
        }
        if (F16 && !(amax < 65000.f)) g_f16_overflow = 1;   // also catches NaN
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == TC_PROD / 32) tmem_dealloc_dyn(tmem_base, (uint32_t)a.tmem_cols);
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
    int nlayers, kpad[CAPTRA_MAX_MLP_LAYERS], npad[CAPTRA_MAX_MLP_LAYERS];
    size_t off_w[CAPTRA_MAX_MLP_LAYERS], off_b[CAPTRA_MAX_MLP_LAYERS], total_floats;
    int wstage_bytes, bias_floats, nsplit, last_npad;
    int nstages, region_cols, tmem_cols, target_occ;
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
        L.kpad[l] = round_up(cin, TC_KC);
        // a non-last layer's accumulator columns are the next layer's K: pad to whole slabs
        L.npad[l] = round_up(d.cout[l], l < d.nlayers - 1 ? TC_KC : 16);
        if (L.npad[l] > 256) {
            if (d.nlayers == 1) {   // one wide layer: 256-column chunks over grid.y
                L.nsplit = ceil_div(L.npad[l], 256);
                L.last_npad = L.npad[l] - 256 * (L.nsplit - 1);
            } else {
                L.supported = false;
            }
        }
        npmax = max(npmax, min(L.npad[l], 256));
        // packed weights of a layer: nslab * 2 planes * 8 chunks * npad * 16 bytes (in floats: /4)
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
    const size_t fixed = (size_t)L.bias_floats * 4 + 256;
    const size_t per_stage = (size_t)TC_A_STAGE + L.wstage_bytes;
    const size_t budget = small ? (size_t)(227 * 1024) / 2 - 2048 : (size_t)225 * 1024;
    L.nstages = (!small && fixed + 4 * per_stage <= budget) ? 4 : 2;
    L.smem_bytes = fixed + (size_t)L.nstages * per_stage;
    if (L.smem_bytes > budget) L.supported = false;
    if (small && (d.nlayers < 2 || npmax > 128)) L.supported = false;
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
    a.nsplit = L.nsplit; a.last_npad = L.last_npad; a.cout_total = d->cout[d->nlayers - 1];
    {
        const int kc = tc_kc(f16, L.small), kmma = f16 ? 16 : 8;
        a.cin0 = d->cin;
        const int rem = d->cin - (L.kpad[0] / kc - 1) * kc;     // channels in layer 0's last slab
        a.klast0 = ceil_div(rem, kmma);
    }
    { const char *e = getenv("CAPTRA_TC_DBG"); a.dbg = e ? atoi(e) : 0; }
    *smem = L.smem_bytes;
    return CAPTRA_OK;
}

template <int MODE, bool F16, bool SMALL>
static int tc_launch_t(TcArgs &a, size_t smem, cudaStream_t stream) {
    auto kern = mlp_tc_kernel<MODE, F16, SMALL>;
    CAPTRA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CAPTRA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    a.ntiles = ceil_div<int64_t>(a.rows, TC_ROWS);
    const int64_t slots = (int64_t)max(1, sm_count() / a.nsplit) * (SMALL ? 2 : 1);
    const int gx = (int)(a.ntiles < slots ? a.ntiles : slots);
    kern<<<dim3(gx, a.nsplit), SMALL ? 192 : 320, smem, stream>>>(a);
    CAPTRA_CHECK_LAUNCH("mlp_tc");
    return CAPTRA_OK;
}
template <int MODE>
static int tc_launch(TcArgs &a, size_t smem, cudaStream_t stream) {
    if (a.small) return a.f16 ? tc_launch_t<MODE, true, true>(a, smem, stream) : tc_launch_t<MODE, false, true>(a, smem, stream);
    return a.f16 ? tc_launch_t<MODE, true, false>(a, smem, stream) : tc_launch_t<MODE, false, false>(a, smem, stream);
}

int tc_sa_mlp_max(int b, int n, int s, int k, int cfeat, const float *xyz, const float *new_xyz,
                  const float *feats, const int *idx, const captra_mlp_desc *d, const void *packed,
                  float *out, int64_t ldo, int col_off, bool f16, cudaStream_t stream) {
    CAPTRA_REQUIRE(k == 32 || k == 64 || k == 128, "sa_mlp_max(tc): nsample must be 32, 64 or 128 (got %d)", k);
    CAPTRA_REQUIRE(d->relu_last, "sa_mlp_max(tc): the max epilogue needs a ReLU after the last layer");
    TcArgs a{};
    size_t smem;
    int rc = tc_fill(a, d, packed, f16, &smem);
    if (rc) return rc;
    a.rows = (int64_t)b * s * k; a.group = k;
    a.out = out; a.ldo = ldo; a.col_off = col_off;
    a.n = n; a.s = s; a.cfeat = cfeat; a.xyz = xyz; a.new_xyz = new_xyz; a.feats = feats; a.idx = idx;
    return tc_launch<0>(a, smem, stream);
}

static int tc_point_mlp_ex(int64_t rows, const float *segA, int64_t ldA, int ca, const float *segB, int64_t ldB, int cb,
                           int bcast, const captra_mlp_desc *d, const void *packed, float *y, int64_t ldy, int col_off,
                           int group, bool f16, cudaStream_t stream, const float *in_scale, const float *in_shift,
                           int rows_per_cloud) {
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
// One CTA per (cloud, 64-channel block): coalesced 256-byte row segments, fp32 partials per thread,
// fp64 combine.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) group_norm_affine_kernel(int npts, int C, int cpg, const float *__restrict__ y, int64_t ld,
                                                                const float *__restrict__ gamma, const float *__restrict__ beta,
                                                                float eps, float *__restrict__ scale, float *__restrict__ shift) {
    __shared__ double s_sum[16][64], s_sq[16][64];
    const int b = blockIdx.y, cb = blockIdx.x * 64;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int c0 = cb + tx * 4;
    float sm[4] = {0, 0, 0, 0}, sq[4] = {0, 0, 0, 0};
    if (c0 < C) {
        const float *base = y + ((size_t)b * npts) * ld + c0;
        const bool vec = (c0 + 4 <= C) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0) && ((ld & 3) == 0);
        for (int r = ty; r < npts; r += 16) {
            float v[4];
            if (vec) {
                const float4 t = __ldg(reinterpret_cast<const float4 *>(base + (size_t)r * ld));
                v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) v[j] = (c0 + j < C) ? __ldg(base + (size_t)r * ld + j) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) { sm[j] += v[j]; sq[j] = fmaf(v[j], v[j], sq[j]); }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) { s_sum[ty][tx * 4 + j] = sm[j]; s_sq[ty][tx * 4 + j] = sq[j]; }
    __syncthreads();
    if (threadIdx.x < 64) {
        double a = 0, q = 0;
        for (int i = 0; i < 16; ++i) { a += s_sum[i][threadIdx.x]; q += s_sq[i][threadIdx.x]; }
        s_sum[0][threadIdx.x] = a; s_sq[0][threadIdx.x] = q;
    }
    __syncthreads();
    if (threadIdx.x < 64) {
        const int c = cb + threadIdx.x;
        if (c < C) {
            const int g0 = (threadIdx.x / cpg) * cpg;   // 64 % cpg == 0 is checked on the host
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

}  // namespace captra

using namespace captra;

extern "C" int captra_group_norm_affine(int clouds, int npts, int c, int channels_per_group, const float *y,
                                        int64_t ldy, const float *gamma, const float *beta, float eps,
                                        float *scale, float *shift, captra_stream_t stream) {
    CAPTRA_REQUIRE(clouds >= 0 && npts >= 1 && c >= 1, "group_norm_affine: bad sizes");
    CAPTRA_REQUIRE(channels_per_group >= 1 && 64 % channels_per_group == 0 && c % channels_per_group == 0,
                   "group_norm_affine: channels_per_group must divide 64 and C");
    if (clouds == 0) return CAPTRA_OK;
    CAPTRA_REQUIRE(y && scale && shift, "group_norm_affine: null pointer");
    CAPTRA_REQUIRE(clouds <= 65535, "group_norm_affine: too many clouds");
    group_norm_affine_kernel<<<dim3(ceil_div(c, 64), clouds), 256, 0, as_stream(stream)>>>(npts, c, channels_per_group, y, ldy, gamma,
                                                                                          beta, eps, scale, shift);
    CAPTRA_CHECK_LAUNCH("group_norm_affine");
    return CAPTRA_OK;
}

namespace captra {
int tc_point_mlp_affine(int64_t rows, const float *x, int64_t ldx, int cin, const float *in_scale, const float *in_shift,
                        int rows_per_cloud, const captra_mlp_desc *d, const void *packed, float *y, int64_t ldy,
                        int col_off, bool f16, cudaStream_t stream) {
    return tc_point_mlp_ex(rows, x, ldx, cin, nullptr, 0, 0, 0, d, packed, y, ldy, col_off, 0, f16, stream, in_scale, in_shift, rows_per_cloud);
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
