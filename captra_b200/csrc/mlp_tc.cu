// mlp_tc.cu -- fused shared per-point MLP on the 5th-gen tensor cores (impl 1): tcgen05.mma with
// TMEM accumulators, operands staged in shared memory (weights by bulk-copy TMA), 3xTF32 split
// arithmetic so the result matches true fp32 to ~1e-6 (the reference's convs are true fp32;
// single-pass TF32 misses the 1e-4 pose tolerance, SURVEY section 7 "hard parts").
#include "mlp_common.cuh"
#include "tc_common.cuh"

#include <math.h>

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
// One persistent CTA per SM walks 128-row tiles.  Warps 0-3 ("row threads", thread r <-> tile row r
// <-> TMEM lane r) produce the A operand and run the epilogues; warp 4 (one elected thread) issues
// tcgen05.mma; warp 5 (one elected thread) streams the pre-split weights with bulk-copy TMA.
//
//   layer l, K slab s (16 input channels):
//     row threads : A slab -> (hi, lo) planes of a_stage[s & 1]
//                   layer 0: gathered from global through the index list (SA) / row pointers (dense)
//                   layer>0: split from the fp32 activation plane kept in shared memory
//     TMA thread  : W slab (hi plane | lo plane, already split and chunk-major in global) -> w_stage[s & 1]
//     MMA thread  : 2 k-steps x 3 terms (lo*hi, hi*lo, hi*hi) of tcgen05.mma kind::tf32 into TMEM,
//                   tcgen05.commit -> empty[s & 1]; after the last slab also -> d_ready
//   epilogue l   : row threads tcgen05.ld their lane, + bias, ReLU, then
//                   l < last : write the fp32 activation plane (chunk-major, same addressing as the
//                              operand planes, so thread r only ever touches row r: no CTA barrier)
//                   l = last : SA   -> max over the K rows of each centroid by redux.sync on the
//                                      (non-negative) float bit patterns, one coalesced row write
//                              dense-> the thread's row straight to global (point-major)
//
// Shared-memory operand layout (K-major, no swizzle): element (row, k) of a 16-channel slab lives at
//   plane + (k/4) * ROWS*16 + row*16 + (k%4)*4      -> LBO = ROWS*16, SBO = 128 in the descriptors.
// ------------------------------------------------------------------------------------------------
constexpr int TC_THREADS = 192;
constexpr int TC_KC = 16;                         // channels per slab
constexpr int TC_A_PLANE = 4 * TC_CHUNK_BYTES;    // bytes of one (hi or lo) A slab plane: 8 KB
constexpr int TC_A_STAGE = 2 * TC_A_PLANE;        // hi + lo
constexpr int TC_TMEM_COLS = 256;

struct TcArgs {
    int nlayers, relu_last, cout_last;
    int kpad[CAPTRA_MAX_MLP_LAYERS], npad[CAPTRA_MAX_MLP_LAYERS];
    const float *wpk[CAPTRA_MAX_MLP_LAYERS];   // [nslab][2 planes][4 chunks][npad][4]
    const float *bias[CAPTRA_MAX_MLP_LAYERS];  // [npad]
    int64_t rows, ntiles;
    int group;                                  // SA: nsample (32|64|128); dense: 0
    float *out; int64_t ldo; int col_off;
    int n, s, cfeat; const float *xyz, *new_xyz, *feats; const int *idx;
    const float *segA; int64_t ldA; int ca; const float *segB; int64_t ldB; int cb; int bcast;
    int wstage_bytes, act_bytes, bias_floats;
    int nsplit, last_npad, cout_total;          // single wide layer split into 256-column chunks over grid.y
};

template <int MODE>  // 0: SA gather loader, 1: dense-row loader; the last epilogue is chosen by a.group
__global__ void __launch_bounds__(TC_THREADS, 1) mlp_tc_kernel(TcArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ uint64_t full_a[2], full_w[2], empty[2], d_ready;
    __shared__ uint32_t tmem_base_s;

    uint8_t *a_stage = smem_raw;                                   // 2 x TC_A_STAGE
    uint8_t *w_stage = a_stage + 2 * TC_A_STAGE;                   // 2 x wstage_bytes
    float *act = reinterpret_cast<float *>(w_stage + 2 * (size_t)a.wstage_bytes);
    float *bias_s = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(act) + a.act_bytes);
    float *red = bias_s + a.bias_floats;                           // [4][256]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // a single wide layer is split over grid.y: this CTA owns output columns [256*y, 256*y + npad)
    int npad_y = 0, col_off = a.col_off, cout_last = a.cout_last;
    const float *wpk0 = a.wpk[0], *bias0 = a.bias[0];
    if (a.nsplit > 1) {
        const int y = blockIdx.y;
        npad_y = (y < a.nsplit - 1) ? 256 : a.last_npad;
        wpk0 += (size_t)y * a.kpad[0] * 256 * 2;
        bias0 += y * 256;
        col_off += y * 256;
        cout_last = min(256, a.cout_total - y * 256);
    }
    auto NPAD = [&](int l) { return a.nsplit > 1 ? npad_y : a.npad[l]; };
    auto WPK = [&](int l) { return l == 0 ? wpk0 : a.wpk[l]; };
    auto BIAS = [&](int l) { return l == 0 ? bias0 : a.bias[l]; };

    if (warp == 4) tmem_alloc<TC_TMEM_COLS>(&tmem_base_s);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) { mbar_init(&full_a[i], 128); mbar_init(&full_w[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(&d_ready, 1);
        fence_mbar_init();
    }
    {   // biases of all layers -> smem
        int off = 0;
        for (int l = 0; l < a.nlayers; ++l) {
            for (int i = tid; i < NPAD(l); i += TC_THREADS) bias_s[off + i] = __ldg(BIAS(l) + i);
            off += NPAD(l);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_d = tmem_base_s;

    if (tid == 128) {
        // ===================== MMA issuer =====================
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
            for (int l = 0; l < a.nlayers; ++l) {
                const int nslab = a.kpad[l] / TC_KC;
                const uint32_t npad = (uint32_t)NPAD(l);
                const uint32_t idesc = make_idesc(2, TC_ROWS, (int)npad);
                const uint32_t b_lbo = npad * 16, b_lo_off = 4 * npad * 16;
                for (int s = 0; s < nslab; ++s, ++it) {
                    const uint32_t st = it & 1, ph = (it >> 1) & 1;
                    mbar_wait(&full_a[st], ph);
                    mbar_wait(&full_w[st], ph);
                    tcgen05_fence_after();
                    const uint32_t abase = smem_u32(a_stage + st * TC_A_STAGE);
                    const uint32_t bbase = smem_u32(w_stage + (size_t)st * a.wstage_bytes);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const uint32_t ao = (uint32_t)(2 * j) * TC_CHUNK_BYTES, bo = (uint32_t)(2 * j) * b_lbo;
                        const uint64_t ah = smem_desc_kmajor_noswz(abase + ao, TC_CHUNK_BYTES, 128);
                        const uint64_t al = smem_desc_kmajor_noswz(abase + TC_A_PLANE + ao, TC_CHUNK_BYTES, 128);
                        const uint64_t bh = smem_desc_kmajor_noswz(bbase + bo, b_lbo, 128);
                        const uint64_t bl = smem_desc_kmajor_noswz(bbase + b_lo_off + bo, b_lbo, 128);
                        umma_tf32(tmem_d, al, bh, idesc, (s | j) ? 1u : 0u);
                        umma_tf32(tmem_d, ah, bl, idesc, 1u);
                        umma_tf32(tmem_d, ah, bh, idesc, 1u);
                    }
                    umma_commit(&empty[st]);
                    if (s == nslab - 1) umma_commit(&d_ready);
                }
            }
        }
    } else if (tid == 160) {
        // ===================== weight producer (bulk-copy TMA) =====================
        uint32_t it = 0;
        for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
            for (int l = 0; l < a.nlayers; ++l) {
                const int nslab = a.kpad[l] / TC_KC;
                const uint32_t bytes = 2u * 4u * (uint32_t)NPAD(l) * 16u;
                for (int s = 0; s < nslab; ++s, ++it) {
                    const uint32_t st = it & 1, ph = (it >> 1) & 1;
                    mbar_wait(&empty[st], ph ^ 1);
                    mbar_arrive_expect_tx(&full_w[st], bytes);
                    bulk_g2s(w_stage + (size_t)st * a.wstage_bytes,
                             reinterpret_cast<const uint8_t *>(WPK(l)) + (size_t)s * bytes, bytes, &full_w[st]);
                }
            }
        }
    } else if (tid < 128) {
        // ===================== row threads: A producer + epilogue =====================
        const int r = tid;
        uint32_t it = 0, dl = 0;
        for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
            const int64_t grow = tile * TC_ROWS + r;
            const bool valid = grow < a.rows;
            // ---- per-row metadata
            const float *frow = nullptr;   // SA: feature row of the gathered point
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
            // layer-0 input element c of this row
            auto in0 = [&](int c) -> float {
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
            auto load_slab0 = [&](int s, float4 (&v)[4]) {
                const int c0 = s * TC_KC;
                const int nvec = MODE == 0 ? a.cfeat : a.ca;       // leading segment length
                const float *base = MODE == 0 ? frow : arow;
                const bool vec_ok = valid && base && (c0 + TC_KC <= nvec) &&
                                    ((reinterpret_cast<uintptr_t>(base + c0) & 15) == 0);
                if (vec_ok) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) v[q] = __ldg(reinterpret_cast<const float4 *>(base + c0) + q);
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q)
                        v[q] = make_float4(in0(c0 + 4 * q), in0(c0 + 4 * q + 1), in0(c0 + 4 * q + 2), in0(c0 + 4 * q + 3));
                }
            };
            auto store_slab = [&](uint32_t st, const float4 (&v)[4]) {
                float *hi = reinterpret_cast<float *>(a_stage + st * TC_A_STAGE);
                float *lo = reinterpret_cast<float *>(a_stage + st * TC_A_STAGE + TC_A_PLANE);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float4 h, l;
                    split_tf32(v[q].x, h.x, l.x); split_tf32(v[q].y, h.y, l.y);
                    split_tf32(v[q].z, h.z, l.z); split_tf32(v[q].w, h.w, l.w);
                    *reinterpret_cast<float4 *>(hi + q * (TC_ROWS * 4) + r * 4) = h;
                    *reinterpret_cast<float4 *>(lo + q * (TC_ROWS * 4) + r * 4) = l;
                }
            };

            int bias_off = 0;
            for (int l = 0; l < a.nlayers; ++l) {
                const int nslab = a.kpad[l] / TC_KC;
                // ---------------- produce the A slabs of this layer ----------------
                float4 cur[4], nxt[4];
                if (l == 0) load_slab0(0, cur);
                for (int s = 0; s < nslab; ++s, ++it) {
                    const uint32_t st = it & 1, ph = (it >> 1) & 1;
                    if (l == 0) {
                        if (s + 1 < nslab) load_slab0(s + 1, nxt);   // one slab of lookahead
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            cur[q] = *reinterpret_cast<const float4 *>(act + (size_t)(4 * s + q) * (TC_ROWS * 4) + r * 4);
                    }
                    mbar_wait(&empty[st], ph ^ 1);
                    store_slab(st, cur);
                    fence_proxy_async_smem();
                    tcgen05_fence_before();
                    mbar_arrive(&full_a[st]);
                    if (l == 0) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) cur[q] = nxt[q];
                    }
                }
                // ---------------- epilogue of this layer ----------------
                mbar_wait(&d_ready, dl & 1);
                ++dl;
                tcgen05_fence_after();
                const bool last = l == a.nlayers - 1;
                const bool relu = !last || a.relu_last;
                const int npad = NPAD(l);
                const uint32_t trow = tmem_d + ((uint32_t)(warp * 32) << 16);
                for (int c0 = 0; c0 < npad; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld_32x16(trow + (uint32_t)c0, v);
                    tmem_ld_wait();
                    float x[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const float t = __uint_as_float(v[j]) + bias_s[bias_off + c0 + j];
                        x[j] = relu ? fmaxf(t, 0.f) : t;
                    }
                    if (!last) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            *reinterpret_cast<float4 *>(act + (size_t)(c0 / 4 + q) * (TC_ROWS * 4) + r * 4) =
                                make_float4(x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
                    } else if (a.group > 0) {
                        // max over the 32 rows of this warp, column by column; lane j keeps column c0+j
                        uint32_t keep = 0;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const uint32_t m = __reduce_max_sync(kFull, valid ? __float_as_uint(x[j]) : 0u);
                            if (lane == j) keep = m;
                        }
                        if (lane < 16) red[warp * 256 + c0 + lane] = __uint_as_float(keep);
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
                if (last && a.group > 0) {
                    // combine the per-warp maxima of the warps that share a centroid and write it out
                    const int wpg = a.group / 32;                 // warps per centroid: 1, 2 or 4
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    const int ngroups = 4 / wpg;
                    for (int o = tid; o < ngroups * npad; o += 128) {
                        const int g = o / npad, c = o - g * npad;
                        const int64_t cen = (tile * TC_ROWS) / a.group + g;
                        if (c >= cout_last || cen * a.group >= a.rows) continue;
                        float m = red[(g * wpg) * 256 + c];
                        for (int w = 1; w < wpg; ++w) m = fmaxf(m, red[(g * wpg + w) * 256 + c]);
                        a.out[cen * a.ldo + col_off + c] = m;
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                }
                tcgen05_fence_before();
                bias_off += npad;
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 4) tmem_dealloc<TC_TMEM_COLS>(tmem_d);
}

// weights -> [slab][plane hi|lo][chunk][npad][4], pre-split, zero padded; bias -> [npad]
__global__ void pack_tc_kernel(int cin, int cout, int kpad, int npad, const float *__restrict__ w,
                               const float *__restrict__ bias, float *__restrict__ wpk, float *__restrict__ bp) {
    const int nslab = kpad / TC_KC;
    const int per_slab = 2 * 4 * npad * 4;
    const int total = nslab * per_slab / 2;  // one thread per (slab, chunk, n, e) -> writes hi and lo
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int e = i & 3;
        const int n = (i >> 2) % npad;
        const int q = ((i >> 2) / npad) & 3;
        const int s = (i >> 2) / npad / 4;
        const int k = s * TC_KC + q * 4 + e;
        const float v = (k < cin && n < cout) ? w[(size_t)n * cin + k] : 0.f;
        uint32_t h;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
        const float hi = __uint_as_float(h);
        const size_t o = (size_t)s * per_slab + ((size_t)q * npad + n) * 4 + e;
        wpk[o] = hi;
        wpk[o + (size_t)4 * npad * 4] = v - hi;
    }
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < npad; c += gridDim.x * blockDim.x)
        bp[c] = (c < cout && bias) ? bias[c] : 0.f;
}

struct TcLayout {
    int nlayers, kpad[CAPTRA_MAX_MLP_LAYERS], npad[CAPTRA_MAX_MLP_LAYERS];
    size_t off_w[CAPTRA_MAX_MLP_LAYERS], off_b[CAPTRA_MAX_MLP_LAYERS], total_floats;
    int wstage_bytes, act_bytes, bias_floats, nsplit, last_npad;
    size_t smem_bytes;
    bool supported;
};

static TcLayout tc_layout(const captra_mlp_desc &d) {
    TcLayout L{};
    L.nlayers = d.nlayers;
    L.supported = true;
    size_t off = 0;
    int cin = d.cin, npmax = 16, actmax = 0;
    L.nsplit = 1;
    for (int l = 0; l < d.nlayers; ++l) {
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
        if (l < d.nlayers - 1) actmax = max(actmax, L.npad[l]);
        L.off_w[l] = off; off += (size_t)L.kpad[l] * L.npad[l] * 2;
        L.off_b[l] = off; off += L.npad[l];
        L.bias_floats += L.npad[l];
        cin = d.cout[l];
    }
    L.total_floats = off;
    L.wstage_bytes = 2 * 4 * npmax * 16;
    L.act_bytes = (actmax / 4) * TC_CHUNK_BYTES;
    L.smem_bytes = (size_t)2 * TC_A_STAGE + 2 * (size_t)L.wstage_bytes + L.act_bytes + (size_t)L.bias_floats * 4 + 4 * 256 * 4;
    if (L.smem_bytes > 225 * 1024) L.supported = false;
    return L;
}

int64_t tc_pack_bytes(const captra_mlp_desc *d) {
    const TcLayout L = tc_layout(*d);
    return L.supported ? (int64_t)(L.total_floats * sizeof(float)) : -1;
}

bool tc_supported(const captra_mlp_desc *d) { return tc_layout(*d).supported; }

int tc_pack(const captra_mlp_desc *d, void *packed, cudaStream_t stream) {
    const TcLayout L = tc_layout(*d);
    CAPTRA_REQUIRE(L.supported, "mlp_pack(tc): layer widths not supported by the tcgen05 path");
    int cin = d->cin;
    for (int l = 0; l < d->nlayers; ++l) {
        float *base = reinterpret_cast<float *>(packed);
        for (int y = 0; y < L.nsplit; ++y) {   // nsplit > 1 only for a single wide layer
            const int npad_y = L.nsplit == 1 ? L.npad[l] : (y < L.nsplit - 1 ? 256 : L.last_npad);
            const int cout_y = L.nsplit == 1 ? d->cout[l] : min(256, d->cout[l] - 256 * y);
            pack_tc_kernel<<<64, 256, 0, stream>>>(cin, cout_y, L.kpad[l], npad_y, d->w[l] + (size_t)y * 256 * cin,
                                                   d->bias[l] ? d->bias[l] + y * 256 : nullptr,
                                                   base + L.off_w[l] + (size_t)y * L.kpad[l] * 256 * 2,
                                                   base + L.off_b[l] + y * 256);
            CAPTRA_CHECK_LAUNCH("mlp_pack(tc)");
        }
        cin = d->cout[l];
    }
    return CAPTRA_OK;
}

static int tc_fill(TcArgs &a, const captra_mlp_desc *d, const void *packed, size_t *smem) {
    const TcLayout L = tc_layout(*d);
    CAPTRA_REQUIRE(L.supported, "mlp(tc): layer widths not supported by the tcgen05 path");
    a.nlayers = d->nlayers; a.relu_last = d->relu_last; a.cout_last = d->cout[d->nlayers - 1];
    for (int l = 0; l < d->nlayers; ++l) {
        a.kpad[l] = L.kpad[l]; a.npad[l] = L.npad[l];
        a.wpk[l] = reinterpret_cast<const float *>(packed) + L.off_w[l];
        a.bias[l] = reinterpret_cast<const float *>(packed) + L.off_b[l];
    }
    a.wstage_bytes = L.wstage_bytes; a.act_bytes = L.act_bytes; a.bias_floats = L.bias_floats;
    a.nsplit = L.nsplit; a.last_npad = L.last_npad; a.cout_total = d->cout[d->nlayers - 1];
    *smem = L.smem_bytes;
    return CAPTRA_OK;
}

template <int MODE>
static int tc_launch(TcArgs &a, size_t smem, cudaStream_t stream) {
    auto kern = mlp_tc_kernel<MODE>;
    CAPTRA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem + 1024));
    a.ntiles = ceil_div<int64_t>(a.rows, TC_ROWS);
    const int nsm = max(1, sm_count() / a.nsplit);
    const int gx = a.ntiles < nsm ? (int)a.ntiles : nsm;
    kern<<<dim3(gx, a.nsplit), TC_THREADS, smem, stream>>>(a);
    CAPTRA_CHECK_LAUNCH("mlp_tc");
    return CAPTRA_OK;
}

int tc_sa_mlp_max(int b, int n, int s, int k, int cfeat, const float *xyz, const float *new_xyz,
                  const float *feats, const int *idx, const captra_mlp_desc *d, const void *packed,
                  float *out, int64_t ldo, int col_off, cudaStream_t stream) {
    CAPTRA_REQUIRE(k == 32 || k == 64 || k == 128, "sa_mlp_max(tc): nsample must be 32, 64 or 128 (got %d)", k);
    CAPTRA_REQUIRE(d->relu_last, "sa_mlp_max(tc): the max epilogue needs a ReLU after the last layer");
    TcArgs a{};
    size_t smem;
    int rc = tc_fill(a, d, packed, &smem);
    if (rc) return rc;
    a.rows = (int64_t)b * s * k; a.group = k;
    a.out = out; a.ldo = ldo; a.col_off = col_off;
    a.n = n; a.s = s; a.cfeat = cfeat; a.xyz = xyz; a.new_xyz = new_xyz; a.feats = feats; a.idx = idx;
    return tc_launch<0>(a, smem, stream);
}

int tc_point_mlp(int64_t rows, const float *segA, int64_t ldA, int ca, const float *segB, int64_t ldB, int cb,
                 int bcast, const captra_mlp_desc *d, const void *packed, float *y, int64_t ldy, int col_off,
                 int group, cudaStream_t stream) {
    CAPTRA_REQUIRE(group == 0 || ((group == 32 || group == 64 || group == 128) && d->relu_last),
                   "point_mlp(tc): grouped max needs group in {32,64,128} and a final ReLU");
    TcArgs a{};
    size_t smem;
    int rc = tc_fill(a, d, packed, &smem);
    if (rc) return rc;
    a.rows = rows; a.group = group;
    a.out = y; a.ldo = ldy; a.col_off = col_off;
    a.segA = segA; a.ldA = ldA; a.ca = ca; a.segB = segB; a.ldB = ldB; a.cb = cb; a.bcast = bcast;
    return tc_launch<1>(a, smem, stream);
}

}  // namespace captra

using namespace captra;

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
