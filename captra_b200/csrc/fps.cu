// fps.cu -- iterative furthest point sampling (reference: sampling_gpu.cu:86-209).
//
// The reference re-reads the cloud and its running min-distance array from global memory in
// every one of the M-1 rounds and reduces through a log2(block)-deep __syncthreads tree.
// B200 design: FPS is a latency chain, so everything a round touches lives in registers --
// each thread owns PPT points (x,y,z,temp = 4*PPT registers) -- and a round is
//   PPT distance updates -> warp arg-max by two `redux.sync` (value, then tie-break key)
//   -> one 8-byte record per warp in shared memory -> ONE __syncthreads
//   -> every warp re-reduces the <=32 records redundantly (no second barrier; records are
//      double-buffered by round parity) -> winner coordinates by a broadcast LDS from the
//      shared-memory copy of the cloud.
//
// Exact tie rule: the reference thread t = k mod block scans k ascending
// with strict '>', and its tournament keeps the left operand on ties, so among equal maxima
// the winner minimises (bitrev_{log2 block}(k mod block), k), block = min(1024, 2^floor(log2 n))
// (cuda_utils.h:10-14).  We reduce on the 64-bit key (ordered(dist), ~((bitrev << 22) | k)),
// which reproduces that total order for any assignment of points to threads.
#include "common.cuh"

#include <math.h>
#include <stdlib.h>

namespace captra {

__device__ __forceinline__ unsigned ordered_bits(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__device__ __forceinline__ unsigned tie_key(int k, int ref_mask, int ref_shift) {
    const unsigned rev = __brev((unsigned)(k & ref_mask)) >> ref_shift;  // bitrev over log2(block) bits
    return ~((rev << 22) | (unsigned)k);
}

// progress != null (captra_fps_ball_query): the kernel lets its stream successor start early (programmatic dependent
// launch) and publishes, every FPS_PUBLISH rounds (one block of the piped ball query), how many centroids of the cloud are final (release store after the
// writer's own new_xyz stores), so that a ball query can consume them while the sampling goes on.
constexpr int FPS_PUBLISH = 16;
template <int NT, int PPT>
__global__ void __launch_bounds__(NT)
fps_reg_kernel(int n, int m, int ref_bits, const float *__restrict__ dataset,
               float *__restrict__ temp, int *__restrict__ idxs, float *__restrict__ new_xyz, int *progress) {
    extern __shared__ float smem[];  // sx[n] sy[n] sz[n]
    if (progress) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    __shared__ uint2 rec[2][32];
    constexpr int NW = NT / 32;

    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *cloud = dataset + (size_t)b * n * 3;
    float *tmp = temp ? temp + (size_t)b * n : nullptr;   // null: running distances start at 1e10 and are not returned
    int *out = idxs + (size_t)b * m;
    float *oxyz = new_xyz ? new_xyz + (size_t)b * m * 3 : nullptr;
    float *sx = smem, *sy = smem + n, *sz = smem + 2 * n;
    const int ref_mask = (1 << ref_bits) - 1;
    const int ref_shift = ref_bits ? 32 - ref_bits : 31;  // n == 1: mask 0, key bits 0

    for (int e = tid; e < n * 3; e += NT) {
        const int pt = e / 3, comp = e - pt * 3;
        smem[comp * n + pt] = __ldg(cloud + e);
    }
    __syncthreads();

    float x[PPT], y[PPT], z[PPT], t[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        const int k = tid + j * NT;
        const bool ok = k < n;
        x[j] = ok ? sx[k] : 0.f; y[j] = ok ? sy[k] : 0.f; z[j] = ok ? sz[k] : 0.f;
        t[j] = ok ? (tmp ? tmp[k] : 1e10f) : -INFINITY;  // -inf never wins, never ties a real point
    }

    int old = 0;
    float ox = sx[0], oy = sy[0], oz = sz[0];
    if (tid == 0) {
        out[0] = 0;
        if (oxyz) { oxyz[0] = ox; oxyz[1] = oy; oxyz[2] = oz; }
    }

    for (int r = 1; r < m; ++r) {
        float tm = -INFINITY;
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            const float d = sqdist_ref(x[j], y[j], z[j], ox, oy, oz);
            t[j] = fminf(d, t[j]);
            tm = fmaxf(tm, t[j]);
        }
        const unsigned um = ordered_bits(tm);
        const unsigned wm = __reduce_max_sync(kFull, um);
        unsigned tb = 0;
        if (um == wm) {
#pragma unroll
            for (int j = 0; j < PPT; ++j)
                if (t[j] == tm) tb = max(tb, tie_key(tid + j * NT, ref_mask, ref_shift));
        }
        const unsigned wt = __reduce_max_sync(kFull, tb);
        unsigned bm, bt;
        if (NW > 1) {
            if (lane == 0) rec[r & 1][warp] = make_uint2(wm, wt);
            __syncthreads();
            const uint2 q = lane < NW ? rec[r & 1][lane] : make_uint2(0u, 0u);
            bm = __reduce_max_sync(kFull, q.x);
            bt = __reduce_max_sync(kFull, q.x == bm ? q.y : 0u);
        } else {
            bm = wm; bt = wt;
        }
        old = (int)((~bt) & 0x3fffffu);
        ox = sx[old]; oy = sy[old]; oz = sz[old];
        if (tid == 0) {
            out[r] = old;
            if (oxyz) { oxyz[r * 3 + 0] = ox; oxyz[r * 3 + 1] = oy; oxyz[r * 3 + 2] = oz; }
            if (progress && ((r + 1) % FPS_PUBLISH == 0 || r == m - 1))
                asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(progress + b), "r"(r + 1) : "memory");
        }
    }
    if (progress && m == 1 && tid == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(progress + b), "r"(1) : "memory");

    // the reference leaves its running min-distances in the caller's scratch
    if (tmp) {
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            const int k = tid + j * NT;
            if (k < n) tmp[k] = t[j];
        }
    }
}

// General-n path: running distances stay in the caller's temp (global/L2), coordinates are
// re-read through the read-only path each round.  Same key, same order.
template <int NT>
__global__ void __launch_bounds__(NT)
fps_stream_kernel(int n, int m, int ref_bits, const float *__restrict__ dataset,
                  float *__restrict__ temp, int *__restrict__ idxs, float *__restrict__ new_xyz) {
    __shared__ uint2 rec[2][32];
    constexpr int NW = NT / 32;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *cloud = dataset + (size_t)b * n * 3;
    float *tmp = temp + (size_t)b * n;
    int *out = idxs + (size_t)b * m;
    float *oxyz = new_xyz ? new_xyz + (size_t)b * m * 3 : nullptr;
    const int ref_mask = (1 << ref_bits) - 1;
    const int ref_shift = ref_bits ? 32 - ref_bits : 31;

    int old = 0;
    float ox = __ldg(cloud + 0), oy = __ldg(cloud + 1), oz = __ldg(cloud + 2);
    if (tid == 0) {
        out[0] = 0;
        if (oxyz) { oxyz[0] = ox; oxyz[1] = oy; oxyz[2] = oz; }
    }
    for (int r = 1; r < m; ++r) {
        unsigned um = 0, tb = 0;
        for (int k = tid; k < n; k += NT) {
            const float d = sqdist_ref(__ldg(cloud + k * 3 + 0), __ldg(cloud + k * 3 + 1),
                                       __ldg(cloud + k * 3 + 2), ox, oy, oz);
            const float d2 = fminf(d, tmp[k]);
            tmp[k] = d2;
            const unsigned u = ordered_bits(d2);
            const unsigned key = tie_key(k, ref_mask, ref_shift);
            if (u > um || (u == um && key > tb)) { um = u; tb = key; }
        }
        const unsigned wm = __reduce_max_sync(kFull, um);
        const unsigned wt = __reduce_max_sync(kFull, um == wm ? tb : 0u);
        if (lane == 0) rec[r & 1][warp] = make_uint2(wm, wt);
        __syncthreads();
        const uint2 q = lane < NW ? rec[r & 1][lane] : make_uint2(0u, 0u);
        const unsigned bm = __reduce_max_sync(kFull, q.x);
        const unsigned bt = __reduce_max_sync(kFull, q.x == bm ? q.y : 0u);
        old = (int)((~bt) & 0x3fffffu);
        ox = __ldg(cloud + old * 3 + 0); oy = __ldg(cloud + old * 3 + 1); oz = __ldg(cloud + old * 3 + 2);
        if (tid == 0) {
            out[r] = old;
            if (oxyz) { oxyz[r * 3 + 0] = ox; oxyz[r * 3 + 1] = oy; oxyz[r * 3 + 2] = oz; }
        }
    }
}


// ------------------------------------------------------------------------------------------------
// Cluster variant for clouds above one CTA's register capacity (8192 points) -- BASELINE cfg5's 16384 -> 4096 and
// the data loader's 20480 -> 4096 resample (datasets/data_utils.py:138-158).  The reference (and fps_stream_kernel)
// re-read the cloud and its distances from L2 in every one of the M-1 rounds.  Here a thread-block CLUSTER of CS
// CTAs owns a cloud: every point and its running distance live in registers (PPT per thread, CS*NT*PPT >= n), and a
// round is
//   PPT distance updates -> warp arg-max (two redux.sync) -> CTA arg-max -> the CTA's record {dist, key, x, y, z}
//   goes into EVERY CTA of the cluster (distributed shared memory, st.shared::cluster + a remote mbarrier arrive)
//   -> every warp re-reduces the CS records redundantly (see the kernel for the synchronisation).
// The winner's coordinates travel inside the record, so the loop touches no global memory at all.  Same 64-bit
// key as fps_reg_kernel => the reference's tie rule bit for bit.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_u32(uint32_t addr, uint32_t v) {
    asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init_cl(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_remote_arrive_release(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_acquire_cluster(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
    return ok != 0;
}

// Round structure: (1) warp arg-max -> per-warp record in LOCAL shared memory -> __syncthreads -> warp 0 reduces the
// CTA's NW records; (2) one lane writes the CTA record {dist, key, x, y, z} into every CTA of the cluster
// (st.shared::cluster) and arrives, with release at cluster scope, on that CTA's mbarrier -- CS remote arrivals per
// barrier and round, from ONE thread per CTA (the hardware cluster barrier, which every warp of the cluster has to
// reach, measured 2-3 us per round here); (3) warp 0 waits on the local mbarrier (acquire.cluster), __syncthreads hands
// the records to the other warps, and every warp reduces the CS records redundantly.  Records and barriers are
// double-buffered by round parity: a CTA can write round r+2's record only after round r+1 completed everywhere,
// i.e. after every peer has consumed round r's.
template <int NT, int PPT, int CS>
__global__ void __launch_bounds__(NT)
fps_cluster_kernel(int n, int m, int ref_bits, const float *__restrict__ dataset,
                   float *__restrict__ temp, int *__restrict__ idxs, float *__restrict__ new_xyz) {
    constexpr int NW = NT / 32;
    static_assert(CS <= 32 && NW <= 32, "one record per lane in the reductions");
    __shared__ uint32_t wrec[2][5][NW];                   // per-warp records of this CTA [round parity]
    __shared__ uint32_t crec[2][5][CS];                   // per-CTA records of the cluster (written remotely)
    __shared__ uint64_t xbar[2];                          // CS arrivals each
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t crank = cluster_ctarank();
    const int b = blockIdx.x / CS;
    const float *cloud = dataset + (size_t)b * n * 3;
    float *tmp = temp ? temp + (size_t)b * n : nullptr;
    int *out = idxs + (size_t)b * m;
    float *oxyz = new_xyz ? new_xyz + (size_t)b * m * 3 : nullptr;
    const int ref_mask = (1 << ref_bits) - 1;
    const int ref_shift = ref_bits ? 32 - ref_bits : 31;
    const bool writer = crank == 0 && tid == 0;

    if (tid == 0) {
        mbar_init_cl(&xbar[0], CS);
        mbar_init_cl(&xbar[1], CS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    float x[PPT], y[PPT], z[PPT], t[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        const int k = (j * CS + (int)crank) * NT + tid;     // consecutive threads <-> consecutive points
        const bool ok = k < n;
        x[j] = ok ? __ldg(cloud + (size_t)k * 3 + 0) : 0.f;
        y[j] = ok ? __ldg(cloud + (size_t)k * 3 + 1) : 0.f;
        z[j] = ok ? __ldg(cloud + (size_t)k * 3 + 2) : 0.f;
        t[j] = ok ? (tmp ? tmp[k] : 1e10f) : -INFINITY;
    }
    float ox = __ldg(cloud + 0), oy = __ldg(cloud + 1), oz = __ldg(cloud + 2);
    if (writer) {
        out[0] = 0;
        if (oxyz) { oxyz[0] = ox; oxyz[1] = oy; oxyz[2] = oz; }
    }
    const uint32_t crec_addr = (uint32_t)__cvta_generic_to_shared(&crec[0][0][crank]);
    const uint32_t xbar_addr = (uint32_t)__cvta_generic_to_shared(&xbar[0]);
    cluster_sync_all();                                   // barriers initialised and every CTA resident before remote traffic

    for (int r = 1; r < m; ++r) {
        const int par = r & 1;
        float tm = -INFINITY;
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            const float d = sqdist_ref(x[j], y[j], z[j], ox, oy, oz);
            t[j] = fminf(d, t[j]);
            tm = fmaxf(tm, t[j]);
        }
        const unsigned um = ordered_bits(tm);
        const unsigned wm = __reduce_max_sync(kFull, um);
        unsigned tb = 0;
        float bx = 0.f, by = 0.f, bz = 0.f;
        if (um == wm) {
#pragma unroll
            for (int j = 0; j < PPT; ++j) {
                const unsigned key = tie_key((j * CS + (int)crank) * NT + tid, ref_mask, ref_shift);
                if (t[j] == tm && key > tb) { tb = key; bx = x[j]; by = y[j]; bz = z[j]; }
            }
        }
        const unsigned wt = __reduce_max_sync(kFull, tb);
        if (um == wm && tb == wt) {                        // exactly one lane: keys are unique
            wrec[par][0][warp] = wm; wrec[par][1][warp] = wt;
            wrec[par][2][warp] = __float_as_uint(bx); wrec[par][3][warp] = __float_as_uint(by); wrec[par][4][warp] = __float_as_uint(bz);
        }
        __syncthreads();
        if (warp == 0) {
            // CTA winner among the NW warp records
            const unsigned qm = lane < NW ? wrec[par][0][lane] : 0u, qt = lane < NW ? wrec[par][1][lane] : 0u;
            const unsigned cm = __reduce_max_sync(kFull, qm);
            const unsigned ct = __reduce_max_sync(kFull, qm == cm ? qt : 0u);
            const int src = __ffs(__ballot_sync(kFull, qm == cm && qt == ct)) - 1;
            // lane c < CS delivers the record to CTA c and arrives on its barrier (release at cluster scope)
            if (lane < CS) {
                const uint32_t ra = mapa_shared(crec_addr + (uint32_t)(par * 5 * CS) * 4u, (uint32_t)lane);
                st_cluster_u32(ra, cm);
                st_cluster_u32(ra + (uint32_t)CS * 4u, ct);
                st_cluster_u32(ra + (uint32_t)(2 * CS) * 4u, wrec[par][2][src]);
                st_cluster_u32(ra + (uint32_t)(3 * CS) * 4u, wrec[par][3][src]);
                st_cluster_u32(ra + (uint32_t)(4 * CS) * 4u, wrec[par][4][src]);
                mbar_remote_arrive_release(mapa_shared(xbar_addr + (uint32_t)par * 8u, (uint32_t)lane));
            }
            if (lane == 0)
                // xbar[1] serves rounds 1, 3, 5, ..., xbar[0] rounds 2, 4, 6, ...: the use count of either is (r - 1) / 2
                while (!mbar_try_wait_acquire_cluster(&xbar[par], (uint32_t)(((r - 1) >> 1) & 1))) {}
            __syncwarp();
        }
        __syncthreads();
        // every warp reduces the CS cluster records redundantly
        const unsigned em = lane < CS ? crec[par][0][lane] : 0u, et = lane < CS ? crec[par][1][lane] : 0u;
        const unsigned bm = __reduce_max_sync(kFull, em);
        const unsigned bt = __reduce_max_sync(kFull, em == bm ? et : 0u);
        const int wi = __ffs(__ballot_sync(kFull, em == bm && et == bt)) - 1;
        ox = __uint_as_float(crec[par][2][wi]);
        oy = __uint_as_float(crec[par][3][wi]);
        oz = __uint_as_float(crec[par][4][wi]);
        if (writer) {
            out[r] = (int)((~bt) & 0x3fffffu);
            if (oxyz) { oxyz[r * 3 + 0] = ox; oxyz[r * 3 + 1] = oy; oxyz[r * 3 + 2] = oz; }
        }
    }
    if (tmp) {
#pragma unroll
        for (int j = 0; j < PPT; ++j) {
            const int k = (j * CS + (int)crank) * NT + tid;
            if (k < n) tmp[k] = t[j];
        }
    }
    cluster_sync_all();                                   // no CTA exits while a peer may still address its shared memory
}

template <int NT, int PPT, int CS>
static int launch_fps_cluster(int b, int n, int m, int ref_bits, const float *dataset, float *temp,
                              int *idxs, float *new_xyz, cudaStream_t stream) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)b * CS);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CAPTRA_CUDA(cudaLaunchKernelEx(&cfg, fps_cluster_kernel<NT, PPT, CS>, n, m, ref_bits, dataset, temp, idxs, new_xyz));
    CAPTRA_CHECK_LAUNCH("furthest_point_sampling(cluster)");
    return CAPTRA_OK;
}

template <int NT, int PPT>
static int launch_fps_reg(int b, int n, int m, int ref_bits, const float *dataset, float *temp,
                          int *idxs, float *new_xyz, cudaStream_t stream, int *progress = nullptr) {
    auto kern = fps_reg_kernel<NT, PPT>;
    const size_t smem = sizeof(float) * 3 * (size_t)n;
    // static smem (the records) counts against the 48 KB default; opting in needs a value > 48 KB
    if (smem + 1024 > 48 * 1024)
        CAPTRA_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem + 1024));
    kern<<<b, NT, smem, stream>>>(n, m, ref_bits, dataset, temp, idxs, new_xyz, progress);
    CAPTRA_CHECK_LAUNCH("furthest_point_sampling");
    return CAPTRA_OK;
}

static int fps_ref_bits(int n) {
    int ref_bits = 0;
    while ((2 << ref_bits) <= n && ref_bits < 10) ++ref_bits;
    return ref_bits;
}

// FPS + gather with progress publication (register-resident shapes only: n <= 8192); progress [b] must be zero
int fps_gather_progress(int b, int n, int m, const float *dataset, int *idxs, float *new_xyz, int *progress, cudaStream_t s) {
    const int ref_bits = fps_ref_bits(n);
    if (n <= 128) return launch_fps_reg<32, 4>(b, n, m, ref_bits, dataset, nullptr, idxs, new_xyz, s, progress);
    if (n <= 512) return launch_fps_reg<128, 4>(b, n, m, ref_bits, dataset, nullptr, idxs, new_xyz, s, progress);
    if (n <= 1024) return launch_fps_reg<256, 4>(b, n, m, ref_bits, dataset, nullptr, idxs, new_xyz, s, progress);
    if (n <= 2048) return launch_fps_reg<256, 8>(b, n, m, ref_bits, dataset, nullptr, idxs, new_xyz, s, progress);
    if (n <= 4096) return launch_fps_reg<512, 8>(b, n, m, ref_bits, dataset, nullptr, idxs, new_xyz, s, progress);
    return launch_fps_reg<1024, 8>(b, n, m, ref_bits, dataset, nullptr, idxs, new_xyz, s, progress);
}

}  // namespace captra

using namespace captra;

extern "C" int captra_fps_gather(int b, int n, int m, const float *dataset, float *temp,
                                 int *idxs, float *new_xyz, captra_stream_t stream) {
    CAPTRA_REQUIRE(b >= 0 && n >= 0 && m >= 0, "fps: negative size");
    if (b == 0 || m <= 0) return CAPTRA_OK;  // sampling_gpu.cu:101 `if (m <= 0) return`
    CAPTRA_REQUIRE(n >= 1, "fps: empty cloud with m > 0");
    CAPTRA_REQUIRE(n < (1 << 22), "fps: n=%d exceeds the 2^22 key width", n);
    CAPTRA_REQUIRE(dataset && idxs, "fps: null pointer");
    CAPTRA_REQUIRE(temp || n <= 32768, "fps: clouds above 32768 points keep their running distances in `temp` (must not be NULL)");
    // block = min(1024, 2^floor(log2 n)) (cuda_utils.h:10-14); exact integer log2 here -- the
    // reference's log()/log() quotient evaluates to the same integer for every n < 2^22
    // (checked exhaustively in tests/test_oracle.py).
    int ref_bits = 0;
    while ((2 << ref_bits) <= n && ref_bits < 10) ++ref_bits;
    cudaStream_t s = as_stream(stream);
    if (n <= 128) return launch_fps_reg<32, 4>(b, n, m, ref_bits, dataset, temp, idxs, new_xyz, s);
    if (n <= 512) return launch_fps_reg<128, 4>(b, n, m, ref_bits, dataset, temp, idxs, new_xyz, s);
    if (n <= 1024) return launch_fps_reg<256, 4>(b, n, m, ref_bits, dataset, temp, idxs, new_xyz, s);
    if (n <= 2048) return launch_fps_reg<256, 8>(b, n, m, ref_bits, dataset, temp, idxs, new_xyz, s);
    if (n <= 4096) return launch_fps_reg<512, 8>(b, n, m, ref_bits, dataset, temp, idxs, new_xyz, s);
    if (n <= 8192) return launch_fps_reg<1024, 8>(b, n, m, ref_bits, dataset, temp, idxs, new_xyz, s);
    // above one CTA's registers: a cluster of 4 / 8 CTAs per cloud (CAPTRA_FPS_CLUSTER=0 forces the streaming kernel)
    static const bool use_cluster = [] { const char *e = getenv("CAPTRA_FPS_CLUSTER"); return !e || atoi(e) != 0; }();
    if (use_cluster && (int64_t)b * 8 < (1LL << 31)) {
        // 4 points per thread: 16 data registers, so the 64-register budget of a 1024-thread CTA holds without spills
        if (n <= 16384) return launch_fps_cluster<1024, 4, 4>(b, n, m, ref_bits, dataset, temp, idxs, new_xyz, s);
        if (n <= 32768) return launch_fps_cluster<1024, 4, 8>(b, n, m, ref_bits, dataset, temp, idxs, new_xyz, s);
    }
    CAPTRA_REQUIRE(temp, "fps: the streaming kernel keeps its running distances in `temp` (must not be NULL)");
    fps_stream_kernel<1024><<<b, 1024, 0, s>>>(n, m, ref_bits, dataset, temp, idxs, new_xyz);
    CAPTRA_CHECK_LAUNCH("furthest_point_sampling(stream)");
    return CAPTRA_OK;
}

extern "C" int furthest_point_sampling_kernel_launcher(int b, int n, int m, const float *dataset,
                                                       float *temp, int *idxs,
                                                       captra_stream_t stream) {
    return captra_fps_gather(b, n, m, dataset, temp, idxs, nullptr, stream);
}
