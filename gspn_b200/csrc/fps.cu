// fps.cu -- farthest point sampling (farthestpointsamplingKernel, tf_sampling_g.cu:105-170)
// re-designed for B200.
//
// The reference keeps the running min-distance in a global scratch array, re-reads every point
// past the first 3072 from global memory each round and reduces with a 9-step __syncthreads tree
// on one CTA per cloud.  FPS is a chain of m-1 strictly sequential rounds, so what matters is the
// latency of ONE round, not bandwidth.  Here a cloud is owned by a thread-block CLUSTER:
//   * every thread keeps its PPT points (x,y,z, packed as fp32x2 pairs) AND their running min-distance in
//     registers for the whole kernel -- global memory is touched once (prologue) and for the m index stores;
//   * per round: PPT/2 packed distance updates (FADD2/FMUL2/FFMA2), a tournament argmax, two redux.sync per
//     warp (max of the distance bits, then min of the tie-break key), then a DSMEM all-to-all: every warp's
//     winner (coordinates travel with the candidate, so the next round starts without a memory load) is pushed
//     into every CTA's candidate table with st.async + mbarrier complete_tx; each CTA waits on its own
//     mbarrier (double-buffered by round parity) and every warp reduces the table with one more warp argmax.
//     No __syncthreads, no cluster barrier, no gpu-scope fence on the round path.
//
// Bit-exactness with the reference: distances use its compiled rounding (common.cuh sqdist_fma);
// its winner among equal maxima is the lowest (k mod 512, k) -- thread-strided scan with strict '>'
// (:130,:146) + left-wins tree (:158).  The key  ((k&511)<<23)|(k>>9)  orders exactly like that,
// independent of how points are mapped to threads here.
#include <cstdio>
#include <cstdlib>
#include "fps_common.cuh"

namespace cg = cooperative_groups;

namespace gspn {

// argmax over a thread's PPT updated distances as a tournament (depth log2 PPT instead of a PPT-long
// dependent chain); the left operand survives ties, so the lowest j -- the lowest key -- wins.
template <int LO, int CNT>
struct Tourney {
    template <int PPT>
    static __device__ __forceinline__ void run(const float (&t)[PPT], float &v, int &j) {
        if constexpr (CNT == 1) {
            v = t[LO];
            j = LO;
        } else {
            float va, vb;
            int ja, jb;
            Tourney<LO, CNT / 2>::run(t, va, ja);
            Tourney<LO + CNT / 2, CNT - CNT / 2>::run(t, vb, jb);
            bool take = vb > va;
            v = take ? vb : va;
            j = take ? jb : ja;
        }
    }
};

// PPT points per thread in registers.  grid = (CLUSTER, b), cluster = (CLUSTER,1,1).
// Requires (blockDim.x*CLUSTER) % 512 == 0 or PPT == 1 so that a thread's points have ascending keys,
// and CLUSTER * (blockDim.x/32) <= kMaxCand when CLUSTER > 1.
// dynamic smem: float4 xyz copy, [PPT][blockDim.x], so only the round's winner lane fetches coordinates.
// CPC (clouds per CTA, CLUSTER > 1 only): 2 = the CTA is two independent halves, each doing what a CTA of blockDim.x / 2 threads does
// for its own cloud (nothing on the round path is CTA-wide: warp collectives and per-half mbarriers only).  Same latency, and the
// FPS state of a batch sits on half as many SMs, each of them full, instead of a third of the registers of twice as many -- which
// leaves whole SMs to the big MLP-chain CTAs of the other lanes in the pipelined step.
template <int PPT, int CLUSTER, int MAXT, bool PROFILE = false, int CPL = (CLUSTER > 1 ? 2 : 1), int CPC = 1>
__global__ void __launch_bounds__(MAXT, 1) fps_resident_kernel(int n, int m, const float *__restrict__ xyz, int *__restrict__ out,
                                                               long long *__restrict__ prof = nullptr, int nclouds = 0) {
    static_assert(CPC == 1 || CLUSTER > 1, "halves only synchronise through their own mbarriers");
    extern __shared__ float4 sxyz[];
    long long t0 = 0, t1 = 0, t2 = 0, t3 = 0, acc[4] = {0, 0, 0, 0};
    // candidate table: 32*CPL entries per parity; entries no warp owns stay "empty" (-1, max key) forever,
    // so the per-round reduce is CPL unconditional loads per lane
    __shared__ Slot wslot_[CPC][2][32 * CPL];
    __shared__ unsigned wkey_[CPC][2][32 * CPL];
    __shared__ __align__(8) uint64_t xbar_[CPC][2];

    const int tpc = blockDim.x / CPC;                             // threads of this cloud's part of the CTA
    const int part = CPC > 1 ? (int)threadIdx.x / tpc : 0;
    const int ltid = (int)threadIdx.x - part * tpc;
    Slot (*wslot)[32 * CPL] = wslot_[part];
    unsigned (*wkey)[32 * CPL] = wkey_[part];
    uint64_t *xbar = xbar_[part];
    int cloud = blockIdx.y * CPC + part;
    const bool live = CPC == 1 || cloud < nclouds;                // odd batch: the last half redoes the last cloud and writes nothing
    if (!live) cloud = nclouds - 1;
    unsigned rank = 0;
    if (CLUSTER > 1) rank = cg::this_cluster().block_rank();
    const int T = tpc * CLUSTER;
    const int gtid = rank * tpc + ltid;
    const int lane = ltid & 31, warp = ltid >> 5, nwarps = tpc >> 5;
    const float *p = xyz + (size_t)cloud * n * 3;
    float4 *sx = sxyz + (size_t)part * PPT * tpc;

    constexpr bool PACKED = (PPT % 2 == 0);
    constexpr int NP = PACKED ? PPT / 2 : 1;
    unsigned long long qx[NP], qy[NP], qz[NP];  // PACKED: the coordinates live as fp32x2 pairs (points 2i, 2i+1)
    float px[PPT], py[PPT], pz[PPT], td[PPT];
#pragma unroll
    for (int j = 0; j < PPT; ++j) {
        int k = gtid + j * T;
        if (k < n) {
            px[j] = __ldg(p + 3 * k); py[j] = __ldg(p + 3 * k + 1); pz[j] = __ldg(p + 3 * k + 2);
            td[j] = 1e38f;  // tf_sampling_g.cu:118
        } else {
            px[j] = py[j] = pz[j] = 0.f;
            td[j] = -1.0f;  // min(d,-1) = -1 never beats a real point (distances are >= 0)
        }
        sx[j * tpc + ltid] = make_float4(px[j], py[j], pz[j], 0.f);  // read back by this thread only
    }
    if (PACKED) {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
            qx[i] = f2_pack(px[2 * i], px[2 * i + 1]);
            qy[i] = f2_pack(py[2 * i], py[2 * i + 1]);
            qz[i] = f2_pack(pz[2 * i], pz[2 * i + 1]);
        }
    }
    float x1 = __ldg(p), y1 = __ldg(p + 1), z1 = __ldg(p + 2);  // old = 0 (:114)
    if (gtid == 0 && live) out[(size_t)cloud * m] = 0;
    const int ncand = CLUSTER * nwarps;
    for (int i = threadIdx.x; i < CPC * 2 * 32 * CPL; i += blockDim.x) {
        (&wslot_[0][0][0])[i] = Slot{0.f, 0.f, 0.f, __float_as_int(-1.0f)};
        (&wkey_[0][0][0])[i] = 0xFFFFFFFFu;
    }
    __syncthreads();
    if (CLUSTER > 1) {
        if (threadIdx.x == 0) {
            for (int i = 0; i < CPC * 2; ++i) f_mbar_init(f_smem_u32(&xbar_[0][0] + i), 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        cluster_barrier();  // every CTA resident and its mbarriers initialised before any DSMEM traffic
    }

    for (int r = 1; r < m; ++r) {
        const int par = r & 1;
        if (CLUSTER > 1 && ltid == 0)  // this round's phase completes after ncand x (16+4) bytes have landed
            f_mbar_expect_tx(f_smem_u32(&xbar[par]), (uint32_t)ncand * 20u);
        if (PROFILE) t0 = clock64();
        if (PACKED) {
            const unsigned long long xx = f2_pack(x1, x1), yy = f2_pack(y1, y1), zz = f2_pack(z1, z1);
#pragma unroll
            for (int i = 0; i < NP; ++i) {
                const unsigned long long dx = f2_sub(qx[i], xx), dy = f2_sub(qy[i], yy), dz = f2_sub(qz[i], zz);
                unsigned long long t = f2_mul(dy, dy);  // same order as the reference's compiled code: dy*dy, fma dx, fma dz
                t = f2_fma(dx, dx, t);
                t = f2_fma(dz, dz, t);
                float d0, d1;
                f2_unpack(t, d0, d1);
                td[2 * i] = fminf(d0, td[2 * i]);
                td[2 * i + 1] = fminf(d1, td[2 * i + 1]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < PPT; ++j) {
                float d = sqdist_fma(px[j], py[j], pz[j], x1, y1, z1);
                td[j] = fminf(d, td[j]);
            }
        }
        float best;
        int bj;
        Tourney<0, PPT>::run(td, best, bj);
        if (PROFILE) { asm volatile("" ::"f"(best), "r"(bj)); t1 = clock64(); }
        Cand c;
        c.dbits = __float_as_int(best);
        c.key = fps_key(gtid + bj * T);
        // every lane starts fetching its own best point's coordinates; the load overlaps the warp reduce
        const float4 q = sx[bj * tpc + ltid];
        int wm = __reduce_max_sync(GSPN_FULL_MASK, c.dbits);
        unsigned kk = (c.dbits == wm) ? c.key : 0xFFFFFFFFu;
        unsigned wk = __reduce_min_sync(GSPN_FULL_MASK, kk);
        if (PROFILE) { asm volatile("" ::"r"(wk)); t2 = clock64(); }
        if (CLUSTER == 1) {
            if (kk == wk) {
                wslot[par][warp] = Slot{q.x, q.y, q.z, wm};
                wkey[par][warp] = wk;
            }
            __syncthreads();
        } else {
            // all-to-all: the winner's (x,y,z,d | key) goes straight into every CTA's table
            const int src = __ffs(__ballot_sync(GSPN_FULL_MASK, kk == wk)) - 1;
            const float qx = __shfl_sync(GSPN_FULL_MASK, q.x, src);
            const float qy = __shfl_sync(GSPN_FULL_MASK, q.y, src);
            const float qz = __shfl_sync(GSPN_FULL_MASK, q.z, src);
            if (lane < CLUSTER) {
                const int slot = rank * nwarps + warp;
                const uint32_t rbar = map_to_rank(f_smem_u32(&xbar[par]), lane);
                st_async_v4(map_to_rank(f_smem_u32(&wslot[par][slot]), lane), __float_as_uint(qx), __float_as_uint(qy),
                            __float_as_uint(qz), (uint32_t)wm, rbar);
                st_async_b32(map_to_rank(f_smem_u32(&wkey[par][slot]), lane), wk, rbar);
            }
            f_mbar_wait(f_smem_u32(&xbar[par]), (uint32_t)(((r - 1) >> 1) & 1));  // barrier par serves rounds par, par+2, ...
        }
        if (PROFILE) t3 = clock64();
        // every warp reduces the candidate table: CPL unconditional loads per lane, then one warp argmax
        Cand w;
        {
            Slot s0 = wslot[par][lane];
            w.dbits = s0.dbits; w.key = wkey[par][lane]; w.x = s0.x; w.y = s0.y; w.z = s0.z;
        }
#pragma unroll
        for (int i = 1; i < CPL; ++i) {
            Slot s = wslot[par][lane + 32 * i];
            unsigned k2 = wkey[par][lane + 32 * i];
            bool take = s.dbits > w.dbits || (s.dbits == w.dbits && k2 < w.key);
            w.dbits = take ? s.dbits : w.dbits; w.key = take ? k2 : w.key;
            w.x = take ? s.x : w.x; w.y = take ? s.y : w.y; w.z = take ? s.z : w.z;
        }
        c = warp_argmax(w);
        x1 = c.x; y1 = c.y; z1 = c.z;
        if (gtid == 0 && live) out[(size_t)cloud * m + r] = fps_unkey(c.key);
        if (PROFILE) {
            asm volatile("" ::"f"(x1));
            long long t4 = clock64();
            acc[0] += t1 - t0; acc[1] += t2 - t1; acc[2] += t3 - t2; acc[3] += t4 - t3;
        }
    }
    if (PROFILE && gtid == 0 && cloud == 0 && live && prof) {
        prof[0] = acc[0]; prof[1] = acc[1]; prof[2] = acc[2]; prof[3] = acc[3];
    }
    if (CLUSTER > 1) cluster_barrier();  // no CTA exits while a peer may still address its smem
}

// Fallback for clouds that do not fit the register-resident kernel: one CTA per cloud, points
// re-read from global/L2 every round, running min-distance in caller workspace (b*n floats) --
// the reference's own structure with the warp-level reduction above.
__global__ void __launch_bounds__(1024, 1) fps_stream_kernel(int n, int m, const float *__restrict__ xyz, float *__restrict__ temp,
                                                             int *__restrict__ out) {
    __shared__ Slot wslot[2][kMaxWarps];
    __shared__ unsigned wkey[2][kMaxWarps];
    const int cloud = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const float *p = xyz + (size_t)cloud * n * 3;
    float *td = temp + (size_t)cloud * n;
    for (int k = threadIdx.x; k < n; k += blockDim.x) td[k] = 1e38f;
    float x1 = __ldg(p), y1 = __ldg(p + 1), z1 = __ldg(p + 2);
    if (threadIdx.x == 0) out[(size_t)cloud * m] = 0;
    for (int r = 1; r < m; ++r) {
        const int par = r & 1;
        Cand c;
        float best = -1.0f;
        c.key = fps_key(threadIdx.x); c.x = c.y = c.z = 0.f;
        // blockDim.x is a multiple of 512, so k mod 512 is constant per thread and keys ascend with k
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
            float x = __ldg(p + 3 * k), y = __ldg(p + 3 * k + 1), z = __ldg(p + 3 * k + 2);
            float d = sqdist_fma(x, y, z, x1, y1, z1);
            float o = td[k];
            float t = fminf(d, o);
            if (t != o) td[k] = t;
            if (t > best) { best = t; c.key = fps_key(k); c.x = x; c.y = y; c.z = z; }
        }
        c.dbits = __float_as_int(best);
        c = warp_argmax(c);
        if (lane == 0) {
            wslot[par][warp] = Slot{c.x, c.y, c.z, c.dbits};
            wkey[par][warp] = c.key;
        }
        __syncthreads();
        Cand w;
        if (lane < nwarps) {
            Slot s = wslot[par][lane];
            w.dbits = s.dbits; w.key = wkey[par][lane]; w.x = s.x; w.y = s.y; w.z = s.z;
        } else {
            w.dbits = __float_as_int(-1.0f); w.key = 0xFFFFFFFFu; w.x = w.y = w.z = 0.f;
        }
        c = warp_argmax(w);
        x1 = c.x; y1 = c.y; z1 = c.z;
        if (threadIdx.x == 0) out[(size_t)cloud * m + r] = fps_unkey(c.key);
    }
}

// Data-prep sized clouds (data_prep.py:65-91: 1e5-5e5 mesh vertices -> 30000 samples): a 16-CTA cluster of 1024 threads keeps the
// running min-distances of up to 524288 points in registers (32 per thread); the coordinates (12 B/point, L2 resident) are
// streamed every round, 1/16 of the cloud per SM.  Two-level reduce: warps -> CTA winner through shared memory, then the
// 16 CTA winners all-to-all with st.async + mbarrier exactly like the resident kernel.
constexpr int kBigCluster = 16, kBigThreads = 1024, kBigPPT = 32;
__global__ void __launch_bounds__(kBigThreads, 1) fps_cluster_stream_kernel(int n, int m, const float *__restrict__ xyz, int *__restrict__ out) {
    __shared__ Slot wslot[2][32];
    __shared__ unsigned wkey[2][32];
    __shared__ Slot cslot[2][32];
    __shared__ unsigned ckey[2][32];
    __shared__ __align__(8) uint64_t xbar[2];
    const int cloud = blockIdx.y;
    const unsigned rank = cg::this_cluster().block_rank();
    const int T = kBigThreads * kBigCluster;  // multiple of 512: a thread's points share k mod 512, keys ascend with j
    const int gtid = rank * kBigThreads + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *p = xyz + (size_t)cloud * n * 3;
    float td[kBigPPT];
#pragma unroll
    for (int j = 0; j < kBigPPT; ++j) td[j] = (gtid + j * T < n) ? 1e38f : -1.0f;
    for (int i = threadIdx.x; i < 2 * 32; i += kBigThreads) {
        (&wslot[0][0])[i] = Slot{0.f, 0.f, 0.f, __float_as_int(-1.0f)}; (&wkey[0][0])[i] = 0xFFFFFFFFu;
        (&cslot[0][0])[i] = Slot{0.f, 0.f, 0.f, __float_as_int(-1.0f)}; (&ckey[0][0])[i] = 0xFFFFFFFFu;
    }
    if (threadIdx.x == 0) {
        f_mbar_init(f_smem_u32(&xbar[0]), 1);
        f_mbar_init(f_smem_u32(&xbar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    float x1 = __ldg(p), y1 = __ldg(p + 1), z1 = __ldg(p + 2);
    if (gtid == 0) out[(size_t)cloud * m] = 0;
    __syncthreads();
    cluster_barrier();
    for (int r = 1; r < m; ++r) {
        const int par = r & 1;
        if (threadIdx.x == 0) f_mbar_expect_tx(f_smem_u32(&xbar[par]), (uint32_t)kBigCluster * 20u);
        float best = -1.0f, bx = 0.f, by = 0.f, bz = 0.f;
        int bj = 0;
#pragma unroll
        for (int j = 0; j < kBigPPT; ++j) {
            const int k = gtid + j * T;
            if (k < n) {
                const float x = __ldg(p + 3 * (size_t)k), y = __ldg(p + 3 * (size_t)k + 1), z = __ldg(p + 3 * (size_t)k + 2);
                const float t = fminf(sqdist_fma(x, y, z, x1, y1, z1), td[j]);
                td[j] = t;
                if (t > best) { best = t; bj = j; bx = x; by = y; bz = z; }
            }
        }
        Cand c;
        c.dbits = __float_as_int(best); c.key = fps_key(gtid + bj * T); c.x = bx; c.y = by; c.z = bz;
        c = warp_argmax(c);
        if (lane == 0) { wslot[par][warp] = Slot{c.x, c.y, c.z, c.dbits}; wkey[par][warp] = c.key; }
        __syncthreads();
        if (warp == 0) {
            Cand w;
            Slot s0 = wslot[par][lane];
            w.dbits = s0.dbits; w.key = wkey[par][lane]; w.x = s0.x; w.y = s0.y; w.z = s0.z;
            w = warp_argmax(w);
            if (lane < kBigCluster) {
                const uint32_t rbar = map_to_rank(f_smem_u32(&xbar[par]), lane);
                st_async_v4(map_to_rank(f_smem_u32(&cslot[par][rank]), lane), __float_as_uint(w.x), __float_as_uint(w.y), __float_as_uint(w.z),
                            (uint32_t)w.dbits, rbar);
                st_async_b32(map_to_rank(f_smem_u32(&ckey[par][rank]), lane), w.key, rbar);
            }
        }
        f_mbar_wait(f_smem_u32(&xbar[par]), (uint32_t)(((r - 1) >> 1) & 1));
        Cand w;
        Slot s0 = cslot[par][lane];
        w.dbits = s0.dbits; w.key = ckey[par][lane]; w.x = s0.x; w.y = s0.y; w.z = s0.z;
        c = warp_argmax(w);
        x1 = c.x; y1 = c.y; z1 = c.z;
        if (gtid == 0) out[(size_t)cloud * m + r] = fps_unkey(c.key);
    }
    cluster_barrier();
}

template <int PPT, int CLUSTER, int MAXT, bool PROFILE, int CPL, int CPC = 1>
static int launch_resident_cpl(int b, int n, int m, const float *inp, int *out, int threads, cudaStream_t s, long long *prof) {
    auto kern = fps_resident_kernel<PPT, CLUSTER, MAXT, PROFILE, CPL, CPC>;
    if (CLUSTER > 8) GSPN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CLUSTER, (b + CPC - 1) / CPC, 1);
    cfg.blockDim = dim3(threads * CPC, 1, 1);
    const size_t smem = sizeof(float4) * (size_t)PPT * threads * CPC;
    if (smem > 48 * 1024) GSPN_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CLUSTER;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    GSPN_CUDA_OK(cudaLaunchKernelEx(&cfg, kern, n, m, inp, out, prof, b));
    return GSPN_OK;
}

template <int PPT, int CLUSTER, int MAXT, bool PROFILE = false>
static int launch_resident(int b, int n, int m, const float *inp, int *out, int threads, cudaStream_t s, long long *prof = nullptr) {
    const int ncand = CLUSTER * (threads / 32);
    if (CLUSTER == 1 || ncand <= 32) return launch_resident_cpl<PPT, CLUSTER, MAXT, PROFILE, 1>(b, n, m, inp, out, threads, s, prof);
    if (ncand <= 64) return launch_resident_cpl<PPT, CLUSTER, MAXT, PROFILE, 2>(b, n, m, inp, out, threads, s, prof);
    if (!PROFILE && ncand <= 128) return launch_resident_cpl<PPT, CLUSTER, MAXT, false, 4>(b, n, m, inp, out, threads, s, prof);
    return GSPN_E_UNSUPPORTED;
}

template <int PPT, int MAXT>
static int dispatch_cluster(int cluster, int b, int n, int m, const float *inp, int *out, int threads, cudaStream_t s) {
    switch (cluster) {
        case 1: return launch_resident<PPT, 1, MAXT>(b, n, m, inp, out, threads, s);
        case 2: return launch_resident<PPT, 2, MAXT>(b, n, m, inp, out, threads, s);
        case 4: return launch_resident<PPT, 4, MAXT>(b, n, m, inp, out, threads, s);
        case 8: return launch_resident<PPT, 8, MAXT>(b, n, m, inp, out, threads, s);
        case 16: return launch_resident<PPT, 16, MAXT>(b, n, m, inp, out, threads, s);
    }
    return GSPN_E_UNSUPPORTED;
}

static int g_fps_pack = 1;  // gspn_fps_tune_pack: clouds per CTA of the (128, 32, 8) mapping
static int launch_cfg(int b, int n, int m, const float *inp, int *out, int threads, int ppt, int cluster, cudaStream_t s) {
    if (g_fps_pack == 2 && threads == 128 && ppt == 32 && cluster == 8 && b >= 2 && b <= 65535 && (long)threads * ppt * cluster >= n)
        return launch_resident_cpl<32, 8, 256, false, 1, 2>(b, n, m, inp, out, threads, s, nullptr);
    if (threads < 32 || threads > 1024 || threads % 32) return GSPN_E_UNSUPPORTED;
    if (ppt > 1 && (threads * cluster) % 512) return GSPN_E_UNSUPPORTED;  // ascending keys per thread
    if ((long)threads * ppt * cluster < n) return GSPN_E_UNSUPPORTED;
    if (cluster > 1 && cluster * (threads / 32) > kMaxCand) return GSPN_E_UNSUPPORTED;  // all-to-all candidate table
    if ((size_t)threads * ppt * 16 > 200 * 1024) return GSPN_E_UNSUPPORTED;
    if (b > 65535) return GSPN_E_UNSUPPORTED;
    switch (ppt) {
        case 1: return dispatch_cluster<1, 1024>(cluster, b, n, m, inp, out, threads, s);
        case 2: return dispatch_cluster<2, 1024>(cluster, b, n, m, inp, out, threads, s);
        case 4: return dispatch_cluster<4, 1024>(cluster, b, n, m, inp, out, threads, s);
        case 8: return dispatch_cluster<8, 1024>(cluster, b, n, m, inp, out, threads, s);
        case 16: return threads <= 512 ? dispatch_cluster<16, 512>(cluster, b, n, m, inp, out, threads, s) : GSPN_E_UNSUPPORTED;
        case 32: return threads <= 256 ? dispatch_cluster<32, 256>(cluster, b, n, m, inp, out, threads, s) : GSPN_E_UNSUPPORTED;
    }
    return GSPN_E_UNSUPPORTED;
}

// (threads, ppt, cluster) by cloud size; see DESIGN.md "FPS" for the measurements behind the table.
static void choose_cfg(int n, int *threads, int *ppt, int *cluster) {
    if (n <= 512) { *threads = ((n + 31) / 32) * 32; *ppt = 1; *cluster = 1; return; }
    if (n <= 1024) { *threads = 512; *ppt = 2; *cluster = 1; return; }
    if (n <= 2048) { *threads = 512; *ppt = 4; *cluster = 1; return; }
    // measured on B200 (tools/fps_sweep.py, profiles/r01_fps_sweep.txt): few fat warps win -- one warp per SM
    // sub-partition keeps the candidate exchange small, registers hold 32 points per thread
    if (n <= 4096) { *threads = 128; *ppt = 8; *cluster = 4; return; }
    if (n <= 8192) { *threads = 128; *ppt = 16; *cluster = 4; return; }
    if (n <= 16384) { *threads = 128; *ppt = 16; *cluster = 8; return; }
    if (n <= 32768) { *threads = 128; *ppt = 32; *cluster = 8; return; }
    if (n <= 65536) { *threads = 256; *ppt = 32; *cluster = 8; return; }
    *threads = 256; *ppt = 32; *cluster = 16;  // up to 131072 (non-portable cluster size)
}

constexpr int kMaxResident = 512 * 16 * 16;
constexpr int kMaxClusterStream = kBigCluster * kBigThreads * kBigPPT;  // 524288

}  // namespace gspn

using namespace gspn;

// fps_bucket.cu: the bucket-pruned single-CTA kernel for 8193 .. 32768 points
size_t gspn_fps_bucket_workspace_bytes(int b, int n);
int gspn_fps_bucket_launch(int b, int n, int m, const float *inp, int *out, void *workspace, long long *prof, cudaStream_t s);
// Opt-in (gspn_fps_tune(1)): exact and 80x fewer distance evaluations, but measured SLOWER per cloud than the full-scan cluster kernel
// (2.4 ms vs 1.04 ms at 32768 -> 2048): off by default.
// fps_pruned.cu (gspn_fps_tune(2)): the same pruning on the 8-CTA cluster (points and distances stay in registers, statically indexed).
// Also exact, also slower than the full scan (1.5 ms): a round's critical path is the one warp whose buckets the sample touches.
size_t gspn_fps_pruned_workspace_bytes(int b, int n);
int gspn_fps_pruned_launch(int b, int n, int m, const float *inp, int *out, void *workspace, long long *prof, cudaStream_t s);
static int g_fps_mode = 0;  // gspn_fps_tune: 0 = full-scan kernels (default), 1 = single-CTA bucket kernel, 2 = pruned cluster kernel
#define g_fps_buckets (g_fps_mode == 1)
static int g_fps_big[3] = {0, 0, 0};  // tuning door (gspn_fps_tune_mapping): (threads, ppt, cluster) for clouds above 16384 points

// Tuning door: per-phase cycle counts of thread 0 (compute+tournament, warp reduce, exchange, table reduce),
// summed over the m-1 rounds, for the (threads, ppt, cluster) shapes the default table uses.
extern "C" int gspn_fps_profile(int b, int n, int m, const float *inp, int *out, int threads, int ppt, int cluster, long long *prof4,
                                gspn_stream_t stream) {
    GSPN_REQUIRE(b > 0 && n > 0 && m > 0);
    GSPN_REQUIRE_PTR(inp); GSPN_REQUIRE_PTR(out); GSPN_REQUIRE_PTR(prof4);
    if ((long)threads * ppt * cluster < n || (cluster > 1 && cluster * (threads / 32) > kMaxCand)) return GSPN_E_UNSUPPORTED;
    cudaStream_t s = as_stream(stream);
    int rc = GSPN_E_UNSUPPORTED;
    if (ppt == 32 && cluster == 8 && threads <= 256) rc = launch_resident<32, 8, 256, true>(b, n, m, inp, out, threads, s, prof4);
    else if (ppt == 16 && cluster == 8 && threads <= 512) rc = launch_resident<16, 8, 512, true>(b, n, m, inp, out, threads, s, prof4);
    else if (ppt == 4 && cluster == 1) rc = launch_resident<4, 1, 1024, true>(b, n, m, inp, out, threads, s, prof4);
    else if (ppt == 4 && cluster == 8) rc = launch_resident<4, 8, 1024, true>(b, n, m, inp, out, threads, s, prof4);
    else if (ppt == 4 && cluster == 16) rc = launch_resident<4, 16, 1024, true>(b, n, m, inp, out, threads, s, prof4);
    if (rc != GSPN_OK) return rc;
    return check_launch();
}

extern "C" int gspn_fps_max_resident_points(void) { return kMaxResident; }

extern "C" size_t gspn_farthest_point_sample_workspace_bytes(int b, int n, int m) {
    (void)m;
    if (b <= 0) return 0;
    if (g_fps_buckets)
        if (const size_t wb = gspn_fps_bucket_workspace_bytes(b, n)) return wb;  // the sorted copy of the bucket-pruned kernel
    if (g_fps_mode == 2)
        if (const size_t wb = gspn_fps_pruned_workspace_bytes(b, n)) return wb;  // the curve-ordered copy
    if (n <= kMaxClusterStream) return 0;
    return sizeof(float) * (size_t)b * (size_t)n;
}

extern "C" void gspn_fps_tune(int mode) { g_fps_mode = (mode == 1 || mode == 2) ? mode : 0; }
extern "C" void gspn_fps_tune_pack(int clouds_per_cta) { g_fps_pack = clouds_per_cta == 2 ? 2 : 1; }
extern "C" void gspn_fps_tune_mapping(int threads, int ppt, int cluster) { g_fps_big[0] = threads; g_fps_big[1] = ppt; g_fps_big[2] = cluster; }

extern "C" int gspn_fps_bucket_profile(int b, int n, int m, const float *inp, int *out, void *workspace, size_t workspace_bytes,
                                       long long *prof3, gspn_stream_t stream) {
    GSPN_REQUIRE(b > 0 && n > 0 && m > 0);
    GSPN_REQUIRE_PTR(inp); GSPN_REQUIRE_PTR(out); GSPN_REQUIRE_PTR(prof3);
    const size_t need = gspn_fps_bucket_workspace_bytes(b, n);
    if (need == 0) return GSPN_E_UNSUPPORTED;
    if (workspace == nullptr || workspace_bytes < need) return GSPN_E_WORKSPACE;
    return gspn_fps_bucket_launch(b, n, m, inp, out, workspace, prof3, as_stream(stream));
}

extern "C" int gspn_fps_pruned_profile(int b, int n, int m, const float *inp, int *out, void *workspace, size_t workspace_bytes,
                                       long long *prof5, gspn_stream_t stream) {
    GSPN_REQUIRE(b > 0 && n > 0 && m > 0);
    GSPN_REQUIRE_PTR(inp); GSPN_REQUIRE_PTR(out); GSPN_REQUIRE_PTR(prof5);
    const size_t need = gspn_fps_pruned_workspace_bytes(b, n);
    if (need == 0) return GSPN_E_UNSUPPORTED;
    if (workspace == nullptr || workspace_bytes < need) return GSPN_E_WORKSPACE;
    return gspn_fps_pruned_launch(b, n, m, inp, out, workspace, prof5, as_stream(stream));
}

extern "C" int gspn_farthest_point_sample_cfg(int b, int n, int m, const float *inp, int *out, int threads, int ppt, int cluster,
                                              gspn_stream_t stream) {
    GSPN_REQUIRE(b >= 0 && n > 0 && m > 0);  // npoint>0 tf_sampling.cpp:99; rank/shape :105
    if (b == 0) return GSPN_OK;
    GSPN_REQUIRE_PTR(inp); GSPN_REQUIRE_PTR(out);
    int t0, p0, c0;
    choose_cfg(n, &t0, &p0, &c0);
    if (n > 16384 && g_fps_big[0] > 0 && (long)g_fps_big[0] * g_fps_big[1] * g_fps_big[2] >= n) { t0 = g_fps_big[0]; p0 = g_fps_big[1]; c0 = g_fps_big[2]; }
    if (threads <= 0) threads = t0;
    if (ppt <= 0) ppt = p0;
    if (cluster <= 0) cluster = c0;
    int rc = launch_cfg(b, n, m, inp, out, threads, ppt, cluster, as_stream(stream));
    if (rc != GSPN_OK) return rc;
    return check_launch();
}

extern "C" int gspn_farthest_point_sample(int b, int n, int m, const float *inp, int *out, void *workspace, size_t workspace_bytes,
                                          gspn_stream_t stream) {
    GSPN_REQUIRE(b >= 0 && n > 0 && m > 0);
    if (b == 0) return GSPN_OK;
    GSPN_REQUIRE_PTR(inp); GSPN_REQUIRE_PTR(out);
    if (g_fps_buckets && workspace != nullptr) {
        const size_t need = gspn_fps_bucket_workspace_bytes(b, n);
        if (need && workspace_bytes >= need) return gspn_fps_bucket_launch(b, n, m, inp, out, workspace, nullptr, as_stream(stream));
    }
    if (g_fps_mode == 2 && workspace != nullptr) {
        const size_t need = gspn_fps_pruned_workspace_bytes(b, n);
        if (need && workspace_bytes >= need) return gspn_fps_pruned_launch(b, n, m, inp, out, workspace, nullptr, as_stream(stream));
    }
    if (n <= kMaxResident) return gspn_farthest_point_sample_cfg(b, n, m, inp, out, 0, 0, 0, stream);
    if (n <= kMaxClusterStream && b <= 65535) {
        GSPN_CUDA_OK(cudaFuncSetAttribute(fps_cluster_stream_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(kBigCluster, b, 1);
        cfg.blockDim = dim3(kBigThreads, 1, 1);
        cfg.stream = as_stream(stream);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = kBigCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        GSPN_CUDA_OK(cudaLaunchKernelEx(&cfg, fps_cluster_stream_kernel, n, m, inp, out));
        return check_launch();
    }
    size_t need = gspn_farthest_point_sample_workspace_bytes(b, n, m);
    if (workspace == nullptr || workspace_bytes < need) return GSPN_E_WORKSPACE;
    fps_stream_kernel<<<b, 1024, 0, as_stream(stream)>>>(n, m, inp, (float *)workspace, out);
    return check_launch();
}
