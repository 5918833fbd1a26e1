// fps_pruned.cu -- exact bucket-pruned farthest point sampling on a thread-block cluster, for clouds of 8193 .. 32768 points
// (SA1 of the model: 32768 -> 2048).  Same results, bit for bit, as farthestpointsamplingKernel (tf_sampling_g.cu:105-170) and as the
// full-scan cluster kernel in fps.cu, whose exchange it keeps; what goes away is the per-round distance update of ALL points.
//
// A round of FPS lowers the running min-distance of the points near the new sample and leaves everything else untouched.  The cloud
// is sorted once along a Hilbert curve over a 32^3 grid (fps_curve_sort_kernel, a counting sort in shared memory) and cut into 1024
// buckets of 32 consecutive points.  The cluster is 8 CTAs x 4 warps = 32 warps; warp w owns buckets w, 32+w, 64+w, ... (the buckets
// one sample touches are neighbours on the curve, so they land in different warps), one bucket per register SLOT j = 0..31, lane =
// point: every thread still keeps 32 points and their running distances in registers for the whole kernel.  Lane j of a warp also
// keeps the state of the warp's bucket j: bounding box, current max distance, tie-break key and coordinates of that maximum.
// Per round:
//     lane j:   bound_j = the reference's own float distance expression evaluated on the per-axis gaps between the sample and box j
//               (every operation is monotone in |operand|, so bound_j <= the float distance to every point of the bucket);
//               bound_j >= max distance of the bucket  =>  min(d, d_new) = d for all its points: skip, bit-exactly.
//     ballot -> the touched slots, warp-uniform; the slot loop is fully unrolled (static register indices) behind a two-level skip,
//               so an untouched warp pays the box test, one ballot and one branch.
//     touched bucket: one distance per lane, redux.sync max, winner lane (key tie-break only on an exact tie), state to lane j.
//     the warp's candidate = best of its 32 bucket maxima, recomputed only when one of them moved; then the same st.async /
//     mbarrier all-to-all and table reduce as fps.cu.
// Measured on config 2: 12.9 bucket updates per round in the whole cluster instead of 1024 -- and 1.55 ms instead of 1.06 ms: every
// round some warp holds the sample's bucket and runs update -> bucket argmax -> warp argmax -> send (~650 dependent cycles) while the
// other 31 wait in the exchange; the full scan's 394 cycles of pipelined FFMA2 are a shorter critical path.  Opt-in
// (gspn_fps_tune(2)), kept with its tests as the record of that result.
#include "fps_common.cuh"

namespace cg = cooperative_groups;

namespace gspn {

constexpr int kFpCluster = 8, kFpThreads = 128, kFpWarps = kFpCluster * kFpThreads / 32;  // 32 cluster warps
constexpr int kFpSlots = 32;                                                              // buckets per warp
constexpr int kFpBuckets = kFpWarps * kFpSlots;                                           // 1024
constexpr int kFpMaxPoints = kFpBuckets * 32;                                             // 32768
constexpr int kFpCells = 32 * 32 * 32;
constexpr int kFpSortThreads = 1024;
constexpr int kFpCellsPerThread = kFpCells / kFpSortThreads;

// float <-> int with the same order (redux.sync min / max on coordinates of either sign)
__device__ __forceinline__ int fp_ord(float f) { int b = __float_as_int(f); return b ^ ((b >> 31) & 0x7FFFFFFF); }
__device__ __forceinline__ float fp_unord(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7FFFFFFF)); }
__device__ __forceinline__ unsigned fp_spread5(unsigned v) {  // bit i -> bit 3i
    return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4) | ((v & 8u) << 6) | ((v & 16u) << 8);
}
// cell (X0,X1,X2 in 0..31) -> its number along a Hilbert curve (Skilling's transpose form): consecutive cells are face neighbours,
// so 32 consecutive points of the sorted cloud form a compact bucket
__device__ __forceinline__ unsigned fp_hilbert(unsigned X0, unsigned X1, unsigned X2) {
#pragma unroll
    for (unsigned Q = 16; Q > 1; Q >>= 1) {
        const unsigned P = Q - 1;
        if (X0 & Q) X0 ^= P;
        if (X1 & Q) X0 ^= P; else { const unsigned t = (X0 ^ X1) & P; X0 ^= t; X1 ^= t; }
        if (X2 & Q) X0 ^= P; else { const unsigned t = (X0 ^ X2) & P; X0 ^= t; X2 ^= t; }
    }
    X1 ^= X0; X2 ^= X1;
    unsigned t = 0;
#pragma unroll
    for (unsigned Q = 16; Q > 1; Q >>= 1)
        if (X2 & Q) t ^= Q - 1;
    X0 ^= t; X1 ^= t; X2 ^= t;
    return (fp_spread5(X0) << 2) | (fp_spread5(X1) << 1) | fp_spread5(X2);
}

// One CTA per cloud: (x, y, z, original index) of every point, in curve order, into sorted[cloud][0..n).
__global__ void __launch_bounds__(kFpSortThreads, 1) fps_curve_sort_kernel(int n, const float *__restrict__ xyz, float4 *__restrict__ sorted) {
    extern __shared__ int cellcnt[];  // kFpCells
    __shared__ float red[kFpSortThreads / 32][6];
    __shared__ int wsum[kFpSortThreads / 32];
    const int cloud = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *p = xyz + (size_t)cloud * n * 3;
    float4 *srt = sorted + (size_t)cloud * kFpMaxPoints;
    float lo0 = 3.4e38f, lo1 = 3.4e38f, lo2 = 3.4e38f, hi0 = -3.4e38f, hi1 = -3.4e38f, hi2 = -3.4e38f;
    for (int k = tid; k < n; k += kFpSortThreads) {
        const float x = __ldg(p + 3 * k), y = __ldg(p + 3 * k + 1), z = __ldg(p + 3 * k + 2);
        lo0 = fminf(lo0, x); hi0 = fmaxf(hi0, x); lo1 = fminf(lo1, y); hi1 = fmaxf(hi1, y); lo2 = fminf(lo2, z); hi2 = fmaxf(hi2, z);
    }
    lo0 = fp_unord(__reduce_min_sync(GSPN_FULL_MASK, fp_ord(lo0))); hi0 = fp_unord(__reduce_max_sync(GSPN_FULL_MASK, fp_ord(hi0)));
    lo1 = fp_unord(__reduce_min_sync(GSPN_FULL_MASK, fp_ord(lo1))); hi1 = fp_unord(__reduce_max_sync(GSPN_FULL_MASK, fp_ord(hi1)));
    lo2 = fp_unord(__reduce_min_sync(GSPN_FULL_MASK, fp_ord(lo2))); hi2 = fp_unord(__reduce_max_sync(GSPN_FULL_MASK, fp_ord(hi2)));
    if (lane == 0) { red[warp][0] = lo0; red[warp][1] = lo1; red[warp][2] = lo2; red[warp][3] = hi0; red[warp][4] = hi1; red[warp][5] = hi2; }
    for (int i = tid; i < kFpCells; i += kFpSortThreads) cellcnt[i] = 0;
    __syncthreads();
    {
        const float a = red[lane][0], b = red[lane][1], c = red[lane][2], d = red[lane][3], e = red[lane][4], f = red[lane][5];
        lo0 = fp_unord(__reduce_min_sync(GSPN_FULL_MASK, fp_ord(a))); lo1 = fp_unord(__reduce_min_sync(GSPN_FULL_MASK, fp_ord(b)));
        lo2 = fp_unord(__reduce_min_sync(GSPN_FULL_MASK, fp_ord(c))); hi0 = fp_unord(__reduce_max_sync(GSPN_FULL_MASK, fp_ord(d)));
        hi1 = fp_unord(__reduce_max_sync(GSPN_FULL_MASK, fp_ord(e))); hi2 = fp_unord(__reduce_max_sync(GSPN_FULL_MASK, fp_ord(f)));
    }
    const float e0 = hi0 - lo0, e1 = hi1 - lo1, e2 = hi2 - lo2;
    const float inv0 = (e0 > 0.f && e0 < 3e38f) ? 32.f / e0 : 0.f, inv1 = (e1 > 0.f && e1 < 3e38f) ? 32.f / e1 : 0.f,
                inv2 = (e2 > 0.f && e2 < 3e38f) ? 32.f / e2 : 0.f;
    auto cell_of = [=](float x, float y, float z) -> unsigned {
        return fp_hilbert((unsigned)min(31, max(0, (int)((x - lo0) * inv0))), (unsigned)min(31, max(0, (int)((y - lo1) * inv1))),
                          (unsigned)min(31, max(0, (int)((z - lo2) * inv2))));
    };
    // a thread's points are the same in the histogram and in the scatter pass: their cells stay in registers between the two
    constexpr int kPer = kFpMaxPoints / kFpSortThreads;  // 32
    unsigned mycell[kPer / 2];  // two 15-bit cells per register
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
        const int k = tid + i * kFpSortThreads;
        unsigned c = 0;
        if (k < n) {
            c = cell_of(__ldg(p + 3 * k), __ldg(p + 3 * k + 1), __ldg(p + 3 * k + 2));
            atomicAdd(&cellcnt[c], 1);
        }
        mycell[i / 2] = (i & 1) ? (mycell[i / 2] | (c << 16)) : c;
    }
    __syncthreads();
    {  // exclusive scan: kFpCellsPerThread consecutive cells per thread
        const int c_lo = tid * kFpCellsPerThread;
        int s = 0;
#pragma unroll 8
        for (int i = 0; i < kFpCellsPerThread; ++i) s += cellcnt[c_lo + ((i + tid) & (kFpCellsPerThread - 1))];  // rotated: no bank conflicts
        int incl = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(GSPN_FULL_MASK, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        const int ws = wsum[lane];  // 32 warps
        int wincl = ws;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(GSPN_FULL_MASK, wincl, d);
            if (lane >= d) wincl += t;
        }
        int base = incl - s + __shfl_sync(GSPN_FULL_MASK, wincl - ws, warp);
        for (int i = 0; i < kFpCellsPerThread; ++i) { const int c = cellcnt[c_lo + i]; cellcnt[c_lo + i] = base; base += c; }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kPer; ++i) {
        const int k = tid + i * kFpSortThreads;
        if (k < n) {
            const int pos = atomicAdd(&cellcnt[(mycell[i / 2] >> (16 * (i & 1))) & 0xFFFFu], 1);
            srt[pos] = make_float4(__ldg(p + 3 * k), __ldg(p + 3 * k + 1), __ldg(p + 3 * k + 2), __int_as_float(k));
        }
    }
}

// the reference's distance expression on the gaps between a sample and a box: a lower bound, in the same float arithmetic, of the
// distance from the sample to any point inside the box
__device__ __forceinline__ float fp_box_bound(float lx, float ly, float lz, float hx, float hy, float hz, float sx, float sy, float sz) {
    const float gx = fmaxf(fmaxf(__fsub_rn(lx, sx), __fsub_rn(sx, hx)), 0.f);
    const float gy = fmaxf(fmaxf(__fsub_rn(ly, sy), __fsub_rn(sy, hy)), 0.f);
    const float gz = fmaxf(fmaxf(__fsub_rn(lz, sz), __fsub_rn(sz, hz)), 0.f);
    float t = __fmul_rn(gy, gy);
    t = __fmaf_rn(gx, gx, t);
    return __fmaf_rn(gz, gz, t);
}

// (max distance, then min key) over the warp; the key reduction only runs on an exact tie of the maximum.  Every lane returns the
// winner lane and the maximum's bits.
__device__ __forceinline__ int fp_warp_winner(int dbits, unsigned key, int &wm, unsigned &wk) {
    wm = __reduce_max_sync(GSPN_FULL_MASK, dbits);  // non-negative floats order as ints; -1.0f (empty) is negative
    const unsigned eq = __ballot_sync(GSPN_FULL_MASK, dbits == wm);
    int src = __ffs(eq) - 1;
    if (eq & (eq - 1)) {  // tie: the reference's order among equal maxima
        const unsigned kk = (dbits == wm) ? key : 0xFFFFFFFFu;
        wk = __reduce_min_sync(GSPN_FULL_MASK, kk);
        src = __ffs(__ballot_sync(GSPN_FULL_MASK, kk == wk)) - 1;
    } else {
        wk = __shfl_sync(GSPN_FULL_MASK, key, src);
    }
    return src;
}

// grid = (kFpCluster, b), cluster = (kFpCluster,1,1), kFpThreads threads.  sorted: the curve-ordered copy (fps_curve_sort_kernel).
// prof (PROF only, cloud 0): [0..3] thread 0's cycles in box test + bucket updates + candidate, (unused), exchange, table reduce;
// [4] bucket updates over all warps.
template <bool PROF>
__global__ void __launch_bounds__(kFpThreads, 1) fps_pruned_kernel(int n, int m, const float *__restrict__ xyz, const float4 *__restrict__ sorted,
                                                                   int *__restrict__ out, long long *__restrict__ prof) {
    __shared__ Slot wslot[2][kFpWarps];
    __shared__ unsigned wkey[2][kFpWarps];
    __shared__ __align__(8) uint64_t xbar[2];
    __shared__ unsigned short sidx[kFpSlots][kFpThreads];  // original index of every point this CTA holds

    const int cloud = blockIdx.y;
    const unsigned rank = cg::this_cluster().block_rank();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cw = rank * (kFpThreads / 32) + warp;  // warp of the cluster
    const float4 *srt = sorted + (size_t)cloud * kFpMaxPoints;
    const float *p = xyz + (size_t)cloud * n * 3;

    float px[kFpSlots], py[kFpSlots], pz[kFpSlots], td[kFpSlots];
    // state of the warp's bucket `lane`
    float blx = 0.f, bly = 0.f, blz = 0.f, bhx = 0.f, bhy = 0.f, bhz = 0.f, bwx = 0.f, bwy = 0.f, bwz = 0.f;
    int bmax = __float_as_int(-1.0f);
    unsigned bkey = 0xFFFFFFFFu;
#pragma unroll
    for (int j = 0; j < kFpSlots; ++j) {
        const int pos = (j * kFpWarps + cw) * 32 + lane;
        const bool ok = pos < n;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) v = __ldcg(srt + pos);
        px[j] = v.x; py[j] = v.y; pz[j] = v.z;
        td[j] = ok ? 1e38f : -1.0f;  // tf_sampling_g.cu:118; padding never wins (distances are >= 0)
        sidx[j][tid] = (unsigned short)__float_as_int(v.w);
        const int ilx = __reduce_min_sync(GSPN_FULL_MASK, ok ? fp_ord(v.x) : 0x7FFFFFFF), ihx = __reduce_max_sync(GSPN_FULL_MASK, ok ? fp_ord(v.x) : (int)0x80000000);
        const int ily = __reduce_min_sync(GSPN_FULL_MASK, ok ? fp_ord(v.y) : 0x7FFFFFFF), ihy = __reduce_max_sync(GSPN_FULL_MASK, ok ? fp_ord(v.y) : (int)0x80000000);
        const int ilz = __reduce_min_sync(GSPN_FULL_MASK, ok ? fp_ord(v.z) : 0x7FFFFFFF), ihz = __reduce_max_sync(GSPN_FULL_MASK, ok ? fp_ord(v.z) : (int)0x80000000);
        if (lane == j) {
            blx = fp_unord(ilx); bly = fp_unord(ily); blz = fp_unord(ilz); bhx = fp_unord(ihx); bhy = fp_unord(ihy); bhz = fp_unord(ihz);
            if ((j * kFpWarps + cw) * 32 < n) bmax = __float_as_int(1e38f);  // non-empty: touched by round 1 whatever the sample
        }
    }
    // the warp's cached candidate (uniform across lanes)
    int cdb = __float_as_int(-1.0f);
    unsigned ckey = 0xFFFFFFFFu;
    float ccx = 0.f, ccy = 0.f, ccz = 0.f;

    float x1 = __ldg(p), y1 = __ldg(p + 1), z1 = __ldg(p + 2);  // old = 0 (:114)
    if (rank == 0 && tid == 0) out[(size_t)cloud * m] = 0;
    for (int i = tid; i < 2 * kFpWarps; i += kFpThreads) {
        (&wslot[0][0])[i] = Slot{0.f, 0.f, 0.f, __float_as_int(-1.0f)};
        (&wkey[0][0])[i] = 0xFFFFFFFFu;
    }
    __syncthreads();
    if (tid == 0) {
        f_mbar_init(f_smem_u32(&xbar[0]), 1);
        f_mbar_init(f_smem_u32(&xbar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    cluster_barrier();  // every CTA resident and its mbarriers initialised before any DSMEM traffic

    long long t0 = 0, t1 = 0, t2 = 0, acc0 = 0, acc2 = 0, acc3 = 0;
    unsigned acc_touched = 0;
    for (int r = 1; r < m; ++r) {
        const int par = r & 1;
        if (tid == 0) f_mbar_expect_tx(f_smem_u32(&xbar[par]), (uint32_t)kFpWarps * 20u);
        if (PROF) t0 = clock64();
        const bool touched = fp_box_bound(blx, bly, blz, bhx, bhy, bhz, x1, y1, z1) < __int_as_float(bmax);
        const unsigned mask = __ballot_sync(GSPN_FULL_MASK, touched);
        if (mask) {
            if (PROF) acc_touched += __popc(mask);
            bool dirty = false;
#pragma unroll
            for (int g = 0; g < kFpSlots; g += 8) {
                if ((mask >> g) & 0xFFu) {
#pragma unroll
                    for (int j = g; j < g + 8; ++j) {
                        if ((mask >> j) & 1u) {
                            const unsigned okey = fps_key(sidx[j][tid]);
                            const float dn = sqdist_fma(px[j], py[j], pz[j], x1, y1, z1);
                            const float dd = fminf(dn, td[j]);
                            td[j] = dd;
                            int bm;
                            unsigned bk;
                            const int src = fp_warp_winner(__float_as_int(dd), okey, bm, bk);
                            const float nx = __shfl_sync(GSPN_FULL_MASK, px[j], src), ny = __shfl_sync(GSPN_FULL_MASK, py[j], src),
                                        nz = __shfl_sync(GSPN_FULL_MASK, pz[j], src);
                            if (lane == j) {
                                dirty = dirty || bm != bmax || bk != bkey;
                                bmax = bm; bkey = bk; bwx = nx; bwy = ny; bwz = nz;
                            }
                        }
                    }
                }
            }
            // distances only fall: the warp's candidate stands unless one of its bucket maxima moved
            if (__any_sync(GSPN_FULL_MASK, dirty)) {
                const int src = fp_warp_winner(bmax, bkey, cdb, ckey);
                ccx = __shfl_sync(GSPN_FULL_MASK, bwx, src);
                ccy = __shfl_sync(GSPN_FULL_MASK, bwy, src);
                ccz = __shfl_sync(GSPN_FULL_MASK, bwz, src);
            }
        }
        if (PROF) { asm volatile("" ::"r"(cdb), "r"(ckey)); t1 = clock64(); }
        if (lane < kFpCluster) {
            const uint32_t rbar = map_to_rank(f_smem_u32(&xbar[par]), lane);
            st_async_v4(map_to_rank(f_smem_u32(&wslot[par][cw]), lane), __float_as_uint(ccx), __float_as_uint(ccy), __float_as_uint(ccz),
                        (uint32_t)cdb, rbar);
            st_async_b32(map_to_rank(f_smem_u32(&wkey[par][cw]), lane), ckey, rbar);
        }
        f_mbar_wait(f_smem_u32(&xbar[par]), (uint32_t)(((r - 1) >> 1) & 1));  // barrier par serves rounds par, par+2, ...
        if (PROF) t2 = clock64();
        {
            const Slot s = wslot[par][lane];
            const unsigned k = wkey[par][lane];
            int gm;
            unsigned gk;
            const int src = fp_warp_winner(s.dbits, k, gm, gk);
            x1 = __shfl_sync(GSPN_FULL_MASK, s.x, src);
            y1 = __shfl_sync(GSPN_FULL_MASK, s.y, src);
            z1 = __shfl_sync(GSPN_FULL_MASK, s.z, src);
            if (rank == 0 && tid == 0) out[(size_t)cloud * m + r] = fps_unkey(gk);
        }
        if (PROF) {
            asm volatile("" ::"f"(x1));
            const long long t3 = clock64();
            acc0 += t1 - t0; acc2 += t2 - t1; acc3 += t3 - t2;
        }
    }
    if (PROF && cloud == 0 && prof) {
        if (rank == 0 && tid == 0) { prof[0] = acc0; prof[1] = 0; prof[2] = acc2; prof[3] = acc3; }
        if (lane == 0) atomicAdd((unsigned long long *)prof + 4, (unsigned long long)acc_touched);
    }
    cluster_barrier();  // no CTA exits while a peer may still address its smem
}

}  // namespace gspn

using namespace gspn;

// used by fps.cu's entry points
size_t gspn_fps_pruned_workspace_bytes(int b, int n) {
    if (b <= 0 || b > 65535 || n <= 8192 || n > kFpMaxPoints) return 0;
    return (size_t)b * kFpMaxPoints * sizeof(float4);
}

int gspn_fps_pruned_launch(int b, int n, int m, const float *inp, int *out, void *workspace, long long *prof, cudaStream_t s) {
    static unsigned char attr_done[64];  // per device; benign race
    int dev = 0;
    GSPN_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return GSPN_E_UNSUPPORTED;
    const int sort_smem = kFpCells * (int)sizeof(int);
    if (!attr_done[dev]) {
        GSPN_CUDA_OK(cudaFuncSetAttribute(fps_curve_sort_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sort_smem));
        attr_done[dev] = 1;
    }
    fps_curve_sort_kernel<<<b, kFpSortThreads, sort_smem, s>>>(n, inp, (float4 *)workspace);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kFpCluster, b, 1);
    cfg.blockDim = dim3(kFpThreads, 1, 1);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kFpCluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    if (prof) GSPN_CUDA_OK(cudaLaunchKernelEx(&cfg, fps_pruned_kernel<true>, n, m, inp, (const float4 *)workspace, out, prof));
    else GSPN_CUDA_OK(cudaLaunchKernelEx(&cfg, fps_pruned_kernel<false>, n, m, inp, (const float4 *)workspace, out, prof));
    return check_launch();
}
