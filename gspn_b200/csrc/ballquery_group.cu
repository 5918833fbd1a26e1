// ballquery_group.cu -- query_ball_point (tf_grouping_g.cu:6-39) and the fused
// ball-query + group_point of sample_and_group (utils/pointnet_util.py:40-48): the ordered-scan kernel used for
// clouds below 4096 points (and whenever the caller gives no workspace); larger clouds go through the exact
// uniform-grid search in grid_search.cu, which shares the row writer (group_rows.cuh) and the entry points below.
//
// Reference: one CTA per cloud, one thread per query, each thread streams all n points from
// global memory with a divergent early exit; then a second kernel copies rows one scalar at a
// time.  Here:
//   * a warp owns QPW queries; the cloud streams through shared memory in tiles that are fetched
//     by the TMA bulk-copy engine (cp.async.bulk + mbarrier, double buffered) when the cloud is
//     16-byte aligned, by coalesced loads otherwise;
//   * each lane tests one point per step against the warp's queries; hits are compacted IN INDEX
//     ORDER with __ballot_sync + popc prefix, so "first nsample in index order" and the first-hit
//     back-fill of the reference hold exactly;
//   * the predicate max(sqrtf(s),1e-20f) < radius is evaluated as  !(s > s_max)  with s_max the
//     largest float whose correctly-rounded square root is below radius (computed on the host by
//     bisection over float bit patterns) -- identical decisions, no sqrt in the inner loop;
//   * the same warp then writes its queries' neighbourhood rows [features | xyz - centre | 0]
//     either as plain f32 rows or straight into the bf16 128B-swizzled tile image that
//     tcgen05.mma consumes (common.cuh), so the grouped tensor is written once, coalesced,
//     in the layout its only consumer wants.
#include <cmath>
#include <cstring>
#include <cstdlib>
#include "common.cuh"
#include "group_rows.cuh"

namespace gspn {

constexpr int kBQWarps = 8;
constexpr int kBQThreads = kBQWarps * 32;
constexpr int kBQTile = 2048;  // points per smem tile (24 KiB as packed xyz)

// largest float s with max(sqrtf(s),1e-20f) < radius, or -1 if no s >= 0 qualifies.
static float ball_threshold(float radius) {
    if (!(radius > 1e-20f)) return -1.0f;
    if (std::isinf(radius)) return 3.402823466e38f;
    uint32_t lo = 0, hi = 0x7F7FFFFFu;  // invariant: sqrtf(lo) < radius (sqrtf(0)=0 < radius)
    auto ok = [&](uint32_t bits) { float s; std::memcpy(&s, &bits, 4); return sqrtf(s) < radius; };
    if (ok(hi)) return 3.402823466e38f;
    while (hi - lo > 1) {
        uint32_t mid = lo + (hi - lo) / 2;
        if (ok(mid)) lo = mid; else hi = mid;
    }
    float s; std::memcpy(&s, &lo, 4);
    return s;
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// TMA bulk copy global -> shared (UBLKCP); bytes multiple of 16, both addresses 16B aligned.
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// grid = (ceil(m / (kBQWarps*QPW)), b).  dynamic smem: 2 point tiles + per-warp index rows.
// qpc = queries per CTA (<= kBQWarps*QPW).  With fewer queries than warps the idle warps skip the search and the whole CTA writes
// the neighbourhood rows together: small, wide levels (SA3: 1024 queries x 131 channels, SA4: 256 x 259) otherwise run on a
// handful of warps that each copy their 32 x ld block alone.
template <int QPW>
__global__ void __launch_bounds__(kBQThreads) ballquery_kernel(int n, int m, float s_max, int nsample, const float *__restrict__ xyz1,
                                                               const float *__restrict__ xyz2, int *__restrict__ idx,
                                                               int *__restrict__ pts_cnt, GroupArgs g, int qpc) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *tiles = reinterpret_cast<float *>(smem_raw);                        // [2][kBQTile*3]
    int *widx = reinterpret_cast<int *>(smem_raw + 2 * kBQTile * 3 * 4);       // [kBQWarps][QPW][nsample]
    __shared__ __align__(8) uint64_t full[2];
    __shared__ float sq[kBQWarps * QPW][3];  // query centres, for the CTA-wide row writer

    const int cloud = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *p = xyz1 + (size_t)cloud * n * 3;
    const float *q = xyz2 + (size_t)cloud * m * 3;
    const int jbase = blockIdx.x * qpc + warp * QPW;
    // TMA bulk path needs the cloud base 16B aligned; tile starts are then aligned too (kBQTile*12 % 16 == 0)
    const bool bulk = ((reinterpret_cast<uintptr_t>(p) & 15u) == 0);

    float qx[QPW], qy[QPW], qz[QPW];
    int cnt[QPW];
    bool active = false;
#pragma unroll
    for (int t = 0; t < QPW; ++t) {
        int j = jbase + t;
        bool ok = j < m && warp * QPW + t < qpc;
        int jj = ok ? j : 0;
        qx[t] = __ldg(q + 3 * jj); qy[t] = __ldg(q + 3 * jj + 1); qz[t] = __ldg(q + 3 * jj + 2);
        cnt[t] = ok ? 0 : nsample;  // out-of-range queries are born finished
        active |= ok;
    }
    int *myidx = widx + (size_t)warp * QPW * nsample;

    const int ntiles = ceil_div(n, kBQTile);
    if (bulk) {
        if (threadIdx.x == 0) { mbar_init(&full[0], 1); mbar_init(&full[1], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncthreads();
    }
    auto issue = [&](int t) {  // one thread: fetch tile t into buffer t&1
        int cntp = min(kBQTile, n - t * kBQTile);
        uint32_t bytes = (uint32_t)cntp * 12u;
        uint32_t body = bytes & ~15u;
        if (body) {
            mbar_expect_tx(&full[t & 1], body);
            bulk_g2s(tiles + (size_t)(t & 1) * kBQTile * 3, p + (size_t)t * kBQTile * 3, body, &full[t & 1]);
        } else {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[t & 1])) : "memory");
        }
    };
    if (bulk && threadIdx.x == 0) issue(0);

    int t = 0;
    for (; t < ntiles; ++t) {
        const int k0 = t * kBQTile;
        const int cntp = min(kBQTile, n - k0);
        float *tile = tiles + (size_t)(t & 1) * kBQTile * 3;
        if (bulk) {
            if (threadIdx.x == 0 && t + 1 < ntiles) issue(t + 1);  // buffer (t+1)&1 was released by the barrier closing tile t-1
            mbar_wait(&full[t & 1], (t >> 1) & 1);
            // the (<16 byte) tail the bulk engine cannot move
            uint32_t bytes = (uint32_t)cntp * 12u, body = bytes & ~15u;
            int tail0 = body >> 2, tail1 = bytes >> 2;
            if (tail0 < tail1) {
                if (threadIdx.x < tail1 - tail0) tile[tail0 + threadIdx.x] = __ldg(p + (size_t)k0 * 3 + tail0 + threadIdx.x);
                __syncthreads();
            }
        } else {
            for (int e = threadIdx.x; e < cntp * 3; e += kBQThreads) tile[e] = __ldg(p + (size_t)k0 * 3 + e);
            __syncthreads();
        }
        if (active) {
            for (int base = 0; base < cntp; base += 32) {
                int kl = base + lane;
                bool valid = kl < cntp;
                int ks = valid ? kl : 0;
                float x = tile[3 * ks], y = tile[3 * ks + 1], z = tile[3 * ks + 2];
                bool alldone = true;
#pragma unroll
                for (int u = 0; u < QPW; ++u) {
                    if (cnt[u] < nsample) {  // warp-uniform
                        float s = sqdist_fma(qx[u], qy[u], qz[u], x, y, z);
                        bool hit = valid && !(s > s_max);
                        unsigned bal = __ballot_sync(GSPN_FULL_MASK, hit);
                        if (bal) {
                            int pos = cnt[u] + __popc(bal & ((1u << lane) - 1u));
                            if (hit && pos < nsample) myidx[u * nsample + pos] = k0 + kl;
                            cnt[u] = min(nsample, cnt[u] + __popc(bal));
                        }
                        alldone &= (cnt[u] >= nsample);
                    }
                }
                if (alldone) { active = false; break; }
            }
        }
        // tile buffer reuse + CTA-wide early exit once every warp has filled all its rows
        if (__syncthreads_or(active ? 1 : 0) == 0) {
            break;
        }
    }
    // drain: when the CTA left early, tile t+1 was already requested; it must land before the CTA retires
    if (bulk && t + 1 < ntiles) mbar_wait(&full[(t + 1) & 1], ((t + 1) >> 1) & 1);

    __syncwarp();
    const bool coop = g.grouped && qpc < kBQWarps * QPW;
#pragma unroll
    for (int u = 0; u < QPW; ++u) {
        int j = jbase + u;
        if (j >= m || warp * QPW + u >= qpc) continue;
        int cn = cnt[u];
        int first = cn > 0 ? myidx[u * nsample] : 0;  // zero-hit row: zeros (reference leaves it unwritten)
        __syncwarp();
        for (int l = lane; l < nsample; l += 32) {
            int v = l < cn ? myidx[u * nsample + l] : first;  // back-fill with the first hit (:29-32)
            myidx[u * nsample + l] = v;
            idx[((size_t)cloud * m + j) * nsample + l] = v;
        }
        if (lane == 0) pts_cnt[(size_t)cloud * m + j] = cn;
        __syncwarp();
        if (coop) {
            if (lane == 0) { sq[warp * QPW + u][0] = qx[u]; sq[warp * QPW + u][1] = qy[u]; sq[warp * QPW + u][2] = qz[u]; }
        } else if (g.grouped) {
            write_group(g, n, m, nsample, cloud, j, myidx + u * nsample, xyz1, qx[u], qy[u], qz[u], lane);
        }
    }
    if (coop) {  // CTA-uniform
        __syncthreads();
        for (int uq = 0; uq < qpc; ++uq) {
            const int j = blockIdx.x * qpc + uq;
            if (j >= m) break;
            write_group(g, n, m, nsample, cloud, j, widx + (size_t)uq * nsample, xyz1, sq[uq][0], sq[uq][1], sq[uq][2], threadIdx.x, kBQThreads);
        }
    }
}

// ---- several nested balls around the same queries in ONE ordered scan (multi_encoding_net, models/model_rpointnet.py:49-61: radii
// 0.5 / 1.0 / 1.5 with nsample 256 / 256 / 512 around the same seeds).  A warp owns a query; the squared distance of a point is
// computed once and compared with every ball's threshold; each ball keeps its own "first nsample in index order" list, and the scan
// stops when all of them are full.  Results are bit-identical to one query_ball_point call per radius.
constexpr int kBQMaxRadii = 4;
struct MultiArgs {
    float s_max[kBQMaxRadii];
    int nsample[kBQMaxRadii];
    int off[kBQMaxRadii];  // offset of ball r's index row inside a warp's shared-memory block
    int *idx[kBQMaxRadii];
    int *cnt[kBQMaxRadii];
    int nrad, row_ints;
};

__global__ void __launch_bounds__(kBQThreads) ballquery_multi_kernel(int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                                                                     const MultiArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float *tiles = reinterpret_cast<float *>(smem_raw);                   // [2][kBQTile*3]
    int *widx = reinterpret_cast<int *>(smem_raw + 2 * kBQTile * 3 * 4);  // [kBQWarps][row_ints]
    const int cloud = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float *p = xyz1 + (size_t)cloud * n * 3;
    const int j = blockIdx.x * kBQWarps + warp;
    const bool mine = j < m;
    const float *q = xyz2 + ((size_t)cloud * m + (mine ? j : 0)) * 3;
    const float qx = __ldg(q), qy = __ldg(q + 1), qz = __ldg(q + 2);
    int cnt[kBQMaxRadii];
#pragma unroll
    for (int r = 0; r < kBQMaxRadii; ++r) cnt[r] = (mine && r < a.nrad) ? 0 : 0x7fffffff;  // absent balls are born full
    int *myidx = widx + (size_t)warp * a.row_ints;
    bool active = mine;
    const int ntiles = ceil_div(n, kBQTile);
    for (int t = 0; t < ntiles; ++t) {
        const int k0 = t * kBQTile;
        const int cntp = min(kBQTile, n - k0);
        float *tile = tiles + (size_t)(t & 1) * kBQTile * 3;
        for (int e = threadIdx.x; e < cntp * 3; e += kBQThreads) tile[e] = __ldg(p + (size_t)k0 * 3 + e);
        __syncthreads();
        if (active) {
            for (int base = 0; base < cntp; base += 32) {
                const int kl = base + lane;
                const bool valid = kl < cntp;
                const int ks = valid ? kl : 0;
                const float s = sqdist_fma(qx, qy, qz, tile[3 * ks], tile[3 * ks + 1], tile[3 * ks + 2]);
                bool alldone = true;
#pragma unroll
                for (int r = 0; r < kBQMaxRadii; ++r) {
                    if (cnt[r] < a.nsample[r]) {  // warp-uniform; absent balls never enter
                        const bool hit = valid && !(s > a.s_max[r]);
                        const unsigned bal = __ballot_sync(GSPN_FULL_MASK, hit);
                        if (bal) {
                            const int pos = cnt[r] + __popc(bal & ((1u << lane) - 1u));
                            if (hit && pos < a.nsample[r]) myidx[a.off[r] + pos] = k0 + kl;
                            cnt[r] = min(a.nsample[r], cnt[r] + __popc(bal));
                        }
                        alldone &= (cnt[r] >= a.nsample[r]);
                    }
                }
                if (alldone) { active = false; break; }
            }
        }
        if (__syncthreads_or(active ? 1 : 0) == 0) break;  // tile buffer reuse + CTA-wide early exit
    }
    __syncwarp();
    if (!mine) return;
#pragma unroll
    for (int r = 0; r < kBQMaxRadii; ++r) {
        if (r >= a.nrad) continue;
        const int ns = a.nsample[r], cn = cnt[r];
        const int first = cn > 0 ? myidx[a.off[r]] : 0;  // zero-hit row: zeros (reference leaves it unwritten)
        int *dst = a.idx[r] + ((size_t)cloud * m + j) * ns;
        for (int l = lane; l < ns; l += 32) dst[l] = l < cn ? myidx[a.off[r] + l] : first;  // back-fill with the first hit (:29-32)
        if (lane == 0) a.cnt[r][(size_t)cloud * m + j] = cn;
    }
}

}  // namespace gspn

using namespace gspn;

// grid_search.cu
int gspn_ballquery_grid_launch(int b, int n, int m, float radius, int nsample, const float *xyz1, const float *xyz2, int *idx, int *pts_cnt,
                               GroupArgs ga, void *workspace, cudaStream_t s);
extern "C" size_t gspn_grid_workspace_bytes(int b, int n);
constexpr int kGridMinPoints = 4096;  // below this the brute-force scan is already a few microseconds

static int g_bq_qpw = 0, g_bq_qpc = 0;  // tuning doors: queries per warp / per CTA of the ordered-scan kernel (0 = choose)
extern "C" void gspn_ballquery_tune(int queries_per_warp, int queries_per_cta) { g_bq_qpw = queries_per_warp; g_bq_qpc = queries_per_cta; }

static int launch_ballquery(int b, int n, int m, float radius, int nsample, const float *xyz1, const float *xyz2, int *idx, int *pts_cnt,
                            GroupArgs g, cudaStream_t s) {
    const float s_max = ball_threshold(radius);
    // QPW: more queries per warp amortise the shared-memory reads and give independent FMA chains,
    // as long as one 8-warp CTA per SM remains
    long warps1 = (long)b * m;
    int qpw = warps1 >= 148L * 8 * 4 ? 2 : 1;  // measured (tools/op_bench.py): 2 beats 1 and 4 on 8 x 32768 -> 2048
    if (g_bq_qpw == 1 || g_bq_qpw == 2 || g_bq_qpw == 4) qpw = g_bq_qpw;  // tuning door (gspn_ballquery_tune)
    if ((size_t)qpw * nsample * 4 * kBQWarps > 96 * 1024) qpw = 1;
    size_t smem = (size_t)2 * kBQTile * 3 * 4 + (size_t)kBQWarps * qpw * nsample * 4;
    if (smem > 200 * 1024) return GSPN_E_UNSUPPORTED;
    // queries per CTA: all warps search unless that leaves the GPU with only a few CTAs of row-copying warps; then fewer
    // queries per CTA (1, 2 or 4) and the CTA writes their rows together (door: GSPN_BQ_QPC)
    // measured on B200 (bench stage times): SA3 (n=512, 1024 queries, ld 192) 30.6 -> 19.3 us, SA4 (n=128, 256 queries, ld 320)
    // 38.3 -> 12.6 us at one query per CTA; SA2 (n=2048: the scan itself is long) gets slower (39 -> 44 us) and keeps 8
    int qpc = kBQWarps * qpw;
    if (g.grouped && g.ld >= 64 && n <= 1024) {
        while (qpc > 1 && (long)b * ceil_div(m, qpc) < 148L * 4) qpc >>= 1;
    }
    if (g_bq_qpc >= 1 && g_bq_qpc <= kBQWarps * qpw) qpc = g_bq_qpc;
    dim3 grid(ceil_div(m, qpc), b);
#define GSPN_BQ_LAUNCH(Q)                                                                                                       \
    do {                                                                                                                        \
        GSPN_CUDA_OK(cudaFuncSetAttribute(ballquery_kernel<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
        ballquery_kernel<Q><<<grid, kBQThreads, smem, s>>>(n, m, s_max, nsample, xyz1, xyz2, idx, pts_cnt, g, qpc);              \
    } while (0)
    if (qpw == 4) GSPN_BQ_LAUNCH(4);
    else if (qpw == 2) GSPN_BQ_LAUNCH(2);
    else GSPN_BQ_LAUNCH(1);
#undef GSPN_BQ_LAUNCH
    return check_launch();
}

extern "C" int gspn_query_ball_point(int b, int n, int m, float radius, int nsample, const float *xyz1, const float *xyz2, int *idx,
                                     int *pts_cnt, void *workspace, size_t workspace_bytes, gspn_stream_t stream) {
    GSPN_REQUIRE(radius > 0.f && nsample > 0);  // tf_grouping.cpp:101,104
    GSPN_REQUIRE(b >= 0 && n > 0 && m >= 0 && b <= 65535);  // :109-114
    if (b == 0 || m == 0) return GSPN_OK;
    GSPN_REQUIRE_PTR(xyz1); GSPN_REQUIRE_PTR(xyz2); GSPN_REQUIRE_PTR(idx); GSPN_REQUIRE_PTR(pts_cnt);
    GroupArgs g = {};
    if (workspace != nullptr && n >= kGridMinPoints && std::isfinite(radius)) {
        if (workspace_bytes < gspn_grid_workspace_bytes(b, n)) return GSPN_E_WORKSPACE;
        return gspn_ballquery_grid_launch(b, n, m, radius, nsample, xyz1, xyz2, idx, pts_cnt, g, workspace, as_stream(stream));
    }
    return launch_ballquery(b, n, m, radius, nsample, xyz1, xyz2, idx, pts_cnt, g, as_stream(stream));
}

extern "C" int gspn_query_ball_point_multi(int b, int n, int m, int nrad, const float *radii, const int *nsamples, const float *xyz1,
                                           const float *xyz2, int *const *idx, int *const *pts_cnt, gspn_stream_t stream) {
    GSPN_REQUIRE(b >= 0 && n > 0 && m >= 0 && b <= 65535 && nrad >= 1 && nrad <= kBQMaxRadii);
    GSPN_REQUIRE_PTR(radii); GSPN_REQUIRE_PTR(nsamples); GSPN_REQUIRE_PTR(idx); GSPN_REQUIRE_PTR(pts_cnt);
    if (b == 0 || m == 0) return GSPN_OK;
    GSPN_REQUIRE_PTR(xyz1); GSPN_REQUIRE_PTR(xyz2);
    MultiArgs a = {};
    a.nrad = nrad;
    int off = 0;
    for (int r = 0; r < kBQMaxRadii; ++r) {
        if (r < nrad) {
            GSPN_REQUIRE(radii[r] > 0.f && nsamples[r] > 0);  // tf_grouping.cpp:101,104
            GSPN_REQUIRE_PTR(idx[r]); GSPN_REQUIRE_PTR(pts_cnt[r]);
            a.s_max[r] = ball_threshold(radii[r]);
            a.nsample[r] = nsamples[r];
            a.off[r] = off;
            a.idx[r] = idx[r]; a.cnt[r] = pts_cnt[r];
            off += nsamples[r];
        } else {
            a.s_max[r] = -1.f; a.nsample[r] = 0; a.off[r] = 0;
        }
    }
    a.row_ints = off;
    const size_t smem = (size_t)2 * kBQTile * 3 * 4 + (size_t)kBQWarps * off * 4;
    if (smem > 200 * 1024) return GSPN_E_UNSUPPORTED;
    GSPN_CUDA_OK(cudaFuncSetAttribute(ballquery_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(m, kBQWarps), b);
    ballquery_multi_kernel<<<grid, kBQThreads, smem, as_stream(stream)>>>(n, m, xyz1, xyz2, a);
    return check_launch();
}

extern "C" size_t gspn_grouped_bytes(long rows, int c_plus_xyz, int grouped_dtype) {
    if (rows <= 0 || c_plus_xyz <= 0) return 0;
    if (grouped_dtype == GSPN_DT_BF16 || grouped_dtype == GSPN_DT_BF16X2) {
        long tiles = ceil_div_l(rows, kTileRows);
        int ld = ceil_div(c_plus_xyz, 64) * 64;
        return (size_t)tiles * (size_t)(ld / 64) * (size_t)kTileBytes * (grouped_dtype == GSPN_DT_BF16X2 ? 2 : 1);
    }
    return (size_t)rows * (size_t)c_plus_xyz * sizeof(float);
}

extern "C" int gspn_ballquery_group(int b, int n, int m, int c, float radius, int nsample, const float *xyz, const float *new_xyz,
                                    const float *shift, const void *points, int points_dtype, int *idx, int *pts_cnt, void *grouped,
                                    int grouped_dtype, int ld, void *workspace, size_t workspace_bytes, gspn_stream_t stream) {
    GSPN_REQUIRE(radius > 0.f && nsample > 0);
    GSPN_REQUIRE(b >= 0 && n > 0 && m >= 0 && c >= 0 && b <= 65535);
    if (points_dtype != GSPN_DT_F32 && points_dtype != GSPN_DT_BF16) return GSPN_E_BAD_DTYPE;
    if (grouped_dtype != GSPN_DT_F32 && grouped_dtype != GSPN_DT_BF16 && grouped_dtype != GSPN_DT_BF16X2) return GSPN_E_BAD_DTYPE;
    if (b == 0 || m == 0) return GSPN_OK;
    GSPN_REQUIRE_PTR(xyz); GSPN_REQUIRE_PTR(new_xyz); GSPN_REQUIRE_PTR(idx); GSPN_REQUIRE_PTR(pts_cnt); GSPN_REQUIRE_PTR(grouped);
    if (c > 0) GSPN_REQUIRE_PTR(points);
    GSPN_REQUIRE(ld >= c + 3);
    if (grouped_dtype != GSPN_DT_F32) GSPN_REQUIRE(ld % 64 == 0);
    GroupArgs g;
    g.shift = shift;
    g.points = points;
    g.c = c;
    g.points_bf16 = points_dtype == GSPN_DT_BF16;
    g.grouped = grouped;
    g.grouped_bf16 = grouped_dtype == GSPN_DT_BF16 ? 1 : (grouped_dtype == GSPN_DT_BF16X2 ? 2 : 0);
    g.ld = ld;
    if (workspace != nullptr && n >= kGridMinPoints && std::isfinite(radius)) {
        if (workspace_bytes < gspn_grid_workspace_bytes(b, n)) return GSPN_E_WORKSPACE;
        return gspn_ballquery_grid_launch(b, n, m, radius, nsample, xyz, new_xyz, idx, pts_cnt, g, workspace, as_stream(stream));
    }
    return launch_ballquery(b, n, m, radius, nsample, xyz, new_xyz, idx, pts_cnt, g, as_stream(stream));
}
