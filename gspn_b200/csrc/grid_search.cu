// grid_search.cu -- exact neighbour search through a per-cloud uniform grid.
//
// The reference's ball query and three_nn are brute-force O(n*m) scans (tf_grouping_g.cu:6-39,
// tf_interpolate.cpp:60-103); at 8 x 32768 points that is 537 M pair tests each and the scan, not memory, is
// what the op costs.  Here the scanned cloud is bucketed once per call into a uniform grid (counting sort:
// bbox -> count -> scan -> scatter, points stored as float4 (x,y,z,index) in cell order), and a query only
// looks at the 3x3x3 block of cells around it (three_nn: a growing block).  Results are IDENTICAL to the brute
// force scans:
//   * ball query: the hit set is decided by the same float predicate on the same operands; the reference's
//     "first nsample in index order" is restored by extracting the nsample smallest indices of the hit set
//     (repeated warp min); a query whose block holds more hits than the hit buffer falls back to the ordered
//     scan, which then terminates early because hits are dense;
//   * three_nn: candidates are inserted with the lexicographic (distance, index) order the ascending strict-<
//     scan produces, and a block of half-width R cells is accepted only if the third best distance is safely
//     inside R*h, so no unvisited point can tie or beat it.
// Cell size h >= 1.001 * radius guarantees (with float rounding of the cell coordinate) that every point
// passing the predicate lies in the 3x3x3 block.
#include <cmath>
#include <cstring>
#include "common.cuh"
#include "group_rows.cuh"

namespace gspn {

constexpr int kGridMinCap = 32768;
// cell budget per cloud: about two cells per scanned point, between 32 Ki and 1 Mi
// (a multiple of 4: every cloud's cell arrays then start 16-byte aligned for the vectorised scan)
static inline int grid_cap(int n) { long c = (2L * n + 3) & ~3L; c = c < kGridMinCap ? kGridMinCap : c; c = c > (1L << 20) ? (1L << 20) : c; return (int)c; }
constexpr int kSortQueriesMin = 65536;  // query sets at least this large are visited in cell order (see sort_queries)
constexpr int kHitCap = 512;  // per-warp hit buffer (indices) of the grid ball query
constexpr int kGQWarps = 8;

struct __align__(16) GridHeader {
    float ox, oy, oz, inv_h, h;
    int gx, gy, gz, ncells;
    int pad[3];
};

// workspace: [GridHeader x b][sorted float4 n x b][cell_start (cap+4) x b][cursor (cap+4) x b], cap = grid_cap(n)
struct GridWs {
    GridHeader *hdr;
    int *cell_start;
    int *cursor;
    float4 *sorted;
};
static size_t grid_ws_bytes(int b, int n) {
    return (size_t)b * (sizeof(GridHeader) + sizeof(int) * 2 * ((size_t)grid_cap(n) + 4) + sizeof(float4) * (size_t)n) + 256;
}
static GridWs carve(void *ws, int b, int n) {
    GridWs g;
    unsigned char *p = (unsigned char *)(((uintptr_t)ws + 127) & ~(uintptr_t)127);
    g.hdr = (GridHeader *)p; p += (size_t)b * sizeof(GridHeader);
    g.sorted = (float4 *)p; p += (size_t)b * n * sizeof(float4);
    g.cell_start = (int *)p; p += (size_t)b * (grid_cap(n) + 4) * sizeof(int);
    g.cursor = (int *)p;
    return g;
}

__device__ __forceinline__ int cell_coord(float v, float o, float inv_h, int g) {
    float u = __fmul_rn(__fsub_rn(v, o), inv_h);
    u = fminf(fmaxf(u, -2.f), (float)g + 1.f);  // far-away queries stay representable; NaN -> -2
    return (int)floorf(u);
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ---- 1. bounding box + grid geometry, one CTA per cloud
__global__ void __launch_bounds__(1024) grid_bbox_kernel(int n, const float *__restrict__ xyz, float h_min, float target_cells, int cap, GridHeader *hdr) {
    __shared__ float red[6][32];
    const int cloud = blockIdx.x;
    const float *p = xyz + (size_t)cloud * n * 3;
    float lo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, hi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
    if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(p) & 15u) == 0 && blockDim.x == 1024) {
        // the cloud as a flat array of 3n floats, read as coalesced float4: component c of vector v belongs to axis
        // (4v + c) mod 3 = (v + c) mod 3; with v = tid + 1024*i and 1024 = 1 (mod 3) that is (tid + i + c) mod 3.  Accumulate by
        // q = (i + c) mod 3 (static after unrolling i by 3) and rotate by tid mod 3 at the end.
        const float4 *p4 = reinterpret_cast<const float4 *>(p);
        const int nvec = (3 * n) >> 2;
        float ql[3] = {3.4e38f, 3.4e38f, 3.4e38f}, qh[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
        for (int v = threadIdx.x; v < nvec; v += 3 * 1024) {
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int vv = v + u * 1024;
                if (vv < nvec) {
                    const float4 f = __ldg(p4 + vv);
                    ql[u % 3] = fminf(ql[u % 3], f.x); qh[u % 3] = fmaxf(qh[u % 3], f.x);
                    ql[(u + 1) % 3] = fminf(ql[(u + 1) % 3], f.y); qh[(u + 1) % 3] = fmaxf(qh[(u + 1) % 3], f.y);
                    ql[(u + 2) % 3] = fminf(ql[(u + 2) % 3], f.z); qh[(u + 2) % 3] = fmaxf(qh[(u + 2) % 3], f.z);
                    ql[u % 3] = fminf(ql[u % 3], f.w); qh[u % 3] = fmaxf(qh[u % 3], f.w);  // c = 3: (u + 3) mod 3 = u
                }
            }
        }
        const int r = threadIdx.x % 3;  // axis a holds class q = (a - r) mod 3
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            lo[a] = r == 0 ? ql[a] : (r == 1 ? ql[(a + 2) % 3] : ql[(a + 1) % 3]);
            hi[a] = r == 0 ? qh[a] : (r == 1 ? qh[(a + 2) % 3] : qh[(a + 1) % 3]);
        }
    } else {
        for (int k = threadIdx.x; k < n; k += blockDim.x) {
#pragma unroll
            for (int a = 0; a < 3; ++a) {
                float v = __ldg(p + 3 * k + a);
                lo[a] = fminf(lo[a], v);
                hi[a] = fmaxf(hi[a], v);
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(GSPN_FULL_MASK, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(GSPN_FULL_MASK, hi[a], o));
        }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
        for (int a = 0; a < 3; ++a) { red[a][warp] = lo[a]; red[3 + a][warp] = hi[a]; }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int nw = blockDim.x >> 5;
        for (int a = 0; a < 3; ++a)
            for (int w = 1; w < nw; ++w) { red[a][0] = fminf(red[a][0], red[a][w]); red[3 + a][0] = fmaxf(red[3 + a][0], red[3 + a][w]); }
        float ex = fmaxf(red[3][0] - red[0][0], 0.f), ey = fmaxf(red[4][0] - red[1][0], 0.f), ez = fmaxf(red[5][0] - red[2][0], 0.f);
        float emax = fmaxf(fmaxf(ex, ey), fmaxf(ez, 1e-30f));
        // cell size: at least h_min, about target_cells cells over the bounding volume, never below emax/1024
        float vol = fmaxf(ex, emax * 1e-3f) * fmaxf(ey, emax * 1e-3f) * fmaxf(ez, emax * 1e-3f);
        float h = fmaxf(fmaxf(h_min, cbrtf(vol / target_cells)), emax / 1024.f);
        int gx, gy, gz;
        for (;;) {
            gx = (int)(ex / h) + 1; gy = (int)(ey / h) + 1; gz = (int)(ez / h) + 1;
            if ((long)gx * gy * gz <= cap) break;
            h *= 1.26f;
        }
        GridHeader g;
        g.ox = red[0][0]; g.oy = red[1][0]; g.oz = red[2][0];
        g.h = h; g.inv_h = 1.0f / h;
        g.gx = gx; g.gy = gy; g.gz = gz; g.ncells = gx * gy * gz;
        g.pad[0] = g.pad[1] = g.pad[2] = 0;
        hdr[cloud] = g;
    }
}

__device__ __forceinline__ int point_cell(const GridHeader &g, float x, float y, float z) {
    int cx = clampi(cell_coord(x, g.ox, g.inv_h, g.gx), 0, g.gx - 1);
    int cy = clampi(cell_coord(y, g.oy, g.inv_h, g.gy), 0, g.gy - 1);
    int cz = clampi(cell_coord(z, g.oz, g.inv_h, g.gz), 0, g.gz - 1);
    return (cz * g.gy + cy) * g.gx + cx;
}

// ---- 2. histogram of points per cell (counts pre-zeroed)
__global__ void __launch_bounds__(256) grid_count_kernel(int n, int cap, const float *__restrict__ xyz, const GridHeader *__restrict__ hdr, int *counts) {
    const int cloud = blockIdx.y;
    const GridHeader g = hdr[cloud];
    const float *p = xyz + (size_t)cloud * n * 3;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x)
        atomicAdd(counts + (size_t)cloud * (cap + 4) + point_cell(g, __ldg(p + 3 * k), __ldg(p + 3 * k + 1), __ldg(p + 3 * k + 2)), 1);
}

// ---- 3. exclusive scan of the counts in place (cell_start) + copy to the scatter cursors, one CTA per cloud.
// 16 cells per thread per pass as four 16-byte vectors (cells beyond ncells hold zero counts and the arrays are cap+4 long),
// block-wide scan of the per-thread sums, running carry across passes.
__global__ void __launch_bounds__(1024) grid_scan_kernel(int cap, const GridHeader *__restrict__ hdr, int *cell_start, int *cursor) {
    __shared__ int wsum[33];
    const int cloud = blockIdx.x;
    const int ncells = hdr[cloud].ncells;
    int *cs = cell_start + (size_t)cloud * (cap + 4);
    int *cu = cursor + (size_t)cloud * (cap + 4);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int carry = 0;
    for (int base = 0; base < ncells; base += 16 * 1024) {
        const int c = base + 16 * threadIdx.x;
        int4 v[4];
        int sum = 0;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            v[g] = (c + 4 * g < ncells) ? *reinterpret_cast<const int4 *>(cs + c + 4 * g) : make_int4(0, 0, 0, 0);
            sum += v[g].x + v[g].y + v[g].z + v[g].w;
        }
        int incl = sum;
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(GSPN_FULL_MASK, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = wsum[lane];
            int inc2 = w;
            for (int o = 1; o < 32; o <<= 1) {
                int t = __shfl_up_sync(GSPN_FULL_MASK, inc2, o);
                if (lane >= o) inc2 += t;
            }
            wsum[lane] = inc2 - w;
            if (lane == 31) wsum[32] = inc2;
        }
        __syncthreads();
        int run = carry + wsum[warp] + incl - sum;
        carry += wsum[32];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            if (c + 4 * g < ncells) {
                int4 o4;
                o4.x = run; run += v[g].x;
                o4.y = run; run += v[g].y;
                o4.z = run; run += v[g].z;
                o4.w = run; run += v[g].w;
                *reinterpret_cast<int4 *>(cs + c + 4 * g) = o4;
                *reinterpret_cast<int4 *>(cu + c + 4 * g) = o4;
            }
        }
        __syncthreads();  // wsum is reused by the next pass
    }
    if (threadIdx.x == 0) cs[ncells] = carry;  // end marker (total number of points)
}

// ---- 4. scatter points into cell order as (x,y,z,index)
__global__ void __launch_bounds__(256) grid_scatter_kernel(int n, int cap, const float *__restrict__ xyz, const GridHeader *__restrict__ hdr, int *cursor,
                                                           float4 *sorted) {
    const int cloud = blockIdx.y;
    const GridHeader g = hdr[cloud];
    const float *p = xyz + (size_t)cloud * n * 3;
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        float x = __ldg(p + 3 * k), y = __ldg(p + 3 * k + 1), z = __ldg(p + 3 * k + 2);
        int pos = atomicAdd(cursor + (size_t)cloud * (cap + 4) + point_cell(g, x, y, z), 1);
        sorted[(size_t)cloud * n + pos] = make_float4(x, y, z, __int_as_float(k));
    }
}

static int build_grid(int b, int n, const float *xyz, float h_min, float target_cells, const GridWs &ws, cudaStream_t s) {
    const int cap = grid_cap(n);
    GSPN_CUDA_OK(cudaMemsetAsync(ws.cell_start, 0, sizeof(int) * (size_t)b * (cap + 4), s));
    grid_bbox_kernel<<<b, 1024, 0, s>>>(n, xyz, h_min, target_cells < (float)cap ? target_cells : (float)cap, cap, ws.hdr);
    dim3 grid(ceil_div(n, 256) < 64 ? ceil_div(n, 256) : 64, b);
    grid_count_kernel<<<grid, 256, 0, s>>>(n, cap, xyz, ws.hdr, ws.cell_start);
    grid_scan_kernel<<<b, 1024, 0, s>>>(cap, ws.hdr, ws.cell_start, ws.cursor);
    grid_scatter_kernel<<<grid, 256, 0, s>>>(n, cap, xyz, ws.hdr, ws.cursor, ws.sorted);
    return check_launch();
}

// ---- ball query through the grid: one warp per query
__global__ void __launch_bounds__(kGQWarps * 32) ballquery_grid_kernel(int n, int m, int cap, float s_max, int nsample, const float *__restrict__ xyz1,
                                                                       const float *__restrict__ xyz2, const GridHeader *__restrict__ hdr,
                                                                       const int *__restrict__ cell_start, const float4 *__restrict__ sorted,
                                                                       int *__restrict__ idx, int *__restrict__ pts_cnt, GroupArgs ga) {
    extern __shared__ int smem_i[];  // [kGQWarps][kHitCap] hit buffer, then [kGQWarps][nsample] output rows
    const int cloud = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = blockIdx.x * kGQWarps + warp;
    if (j >= m) return;
    int *hits = smem_i + warp * kHitCap;
    int *row = smem_i + kGQWarps * kHitCap + warp * nsample;
    const GridHeader g = hdr[cloud];
    const int *cs = cell_start + (size_t)cloud * (cap + 4);
    const float4 *sp = sorted + (size_t)cloud * n;
    const float *q = xyz2 + ((size_t)cloud * m + j) * 3;
    const float qx = __ldg(q), qy = __ldg(q + 1), qz = __ldg(q + 2);
    const int cx = cell_coord(qx, g.ox, g.inv_h, g.gx), cy = cell_coord(qy, g.oy, g.inv_h, g.gy), cz = cell_coord(qz, g.oz, g.inv_h, g.gz);
    const int x0 = max(cx - 1, 0), x1 = min(cx + 1, g.gx - 1);
    int H = 0;
    {
        // The 3x3x3 block is nine contiguous ranges of the cell-sorted array (three x-adjacent cells each).  Lanes 0..8 fetch
        // one range's bounds each -- one parallel round trip instead of nine dependent ones -- and the ranges are then walked
        // as ONE concatenated candidate list, 32 candidates per step, instead of at least one (mostly empty) step per range.
        int rbeg = 0, rlen = 0;
        if (lane < 9 && x0 <= x1) {
            const int zc = cz + lane / 3 - 1, yc = cy + lane % 3 - 1;
            if (zc >= 0 && zc < g.gz && yc >= 0 && yc < g.gy) {
                const int c0 = (zc * g.gy + yc) * g.gx + x0;
                rbeg = __ldg(cs + c0);
                rlen = __ldg(cs + c0 + (x1 - x0) + 1) - rbeg;
            }
        }
        int incl = rlen;
#pragma unroll
        for (int o = 1; o < 16; o <<= 1) {
            const int t = __shfl_up_sync(GSPN_FULL_MASK, incl, o);
            if (lane >= o) incl += t;
        }
        const int excl = incl - rlen;
        const int T = __shfl_sync(GSPN_FULL_MASK, incl, 8);  // candidates in the block
        int pre[9], bg[9];
#pragma unroll
        for (int r = 0; r < 9; ++r) {
            pre[r] = __shfl_sync(GSPN_FULL_MASK, excl, r);
            bg[r] = __shfl_sync(GSPN_FULL_MASK, rbeg, r);
        }
        for (int t0 = 0; t0 < T; t0 += 32) {
            const int t = t0 + lane;
            bool hit = false;
            int k = 0;
            if (t < T) {
                int base = bg[0], p0 = pre[0];  // the last range whose start offset is <= t (empty ranges share their successor's)
#pragma unroll
                for (int r = 1; r < 9; ++r)
                    if (t >= pre[r]) { base = bg[r]; p0 = pre[r]; }
                const float4 p = __ldg(sp + base + (t - p0));
                k = __float_as_int(p.w);
                hit = !(sqdist_fma(qx, qy, qz, p.x, p.y, p.z) > s_max);  // same operands, same rounding as the ordered scan
            }
            const unsigned bal = __ballot_sync(GSPN_FULL_MASK, hit);
            if (bal) {
                const int pos = H + __popc(bal & ((1u << lane) - 1u));
                if (hit && pos < kHitCap) hits[pos] = k;
                H += __popc(bal);
            }
        }
    }
    __syncwarp();
    int cn;
    if (H <= kHitCap) {
        // the nsample smallest indices of the hit set, ascending == first nsample hits of the index-order scan
        cn = min(H, nsample);
        int last = -1;
        for (int r = 0; r < cn; ++r) {
            int local = 0x7FFFFFFF;
            for (int e = lane; e < H; e += 32) {
                int v = hits[e];
                if (v > last && v < local) local = v;
            }
            last = __reduce_min_sync(GSPN_FULL_MASK, local);
            if (lane == 0) row[r] = last;
        }
    } else {
        // dense neighbourhood: the ordered scan of the reference, which now ends after a few hundred points
        const float *p = xyz1 + (size_t)cloud * n * 3;
        cn = 0;
        for (int base = 0; base < n && cn < nsample; base += 32) {
            const int k = base + lane;
            bool hit = false;
            if (k < n) hit = !(sqdist_fma(qx, qy, qz, __ldg(p + 3 * k), __ldg(p + 3 * k + 1), __ldg(p + 3 * k + 2)) > s_max);
            const unsigned bal = __ballot_sync(GSPN_FULL_MASK, hit);
            if (bal) {
                const int pos = cn + __popc(bal & ((1u << lane) - 1u));
                if (hit && pos < nsample) row[pos] = k;
                cn = min(nsample, cn + __popc(bal));
            }
        }
    }
    __syncwarp();
    const int first = cn > 0 ? row[0] : 0;  // zero-hit row: zeros (the reference leaves it unwritten)
    __syncwarp();
    for (int l = lane; l < nsample; l += 32) {
        const int v = l < cn ? row[l] : first;  // back-fill with the first hit (tf_grouping_g.cu:29-32)
        row[l] = v;
        idx[((size_t)cloud * m + j) * nsample + l] = v;
    }
    if (lane == 0) pts_cnt[(size_t)cloud * m + j] = cn;
    __syncwarp();
    if (ga.grouped) write_group(ga, n, m, nsample, cloud, j, row, xyz1, qx, qy, qz, lane);
}

// ---- three_nn through the grid: one thread per unknown point, growing cell block
__device__ __forceinline__ void insert3(float d, int k, float &b1, float &b2, float &b3, int &i1, int &i2, int &i3) {
    // order: (distance, index) ascending -- what the ascending strict-'<' scan of threenn_cpu yields
    if (d < b3 || (d == b3 && k < i3)) {
        if (d < b1 || (d == b1 && k < i1)) { b3 = b2; i3 = i2; b2 = b1; i2 = i1; b1 = d; i1 = k; }
        else if (d < b2 || (d == b2 && k < i2)) { b3 = b2; i3 = i2; b2 = d; i2 = k; }
        else { b3 = d; i3 = k; }
    }
}

// TOP1: nn_distance's one-directional 1-NN (tf_nndistance.cpp:21-43) -- the same search keeping only the best; FMA selects
// the rounding of the compiled NmDistanceKernel instead of the CPU loop's.
template <bool TOP1, bool FMA>
__global__ void __launch_bounds__(128) three_nn_grid_kernel(int n, int m, int cap, const float *__restrict__ xyz1, const GridHeader *__restrict__ hdr,
                                                            const int *__restrict__ cell_start, const float4 *__restrict__ sorted,
                                                            float *__restrict__ dist, int *__restrict__ idx, float *__restrict__ weight,
                                                            const float4 *__restrict__ qsorted) {
    const int cloud = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const GridHeader g = hdr[cloud];
    const int *cs = cell_start + (size_t)cloud * (cap + 4);
    const float4 *sp = sorted + (size_t)cloud * m;
    // queries in input order, or (qsorted) in the cell order of a grid over the queries themselves: neighbouring lanes then
    // hold neighbouring points, walk the same cells of the known set and leave the candidate loops together
    int j = t;
    float x1, y1, z1;
    if (qsorted) {
        const float4 q = __ldg(qsorted + (size_t)cloud * n + t);
        x1 = q.x; y1 = q.y; z1 = q.z;
        j = __float_as_int(q.w);
    } else {
        const float *u = xyz1 + ((size_t)cloud * n + t) * 3;
        x1 = __ldg(u); y1 = __ldg(u + 1); z1 = __ldg(u + 2);
    }
    const int cx = clampi(cell_coord(x1, g.ox, g.inv_h, g.gx), -1, g.gx), cy = clampi(cell_coord(y1, g.oy, g.inv_h, g.gy), -1, g.gy),
              cz = clampi(cell_coord(z1, g.oz, g.inv_h, g.gz), -1, g.gz);
    // a query outside the known points' bounding box is still handled: its distance to the block faces only grows
    const int rmax = max(max(g.gx, g.gy), g.gz) + 1;
    float b1, b2, b3;
    int i1, i2, i3;
    for (int R = 1;; R = (R < 4 ? R + 1 : R * 2)) {
        b1 = b2 = b3 = __int_as_float(0x7f800000);  // +inf: 1e40 as float (tf_interpolate.cpp:66,91)
        i1 = i2 = i3 = 0;
        const int xa = max(cx - R, 0), xb = min(cx + R, g.gx - 1);
        if (R == 1) {
            // first pass (almost always the only one): the bounds of all nine ranges are requested together, so the thread
            // pays one memory round trip for them instead of nine dependent ones
            int bb[9], ee[9];
#pragma unroll
            for (int r = 0; r < 9; ++r) {
                const int zc = cz + r / 3 - 1, yc = cy + r % 3 - 1;
                const bool ok = xa <= xb && zc >= 0 && zc < g.gz && yc >= 0 && yc < g.gy;
                const int c0 = ok ? (zc * g.gy + yc) * g.gx + xa : 0;
                bb[r] = ok ? __ldg(cs + c0) : 0;
                ee[r] = ok ? __ldg(cs + c0 + (xb - xa) + 1) : 0;
            }
#pragma unroll
            for (int r = 0; r < 9; ++r) {
                for (int i = bb[r]; i < ee[r]; ++i) {
                    const float4 p = __ldg(sp + i);
                    const float d = FMA ? sqdist_fma(p.x, p.y, p.z, x1, y1, z1) : sqdist_nofma(p.x, p.y, p.z, x1, y1, z1);
                    const int k = __float_as_int(p.w);
                    if (TOP1) {
                        if (d < b1 || (d == b1 && k < i1)) { b1 = d; i1 = k; }
                    } else {
                        insert3(d, k, b1, b2, b3, i1, i2, i3);
                    }
                }
            }
        } else if (xa <= xb) {
            for (int zc = max(cz - R, 0); zc <= min(cz + R, g.gz - 1); ++zc)
                for (int yc = max(cy - R, 0); yc <= min(cy + R, g.gy - 1); ++yc) {
                    const int c0 = (zc * g.gy + yc) * g.gx + xa;
                    const int beg = __ldg(cs + c0), end = __ldg(cs + c0 + (xb - xa) + 1);
                    for (int i = beg; i < end; ++i) {
                        const float4 p = __ldg(sp + i);
                        const float d = FMA ? sqdist_fma(p.x, p.y, p.z, x1, y1, z1) : sqdist_nofma(p.x, p.y, p.z, x1, y1, z1);
                        const int k = __float_as_int(p.w);
                        if (TOP1) {
                            if (d < b1 || (d == b1 && k < i1)) { b1 = d; i1 = k; }
                        } else {
                            insert3(d, k, b1, b2, b3, i1, i2, i3);
                        }
                    }
                }
        }
        if (R >= rmax) break;  // the block covers the whole grid
        // every unvisited point is farther than R*h from the query along some axis; accept only with a safety margin
        // far above float rounding (1e-3 relative on the radius), so no unvisited point can tie or beat b3
        const float cover = (float)R * g.h * 0.999f;
        if ((TOP1 ? b1 : b3) < cover * cover) break;
    }
    if (TOP1) {
        dist[(size_t)cloud * n + j] = b1;
        idx[(size_t)cloud * n + j] = i1;
        return;
    }
    const size_t o = ((size_t)cloud * n + j) * 3;
    dist[o] = b1; dist[o + 1] = b2; dist[o + 2] = b3;
    idx[o] = i1; idx[o + 1] = i2; idx[o + 2] = i3;
    if (weight) {  // pointnet_util.py:157-160
        float r1 = __fdiv_rn(1.0f, fmaxf(b1, 1e-10f)), r2 = __fdiv_rn(1.0f, fmaxf(b2, 1e-10f)), r3 = __fdiv_rn(1.0f, fmaxf(b3, 1e-10f));
        float norm = __fadd_rn(__fadd_rn(r1, r2), r3);
        weight[o] = __fdiv_rn(r1, norm); weight[o + 1] = __fdiv_rn(r2, norm); weight[o + 2] = __fdiv_rn(r3, norm);
    }
}

}  // namespace gspn

using namespace gspn;

// host-side copy of the threshold used by the brute-force kernel (ballquery_group.cu)
static float grid_ball_threshold(float radius) {
    if (!(radius > 1e-20f)) return -1.0f;
    if (std::isinf(radius)) return 3.402823466e38f;
    uint32_t lo = 0, hi = 0x7F7FFFFFu;
    auto ok = [&](uint32_t bits) { float s; std::memcpy(&s, &bits, 4); return sqrtf(s) < radius; };
    if (ok(hi)) return 3.402823466e38f;
    while (hi - lo > 1) {
        uint32_t mid = lo + (hi - lo) / 2;
        if (ok(mid)) lo = mid; else hi = mid;
    }
    float s; std::memcpy(&s, &lo, 4);
    return s;
}

extern "C" size_t gspn_grid_workspace_bytes(int b, int n) { return (b <= 0 || n <= 0) ? 0 : grid_ws_bytes(b, n); }
// scanned set + (large query sets) a second grid that orders the queries
extern "C" size_t gspn_grid_query_workspace_bytes(int b, int n_queries, int n_scanned) {
    if (b <= 0 || n_queries <= 0 || n_scanned <= 0) return 0;
    const size_t first = (grid_ws_bytes(b, n_scanned) + 255) & ~(size_t)255;
    return n_queries >= kSortQueriesMin ? first + grid_ws_bytes(b, n_queries) : grid_ws_bytes(b, n_scanned);
}

// called by gspn_query_ball_point / gspn_ballquery_group when a workspace is supplied (ballquery_group.cu)
int gspn_ballquery_grid_launch(int b, int n, int m, float radius, int nsample, const float *xyz1, const float *xyz2, int *idx, int *pts_cnt,
                               GroupArgs ga, void *workspace, cudaStream_t s) {
    if (!std::isfinite(radius)) return GSPN_E_UNSUPPORTED;
    GridWs ws = carve(workspace, b, n);
    int rc = build_grid(b, n, xyz1, radius * 1.001f, 16384.f, ws, s);
    if (rc != GSPN_OK) return rc;
    const size_t smem = sizeof(int) * (size_t)kGQWarps * (kHitCap + nsample);
    if (smem > 200 * 1024) return GSPN_E_UNSUPPORTED;
    if (smem > 48 * 1024) GSPN_CUDA_OK(cudaFuncSetAttribute(ballquery_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(m, kGQWarps), b);
    ballquery_grid_kernel<<<grid, kGQWarps * 32, smem, s>>>(n, m, grid_cap(n), grid_ball_threshold(radius), nsample, xyz1, xyz2, ws.hdr, ws.cell_start,
                                                           ws.sorted, idx, pts_cnt, ga);
    return check_launch();
}

// workspace_bytes >= grid_ws_bytes(b, m) + grid_ws_bytes(b, n) and a large query set: the queries are bucketed too (their own
// grid, about four per cell) and visited in cell order.  Measured on B200: the search kernel gets 18-26 % faster, the ordering
// costs ~37 us at 8 x 32768 queries -- a net loss there (FP4's three_nn: stage 100 -> 124 us), a net win from 65536 queries per
// cloud up (nn_distance at 8 x 131072^2: 5.3 -> 3.9 ms), hence the threshold.
static const float4 *sort_queries(int b, int n, int m, const float *xyz1, void *workspace, size_t workspace_bytes, cudaStream_t s, int *rc) {
    *rc = GSPN_OK;
    const size_t first = (grid_ws_bytes(b, m) + 255) & ~(size_t)255;
    if (n < kSortQueriesMin || workspace_bytes < first + grid_ws_bytes(b, n)) return nullptr;
    GridWs qs = carve((unsigned char *)workspace + first, b, n);
    *rc = build_grid(b, n, xyz1, 0.f, (float)n * 0.25f, qs, s);
    return *rc == GSPN_OK ? qs.sorted : nullptr;
}

int gspn_three_nn_grid_launch(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx, float *weight, void *workspace,
                              size_t workspace_bytes, cudaStream_t s) {
    GridWs ws = carve(workspace, b, m);
    // about one known point per cell: the 3x3x3 block then holds the three nearest for almost every query
    float target = (float)m;
    int rc = build_grid(b, m, xyz2, 0.f, target, ws, s);
    if (rc != GSPN_OK) return rc;
    const float4 *qsorted = sort_queries(b, n, m, xyz1, workspace, workspace_bytes, s, &rc);
    if (rc != GSPN_OK) return rc;
    dim3 grid(ceil_div(n, 128), b);
    three_nn_grid_kernel<false, false><<<grid, 128, 0, s>>>(n, m, grid_cap(m), xyz1, ws.hdr, ws.cell_start, ws.sorted, dist, idx, weight, qsorted);
    return check_launch();
}

// one direction of nn_distance through the grid (grid over xyz2, queries xyz1)
int gspn_nn_one_way_grid_launch(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx, int fma, void *workspace,
                                size_t workspace_bytes, cudaStream_t s) {
    GridWs ws = carve(workspace, b, m);
    float target = (float)m;
    int rc = build_grid(b, m, xyz2, 0.f, target, ws, s);
    if (rc != GSPN_OK) return rc;
    const float4 *qsorted = sort_queries(b, n, m, xyz1, workspace, workspace_bytes, s, &rc);
    if (rc != GSPN_OK) return rc;
    dim3 grid(ceil_div(n, 128), b);
    if (fma) three_nn_grid_kernel<true, true><<<grid, 128, 0, s>>>(n, m, grid_cap(m), xyz1, ws.hdr, ws.cell_start, ws.sorted, dist, idx, nullptr, qsorted);
    else three_nn_grid_kernel<true, false><<<grid, 128, 0, s>>>(n, m, grid_cap(m), xyz1, ws.hdr, ws.cell_start, ws.sorted, dist, idx, nullptr, qsorted);
    return check_launch();
}
