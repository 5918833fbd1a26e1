// nn_search.cu -- brute-force nearest-neighbour scans of the FP path: three_nn (top-3) and
// nn_distance (top-1, both directions).  FP32-ALU bound (O(n*m) pair evaluations), not HBM bound:
// each CTA stages the scanned cloud through shared memory in float4-padded tiles so one
// broadcast LDS.128 feeds 32 lanes, and every thread owns kQ query points to amortise it.
#include <math_constants.h>
#include "common.cuh"

namespace gspn {

constexpr int kNNThreads = 128;
constexpr int kNNTile = 1024;  // scanned points per shared-memory tile (16 KiB as float4)

// cooperative AoS (x,y,z)*cnt -> float4 smem tile
__device__ __forceinline__ void load_tile(float4 *tile, const float *__restrict__ src, int cnt) {
    for (int k = threadIdx.x; k < cnt; k += blockDim.x)
        tile[k] = make_float4(__ldg(src + 3 * k), __ldg(src + 3 * k + 1), __ldg(src + 3 * k + 2), 0.f);
}

// ---- three_nn (tf_interpolate.cpp:60-103): strict '<' insertion into an ascending top-3,
// float distance (== the reference's double compare of float values), bests start above
// every finite float (1e40 -> +inf when stored as float).
template <int Q>
__global__ void __launch_bounds__(kNNThreads) three_nn_kernel(int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                                                             float *__restrict__ dist, int *__restrict__ idx, float *__restrict__ weight) {
    __shared__ float4 tile[kNNTile];
    const int cloud = blockIdx.y;
    const float *u = xyz1 + (size_t)cloud * n * 3;
    const float *kn = xyz2 + (size_t)cloud * m * 3;
    const int j0 = (blockIdx.x * kNNThreads + threadIdx.x) * Q;

    float x1[Q], y1[Q], z1[Q], b1[Q], b2[Q], b3[Q];
    int i1[Q], i2[Q], i3[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        int j = min(j0 + q, n - 1);
        x1[q] = __ldg(u + 3 * j); y1[q] = __ldg(u + 3 * j + 1); z1[q] = __ldg(u + 3 * j + 2);
        b1[q] = b2[q] = b3[q] = CUDART_INF_F;
        i1[q] = i2[q] = i3[q] = 0;
    }
    for (int k0 = 0; k0 < m; k0 += kNNTile) {
        int cnt = min(kNNTile, m - k0);
        __syncthreads();
        load_tile(tile, kn + (size_t)k0 * 3, cnt);
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < cnt; ++k) {
            float4 p = tile[k];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                float d = sqdist_nofma(p.x, p.y, p.z, x1[q], y1[q], z1[q]);
                if (d < b3[q]) {
                    int kk = k0 + k;
                    if (d < b1[q]) { b3[q] = b2[q]; i3[q] = i2[q]; b2[q] = b1[q]; i2[q] = i1[q]; b1[q] = d; i1[q] = kk; }
                    else if (d < b2[q]) { b3[q] = b2[q]; i3[q] = i2[q]; b2[q] = d; i2[q] = kk; }
                    else { b3[q] = d; i3[q] = kk; }
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        int j = j0 + q;
        if (j < n) {
            size_t o = ((size_t)cloud * n + j) * 3;
            dist[o] = b1[q]; dist[o + 1] = b2[q]; dist[o + 2] = b3[q];
            idx[o] = i1[q]; idx[o + 1] = i2[q]; idx[o + 2] = i3[q];
            if (weight) {
                // pointnet_util.py:157-160: dist=max(dist,1e-10); w=(1/dist)/sum(1/dist)
                float r1 = __fdiv_rn(1.0f, fmaxf(b1[q], 1e-10f));
                float r2 = __fdiv_rn(1.0f, fmaxf(b2[q], 1e-10f));
                float r3 = __fdiv_rn(1.0f, fmaxf(b3[q], 1e-10f));
                float norm = __fadd_rn(__fadd_rn(r1, r2), r3);
                weight[o] = __fdiv_rn(r1, norm); weight[o + 1] = __fdiv_rn(r2, norm); weight[o + 2] = __fdiv_rn(r3, norm);
            }
        }
    }
}

// ---- nn_distance, one direction (tf_nndistance.cpp:21-43 / tf_nndistance_g.cu:5-127):
// ascending scan, `k==0 || d<best`, so the lowest index wins ties.
template <int Q, bool FMA>
__global__ void __launch_bounds__(kNNThreads) nn_one_way_kernel(int n, int m, const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                                                               float *__restrict__ dist, int *__restrict__ idx) {
    __shared__ float4 tile[kNNTile];
    const int cloud = blockIdx.y;
    const float *u = xyz1 + (size_t)cloud * n * 3;
    const float *kn = xyz2 + (size_t)cloud * m * 3;
    const int j0 = (blockIdx.x * kNNThreads + threadIdx.x) * Q;
    float x1[Q], y1[Q], z1[Q], best[Q];
    int besti[Q];
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        int j = min(j0 + q, n - 1);
        x1[q] = __ldg(u + 3 * j); y1[q] = __ldg(u + 3 * j + 1); z1[q] = __ldg(u + 3 * j + 2);
        best[q] = 0.f; besti[q] = 0;
    }
    for (int k0 = 0; k0 < m; k0 += kNNTile) {
        int cnt = min(kNNTile, m - k0);
        __syncthreads();
        load_tile(tile, kn + (size_t)k0 * 3, cnt);
        __syncthreads();
        int k = 0;
        if (k0 == 0) {  // k==0 always wins (best may be NaN/inf there)
            float4 p = tile[0];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                best[q] = FMA ? sqdist_fma(p.x, p.y, p.z, x1[q], y1[q], z1[q]) : sqdist_nofma(p.x, p.y, p.z, x1[q], y1[q], z1[q]);
                besti[q] = 0;
            }
            k = 1;
        }
#pragma unroll 4
        for (; k < cnt; ++k) {
            float4 p = tile[k];
#pragma unroll
            for (int q = 0; q < Q; ++q) {
                float d = FMA ? sqdist_fma(p.x, p.y, p.z, x1[q], y1[q], z1[q]) : sqdist_nofma(p.x, p.y, p.z, x1[q], y1[q], z1[q]);
                if (d < best[q]) { best[q] = d; besti[q] = k0 + k; }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < Q; ++q) {
        int j = j0 + q;
        if (j < n) {
            dist[(size_t)cloud * n + j] = best[q];
            idx[(size_t)cloud * n + j] = besti[q];
        }
    }
}

template <bool FMA>
static void launch_nn(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx, cudaStream_t s) {
    // few query points -> 1 per thread so the grid still covers the SMs
    if ((long)b * n >= 148L * 8 * kNNThreads * 2) {
        dim3 grid(ceil_div(n, kNNThreads * 2), b);
        nn_one_way_kernel<2, FMA><<<grid, kNNThreads, 0, s>>>(n, m, xyz1, xyz2, dist, idx);
    } else {
        dim3 grid(ceil_div(n, kNNThreads), b);
        nn_one_way_kernel<1, FMA><<<grid, kNNThreads, 0, s>>>(n, m, xyz1, xyz2, dist, idx);
    }
}

}  // namespace gspn

using namespace gspn;

// grid_search.cu
int gspn_three_nn_grid_launch(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx, float *weight, void *workspace,
                              size_t workspace_bytes, cudaStream_t s);
int gspn_nn_one_way_grid_launch(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx, int fma, void *workspace,
                                size_t workspace_bytes, cudaStream_t s);
extern "C" size_t gspn_grid_workspace_bytes(int b, int n);

extern "C" int gspn_three_nn(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx, float *weight,
                             void *workspace, size_t workspace_bytes, gspn_stream_t stream) {
    GSPN_REQUIRE(b >= 0 && n >= 0 && m > 0 && b <= 65535);  // tf_interpolate.cpp:163-169
    if (b == 0 || n == 0) return GSPN_OK;
    GSPN_REQUIRE_PTR(xyz1); GSPN_REQUIRE_PTR(xyz2); GSPN_REQUIRE_PTR(dist); GSPN_REQUIRE_PTR(idx);
    cudaStream_t s = as_stream(stream);
    if (workspace != nullptr && m >= 1024 && (long)n * m >= (1L << 22)) {  // grid over the known points
        if (workspace_bytes < gspn_grid_workspace_bytes(b, m)) return GSPN_E_WORKSPACE;
        return gspn_three_nn_grid_launch(b, n, m, xyz1, xyz2, dist, idx, weight, workspace, workspace_bytes, s);
    }
    if ((long)b * n >= 148L * 8 * kNNThreads * 2) {
        dim3 grid(ceil_div(n, kNNThreads * 2), b);
        three_nn_kernel<2><<<grid, kNNThreads, 0, s>>>(n, m, xyz1, xyz2, dist, idx, weight);
    } else {
        dim3 grid(ceil_div(n, kNNThreads), b);
        three_nn_kernel<1><<<grid, kNNThreads, 0, s>>>(n, m, xyz1, xyz2, dist, idx, weight);
    }
    return check_launch();
}

extern "C" int gspn_nn_distance(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist1, int *idx1, float *dist2, int *idx2,
                                int rounding, void *workspace, size_t workspace_bytes, gspn_stream_t stream) {
    GSPN_REQUIRE(b >= 0 && n > 0 && m > 0 && b <= 65535);  // tf_nndistance.cpp:51-58
    GSPN_REQUIRE(rounding == 0 || rounding == 1);
    if (b == 0) return GSPN_OK;
    GSPN_REQUIRE_PTR(xyz1); GSPN_REQUIRE_PTR(xyz2); GSPN_REQUIRE_PTR(dist1); GSPN_REQUIRE_PTR(idx1); GSPN_REQUIRE_PTR(dist2); GSPN_REQUIRE_PTR(idx2);
    cudaStream_t s = as_stream(stream);
    if (workspace != nullptr && n >= 2048 && m >= 2048) {  // both directions through a grid over the scanned set
        if (workspace_bytes < gspn_grid_workspace_bytes(b, n > m ? n : m)) return GSPN_E_WORKSPACE;
        int rc = gspn_nn_one_way_grid_launch(b, n, m, xyz1, xyz2, dist1, idx1, rounding, workspace, workspace_bytes, s);
        if (rc != GSPN_OK) return rc;
        return gspn_nn_one_way_grid_launch(b, m, n, xyz2, xyz1, dist2, idx2, rounding, workspace, workspace_bytes, s);
    }
    if (rounding) {
        launch_nn<true>(b, n, m, xyz1, xyz2, dist1, idx1, s);
        launch_nn<true>(b, m, n, xyz2, xyz1, dist2, idx2, s);
    } else {
        launch_nn<false>(b, n, m, xyz1, xyz2, dist1, idx1, s);
        launch_nn<false>(b, m, n, xyz2, xyz1, dist2, idx2, s);
    }
    return check_launch();
}

// One direction only: for every query the nearest reference point (lowest index on ties).  The model's brute-force
// nearest-seed / nearest-cropped-point argmins (models/model_rpointnet.py:1136 and :1032-1033: argmin over
// reduce_sum(square(a - b), -1)) and test.py's sklearn ball-tree 1-NN (test.py:165-166,184-185) are this search.
extern "C" int gspn_nearest_point(int b, int n, int m, const float *queries, const float *refs, float *dist, int *idx, int rounding,
                                  void *workspace, size_t workspace_bytes, gspn_stream_t stream) {
    GSPN_REQUIRE(b >= 0 && n >= 0 && m > 0 && b <= 65535);
    GSPN_REQUIRE(rounding == 0 || rounding == 1);
    if (b == 0 || n == 0) return GSPN_OK;
    GSPN_REQUIRE_PTR(queries); GSPN_REQUIRE_PTR(refs); GSPN_REQUIRE_PTR(dist); GSPN_REQUIRE_PTR(idx);
    cudaStream_t s = as_stream(stream);
    if (workspace != nullptr && m >= 2048 && (long)n * m >= (1L << 22)) {  // grid over the reference set
        if (workspace_bytes < gspn_grid_workspace_bytes(b, m)) return GSPN_E_WORKSPACE;
        return gspn_nn_one_way_grid_launch(b, n, m, queries, refs, dist, idx, rounding, workspace, workspace_bytes, s);
    }
    if (rounding) launch_nn<true>(b, n, m, queries, refs, dist, idx, s);
    else launch_nn<false>(b, n, m, queries, refs, dist, idx, s);
    return check_launch();
}

