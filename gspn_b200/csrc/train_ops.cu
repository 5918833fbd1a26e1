// train_ops.cu -- fp32 kernels of the TRAINING form of the shared MLP (SURVEY.md 8f rank 1 / BASELINE config 4):
// tf_util.conv2d 1x1 + bias -> tf.contrib.layers.batch_norm with batch moments (utils/tf_util.py:170-184,515-534)
// -> ReLU -> tf.reduce_max over nsample (utils/pointnet_util.py:109-124), and their backward passes.  The GEMMs
// reuse gspn_mlp_layer_f32 (forward and dX); here: column statistics, BN+ReLU apply, max-pool with argmax, the BN/ReLU/
// pool backward, the weight gradient (split-K over rows) and the strided scatter of grouped-row gradients.
#include "common.cuh"

namespace gspn {

// ---- column sums: s1[c] = sum_r f(r,c), s2[c] = sum_r g(r,c); double accumulation across CTAs
// MODE 0: f = z, g = z*z                       (batch moments)
// MODE 1: f = dy', g = dy' * xhat              (BN backward sums) with dy' the ReLU/pool-masked upstream gradient
struct BwdArgs {
    const float *z;       // (rows,c) pre-BN activations saved by the forward
    const float *dy;      // pool==1: (rows,c); pool>1: (rows/pool,c)
    const int *argmax;    // pool>1: (rows/pool,c) winning row offset inside the group
    const float *mean, *invstd, *gamma, *beta;
    int pool, relu;
};

__device__ __forceinline__ float masked_dy(const BwdArgs &a, long r, int col, int c, float zv) {
    float y = fmaf((zv - a.mean[col]) * a.invstd[col], a.gamma[col], a.beta[col]);
    if (a.relu && !(y > 0.f)) return 0.f;
    if (a.pool > 1) {
        long g = r / a.pool;
        int k = (int)(r - g * a.pool);
        return a.argmax[g * c + col] == k ? a.dy[g * c + col] : 0.f;
    }
    return a.dy[r * c + col];
}

template <int MODE>
__global__ void __launch_bounds__(256) col_sums_kernel(long rows, int c, const float *__restrict__ z, BwdArgs a, double *__restrict__ s1,
                                                       double *__restrict__ s2) {
    __shared__ float sh1[4][64], sh2[4][64];
    const int col = blockIdx.y * 64 + threadIdx.x;
    const long r0 = (long)blockIdx.x * 256;
    float p1 = 0.f, p2 = 0.f;
    if (col < c) {
        for (long r = r0 + threadIdx.y; r < rows && r < r0 + 256; r += 4) {
            float zv = z[r * c + col];
            if (MODE == 0) { p1 += zv; p2 = fmaf(zv, zv, p2); }
            else {
                float d = masked_dy(a, r, col, c, zv);
                p1 += d;
                p2 = fmaf(d, (zv - a.mean[col]) * a.invstd[col], p2);
            }
        }
    }
    sh1[threadIdx.y][threadIdx.x] = p1;
    sh2[threadIdx.y][threadIdx.x] = p2;
    __syncthreads();
    if (threadIdx.y == 0 && col < c) {
        float t1 = sh1[0][threadIdx.x] + sh1[1][threadIdx.x] + sh1[2][threadIdx.x] + sh1[3][threadIdx.x];
        float t2 = sh2[0][threadIdx.x] + sh2[1][threadIdx.x] + sh2[2][threadIdx.x] + sh2[3][threadIdx.x];
        atomicAdd(s1 + col, (double)t1);
        atomicAdd(s2 + col, (double)t2);
    }
}

// ---- y = act((z - mean) * invstd * gamma + beta)
__global__ void __launch_bounds__(256) bn_act_kernel(long total, int c, const float *__restrict__ z, const float *__restrict__ mean,
                                                     const float *__restrict__ invstd, const float *__restrict__ gamma,
                                                     const float *__restrict__ beta, int relu, float *__restrict__ y) {
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        int col = (int)(e % c);
        float v = fmaf((z[e] - mean[col]) * invstd[col], gamma[col], beta[col]);
        y[e] = relu ? fmaxf(v, 0.f) : v;
    }
}

// ---- max over groups of k rows with argmax (first maximum wins, like tf.reduce_max's gradient convention of ties is
// irrelevant for continuous inputs)
__global__ void __launch_bounds__(256) maxpool_argmax_kernel(long groups, int k, int c, const float *__restrict__ y, float *__restrict__ out,
                                                             int *__restrict__ argmax) {
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < groups * c; e += (long)gridDim.x * blockDim.x) {
        long g = e / c;
        int col = (int)(e - g * c);
        const float *src = y + g * k * c + col;
        float best = src[0];
        int bi = 0;
        for (int s = 1; s < k; ++s) {
            float v = src[(size_t)s * c];
            if (v > best) { best = v; bi = s; }
        }
        out[e] = best;
        argmax[e] = bi;
    }
}

// ---- dz = gamma*invstd * (dy' - s1/N - xhat*s2/N)      (batch-norm backward through the batch moments)
__global__ void __launch_bounds__(256) bn_bwd_kernel(long rows, long norm_rows, int c, BwdArgs a, const double *__restrict__ s1,
                                                     const double *__restrict__ s2, float *__restrict__ dz) {
    const long total = rows * c;
    const float invn = 1.0f / (float)norm_rows;  // rows of the WHOLE batch the moments were taken over (all ranks under SyncBN)
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        long r = e / c;
        int col = (int)(e - r * c);
        float zv = a.z[e];
        float xhat = (zv - a.mean[col]) * a.invstd[col];
        float d = masked_dy(a, r, col, c, zv);
        dz[e] = a.gamma[col] * a.invstd[col] * (d - (float)s1[col] * invn - xhat * (float)s2[col] * invn);
    }
}

// ---- weight gradient dW (cin,cout) += x^T (rows,cin; stride ldx) * dz (rows,cout), split over row slabs; dbias = colsum(dz)
constexpr int WG_ROWS = 512;
__global__ void __launch_bounds__(256) wgrad_kernel(long rows, int cin, int cout, const float *__restrict__ x, int ldx,
                                                    const float *__restrict__ dz, float *__restrict__ dW, float *__restrict__ dbias) {
    // CTA: 16x16 output tile of dW (blockIdx.y, blockIdx.z), rows slab blockIdx.x; threads (16,16)
    __shared__ float xs[32][17], ds[32][17];
    const int ti = threadIdx.y, tj = threadIdx.x;
    const int i0 = blockIdx.y * 16, j0 = blockIdx.z * 16;
    const long r0 = (long)blockIdx.x * WG_ROWS, r1 = min(rows, r0 + WG_ROWS);
    float acc = 0.f, bsum = 0.f;
    for (long rb = r0; rb < r1; rb += 32) {
        // stage 32 rows x 16 columns of x and dz
        for (int e = ti * 16 + tj; e < 32 * 16; e += 256) {
            int rr = e >> 4, cc = e & 15;
            long r = rb + rr;
            xs[rr][cc] = (r < r1 && i0 + cc < cin) ? x[r * ldx + i0 + cc] : 0.f;
            ds[rr][cc] = (r < r1 && j0 + cc < cout) ? dz[r * cout + j0 + cc] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < 32; ++rr) {
            acc = fmaf(xs[rr][ti], ds[rr][tj], acc);
            if (ti == 0) bsum += ds[rr][tj];
        }
        __syncthreads();
    }
    if (i0 + ti < cin && j0 + tj < cout) atomicAdd(dW + (size_t)(i0 + ti) * cout + j0 + tj, acc);
    if (dbias && blockIdx.y == 0 && ti == 0 && j0 + tj < cout) atomicAdd(dbias + j0 + tj, bsum);
}

// ---- grad of the fused grouping rows w.r.t. the features: grad_points[b, idx[b,j,s], :c] += grad_rows[(b,j,s), :c]
__global__ void __launch_bounds__(256) group_rows_grad_kernel(long total, int n, int c, long mk, int ld, const float *__restrict__ grad_rows,
                                                              const int *__restrict__ idx, float *__restrict__ grad_points) {
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        long row = e / c;
        int l = (int)(e - row * c);
        long bi = row / mk;
        atomicAdd(grad_points + (bi * n + __ldg(idx + row)) * c + l, __ldg(grad_rows + row * ld + l));
    }
}

static inline unsigned ew_blocks(long total) {
    long b = ceil_div_l(total, 256);
    return (unsigned)(b < 148L * 32 ? (b > 0 ? b : 1) : 148L * 32);
}

}  // namespace gspn

using namespace gspn;

extern "C" int gspn_col_moments_f32(long rows, int c, const float *z, double *sum, double *sumsq, gspn_stream_t stream) {
    GSPN_REQUIRE(rows > 0 && c > 0);
    GSPN_REQUIRE_PTR(z); GSPN_REQUIRE_PTR(sum); GSPN_REQUIRE_PTR(sumsq);
    cudaStream_t s = as_stream(stream);
    GSPN_CUDA_OK(cudaMemsetAsync(sum, 0, sizeof(double) * c, s));
    GSPN_CUDA_OK(cudaMemsetAsync(sumsq, 0, sizeof(double) * c, s));
    dim3 grid((unsigned)ceil_div_l(rows, 256), ceil_div(c, 64)), block(64, 4);
    BwdArgs a = {};
    col_sums_kernel<0><<<grid, block, 0, s>>>(rows, c, z, a, sum, sumsq);
    return check_launch();
}

extern "C" int gspn_bn_act_f32(long rows, int c, const float *z, const float *mean, const float *invstd, const float *gamma, const float *beta,
                               int relu, float *y, gspn_stream_t stream) {
    GSPN_REQUIRE(rows > 0 && c > 0);
    GSPN_REQUIRE_PTR(z); GSPN_REQUIRE_PTR(mean); GSPN_REQUIRE_PTR(invstd); GSPN_REQUIRE_PTR(gamma); GSPN_REQUIRE_PTR(beta); GSPN_REQUIRE_PTR(y);
    bn_act_kernel<<<ew_blocks(rows * c), 256, 0, as_stream(stream)>>>(rows * c, c, z, mean, invstd, gamma, beta, relu, y);
    return check_launch();
}

extern "C" int gspn_maxpool_argmax_f32(long groups, int k, int c, const float *y, float *out, int *argmax, gspn_stream_t stream) {
    GSPN_REQUIRE(groups > 0 && k > 0 && c > 0);
    GSPN_REQUIRE_PTR(y); GSPN_REQUIRE_PTR(out); GSPN_REQUIRE_PTR(argmax);
    maxpool_argmax_kernel<<<ew_blocks(groups * c), 256, 0, as_stream(stream)>>>(groups, k, c, y, out, argmax);
    return check_launch();
}

extern "C" int gspn_bn_act_pool_bwd_f32(long rows, int c, int pool, int relu, int bn, const float *z, const float *dy, const int *argmax,
                                        const float *mean, const float *invstd, const float *gamma, const float *beta, double *s1, double *s2,
                                        float *dz, float *dgamma, float *dbeta, gspn_stream_t stream) {
    GSPN_REQUIRE(rows > 0 && c > 0 && pool >= 1 && rows % pool == 0);
    GSPN_REQUIRE_PTR(z); GSPN_REQUIRE_PTR(dy); GSPN_REQUIRE_PTR(mean); GSPN_REQUIRE_PTR(invstd); GSPN_REQUIRE_PTR(gamma); GSPN_REQUIRE_PTR(beta);
    GSPN_REQUIRE_PTR(s1); GSPN_REQUIRE_PTR(s2); GSPN_REQUIRE_PTR(dz);
    if (pool > 1) GSPN_REQUIRE_PTR(argmax);
    (void)dgamma; (void)dbeta;  // dgamma = s2, dbeta = s1: the caller reads them from the sums
    cudaStream_t s = as_stream(stream);
    GSPN_CUDA_OK(cudaMemsetAsync(s1, 0, sizeof(double) * c, s));
    GSPN_CUDA_OK(cudaMemsetAsync(s2, 0, sizeof(double) * c, s));
    BwdArgs a = {z, dy, argmax, mean, invstd, gamma, beta, pool, relu};
    dim3 grid((unsigned)ceil_div_l(rows, 256), ceil_div(c, 64)), block(64, 4);
    // bn == 0 (bn=False call sites): mean=0, invstd=gamma=1, beta=bias-free identity; the sums stay zero so dz = dy'
    if (bn) col_sums_kernel<1><<<grid, block, 0, s>>>(rows, c, z, a, s1, s2);
    bn_bwd_kernel<<<ew_blocks(rows * c), 256, 0, s>>>(rows, rows, c, a, s1, s2, dz);
    return check_launch();
}

// The same backward in two calls, for batch norm over the WHOLE batch of a data-parallel job (SyncBN): the caller all-reduces
// s1 / s2 between them and passes the global row count.
extern "C" int gspn_bn_bwd_sums_f32(long rows, int c, int pool, int relu, const float *z, const float *dy, const int *argmax, const float *mean,
                                    const float *invstd, const float *gamma, const float *beta, double *s1, double *s2, gspn_stream_t stream) {
    GSPN_REQUIRE(rows > 0 && c > 0 && pool >= 1 && rows % pool == 0);
    GSPN_REQUIRE_PTR(z); GSPN_REQUIRE_PTR(dy); GSPN_REQUIRE_PTR(mean); GSPN_REQUIRE_PTR(invstd); GSPN_REQUIRE_PTR(gamma); GSPN_REQUIRE_PTR(beta);
    GSPN_REQUIRE_PTR(s1); GSPN_REQUIRE_PTR(s2);
    if (pool > 1) GSPN_REQUIRE_PTR(argmax);
    cudaStream_t s = as_stream(stream);
    GSPN_CUDA_OK(cudaMemsetAsync(s1, 0, sizeof(double) * c, s));
    GSPN_CUDA_OK(cudaMemsetAsync(s2, 0, sizeof(double) * c, s));
    BwdArgs a = {z, dy, argmax, mean, invstd, gamma, beta, pool, relu};
    dim3 grid((unsigned)ceil_div_l(rows, 256), ceil_div(c, 64)), block(64, 4);
    col_sums_kernel<1><<<grid, block, 0, s>>>(rows, c, z, a, s1, s2);
    return check_launch();
}

extern "C" int gspn_bn_bwd_apply_f32(long rows, long total_rows, int c, int pool, int relu, const float *z, const float *dy, const int *argmax,
                                     const float *mean, const float *invstd, const float *gamma, const float *beta, const double *s1,
                                     const double *s2, float *dz, gspn_stream_t stream) {
    GSPN_REQUIRE(rows > 0 && total_rows >= rows && c > 0 && pool >= 1 && rows % pool == 0);
    GSPN_REQUIRE_PTR(z); GSPN_REQUIRE_PTR(dy); GSPN_REQUIRE_PTR(mean); GSPN_REQUIRE_PTR(invstd); GSPN_REQUIRE_PTR(gamma); GSPN_REQUIRE_PTR(beta);
    GSPN_REQUIRE_PTR(s1); GSPN_REQUIRE_PTR(s2); GSPN_REQUIRE_PTR(dz);
    if (pool > 1) GSPN_REQUIRE_PTR(argmax);
    BwdArgs a = {z, dy, argmax, mean, invstd, gamma, beta, pool, relu};
    bn_bwd_kernel<<<ew_blocks(rows * c), 256, 0, as_stream(stream)>>>(rows, total_rows, c, a, s1, s2, dz);
    return check_launch();
}

extern "C" int gspn_mlp_wgrad_f32(long rows, int cin, int cout, const float *x, int ldx, const float *dz, float *dW, float *dbias,
                                  gspn_stream_t stream) {
    GSPN_REQUIRE(rows > 0 && cin > 0 && cout > 0 && ldx >= cin);
    GSPN_REQUIRE_PTR(x); GSPN_REQUIRE_PTR(dz); GSPN_REQUIRE_PTR(dW);
    cudaStream_t s = as_stream(stream);
    GSPN_CUDA_OK(cudaMemsetAsync(dW, 0, sizeof(float) * (size_t)cin * cout, s));
    if (dbias) GSPN_CUDA_OK(cudaMemsetAsync(dbias, 0, sizeof(float) * cout, s));
    dim3 grid((unsigned)ceil_div_l(rows, WG_ROWS), ceil_div(cin, 16), ceil_div(cout, 16)), block(16, 16);
    wgrad_kernel<<<grid, block, 0, s>>>(rows, cin, cout, x, ldx, dz, dW, dbias);
    return check_launch();
}

extern "C" int gspn_group_rows_grad(int b, int n, int c, int m, int nsample, int ld, const float *grad_rows, const int *idx, float *grad_points,
                                    gspn_stream_t stream) {
    GSPN_REQUIRE(b >= 0 && n > 0 && c > 0 && m >= 0 && nsample > 0 && ld >= c);
    if (b == 0) return GSPN_OK;
    GSPN_REQUIRE_PTR(grad_points);
    cudaStream_t s = as_stream(stream);
    GSPN_CUDA_OK(cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)b * n * c, s));
    if (m == 0) return GSPN_OK;
    GSPN_REQUIRE_PTR(grad_rows); GSPN_REQUIRE_PTR(idx);
    long mk = (long)m * nsample, total = (long)b * mk * c;
    group_rows_grad_kernel<<<ew_blocks(total), 256, 0, s>>>(total, n, c, mk, ld, grad_rows, idx, grad_points);
    return check_launch();
}
