// mlp_f32.cu -- reference-precision (fp32, CUDA-core) shared-MLP layer:
//   y = act((x @ w) * scale + shift)  [+ max over groups of `pool` consecutive rows]
// i.e. tf_util.conv2d 1x1 + bias + batch_norm(inference) + ReLU (utils/tf_util.py:170-184,530-534)
// and tf.reduce_max over nsample (utils/pointnet_util.py:124).  This is the path the fp32 parity
// tests pin (tolerance 1e-5); the throughput path is the tcgen05 chain in mlp_tc.cu.
#include <cfloat>
#include "common.cuh"

namespace gspn {

constexpr int BM = 64, BN = 64, BK = 16;

// 256 threads, each a 4x4 micro-tile.  pool must divide BM (or be 1).
__global__ void __launch_bounds__(256) mlp_layer_f32_kernel(long rows, int cin, int cout, const float *__restrict__ x, int ldx,
                                                            const float *__restrict__ w, const float *__restrict__ scale,
                                                            const float *__restrict__ shift, int relu, int pool, float *__restrict__ y) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    __shared__ float Cs[BM][BN + 1];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const long row0 = (long)blockIdx.x * BM;
    const int col0 = blockIdx.y * BN;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < cin; k0 += BK) {
        // A tile: BM x BK (x row-major, stride ldx); 1024 elements / 256 threads
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            int e = tid + t * 256;
            int r = e >> 4, k = e & 15;
            long gr = row0 + r;
            As[k][r] = (gr < rows && k0 + k < cin) ? __ldg(x + gr * ldx + k0 + k) : 0.f;
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            int e = tid + t * 256;
            int k = e >> 6, c = e & 63;
            Bs[k][c] = (k0 + k < cin && col0 + c < cout) ? __ldg(w + (size_t)(k0 + k) * cout + col0 + c) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int c = col0 + tx * 4 + j;
        float sc = c < cout ? __ldg(scale + c) : 0.f, sh = c < cout ? __ldg(shift + c) : 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float v = fmaf(acc[i][j], sc, sh);
            if (relu) v = fmaxf(v, 0.f);
            acc[i][j] = v;
        }
    }
    if (pool <= 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            long r = row0 + ty * 4 + i;
            if (r >= rows) continue;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int c = col0 + tx * 4 + j;
                if (c < cout) y[r * cout + c] = acc[i][j];
            }
        }
        return;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) Cs[ty * 4 + i][tx * 4 + j] = acc[i][j];
    __syncthreads();
    const int groups = BM / pool;
    for (int e = tid; e < groups * BN; e += 256) {
        int gi = e / BN, c = e - gi * BN;
        long r = row0 + (long)gi * pool;
        if (r >= rows || col0 + c >= cout) continue;
        float v = -FLT_MAX;
        for (int s = 0; s < pool && r + s < rows; ++s) v = fmaxf(v, Cs[gi * pool + s][c]);
        y[(r / pool) * cout + col0 + c] = v;
    }
}

// y[g,c] = max_s x[g*k+s, c]
__global__ void __launch_bounds__(256) max_pool_rows_kernel(long groups, int k, int c, const float *__restrict__ x, float *__restrict__ y) {
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < groups * c; e += (long)gridDim.x * blockDim.x) {
        long gi = e / c;
        int col = (int)(e - gi * c);
        const float *src = x + gi * k * c + col;
        float v = __ldg(src);
        for (int s = 1; s < k; ++s) v = fmaxf(v, __ldg(src + (size_t)s * c));
        y[e] = v;
    }
}

}  // namespace gspn

using namespace gspn;

extern "C" int gspn_mlp_layer_f32(long rows, int cin, int cout, const float *x, int ldx, const float *w, const float *scale,
                                  const float *shift, int relu, int pool, float *y, gspn_stream_t stream) {
    GSPN_REQUIRE(rows >= 0 && cin > 0 && cout > 0 && ldx >= cin && pool >= 1);
    if (rows == 0) return GSPN_OK;
    GSPN_REQUIRE_PTR(x); GSPN_REQUIRE_PTR(w); GSPN_REQUIRE_PTR(scale); GSPN_REQUIRE_PTR(shift); GSPN_REQUIRE_PTR(y);
    if (pool > 1 && (BM % pool != 0 || rows % pool != 0)) return GSPN_E_UNSUPPORTED;  // caller pools with gspn_max_pool_rows
    long bx = ceil_div_l(rows, BM);
    GSPN_REQUIRE(bx < (1L << 31));
    dim3 grid((unsigned)bx, ceil_div(cout, BN));
    mlp_layer_f32_kernel<<<grid, 256, 0, as_stream(stream)>>>(rows, cin, cout, x, ldx, w, scale, shift, relu, pool, y);
    return check_launch();
}

extern "C" int gspn_max_pool_rows(long groups, int k, int c, const float *x, float *y, gspn_stream_t stream) {
    GSPN_REQUIRE(groups >= 0 && k > 0 && c > 0);
    if (groups == 0) return GSPN_OK;
    GSPN_REQUIRE_PTR(x); GSPN_REQUIRE_PTR(y);
    long total = groups * c;
    long blk = ceil_div_l(total, 256);
    if (blk > 148L * 64) blk = 148L * 64;
    max_pool_rows_kernel<<<(unsigned)blk, 256, 0, as_stream(stream)>>>(groups, k, c, x, y);
    return check_launch();
}
