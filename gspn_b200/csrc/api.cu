// api.cu -- library-level entry points of include/gspn_b200.h (version, error text).
#include "common.cuh"

namespace gspn {
static thread_local cudaError_t g_last_cuda_error = cudaSuccess;
void set_last_cuda_error(cudaError_t e) { g_last_cuda_error = e; }
}  // namespace gspn

extern "C" int gspn_version(void) { return 2000; }

// ---- measurement aid: dependent-free FFMA stream (8 accumulators per thread, packed pairs so that FFMA2 can issue), the
// denominator of the "pair evaluations / s" roofline of the search kernels (FPS, ball query, three_nn, nn_distance)
namespace gspn {
__global__ void __launch_bounds__(256) fma_peak_kernel(int iters, float seed, float *out) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = seed + (float)(threadIdx.x + i);
    const float m = 1.0000001f, c = 1e-7f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], m, c);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 123.456f) out[0] = s;  // never true: keeps the loop alive
}
}  // namespace gspn

extern "C" int gspn_fp32_peak_probe(int blocks, int iters, float *scratch, double *flops_out, gspn_stream_t stream) {
    GSPN_REQUIRE(blocks > 0 && iters > 0);
    GSPN_REQUIRE_PTR(scratch); GSPN_REQUIRE_PTR(flops_out);
    gspn::fma_peak_kernel<<<blocks, 256, 0, gspn::as_stream(stream)>>>(iters, 1.0f, scratch);
    *flops_out = 2.0 * 16.0 * (double)iters * 256.0 * (double)blocks;
    return gspn::check_launch();
}

extern "C" const char *gspn_last_cuda_error(void) { return cudaGetErrorString(gspn::g_last_cuda_error); }

extern "C" const char *gspn_error_string(int code) {
    switch (code) {
        case GSPN_OK: return "ok";
        case GSPN_E_BAD_SHAPE: return "invalid argument: shape or attribute check failed";
        case GSPN_E_NULL_PTR: return "invalid argument: null pointer";
        case GSPN_E_BAD_DTYPE: return "invalid argument: unknown dtype";
        case GSPN_E_WORKSPACE: return "workspace missing or too small";
        case GSPN_E_CUDA: return "CUDA error (see gspn_last_cuda_error)";
        case GSPN_E_UNSUPPORTED: return "valid request outside what this build implements";
    }
    return "unknown error code";
}
