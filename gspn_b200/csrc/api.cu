// api.cu -- library-level entry points of include/gspn_b200.h (version, error text).
#include "common.cuh"

namespace gspn {
static thread_local cudaError_t g_last_cuda_error = cudaSuccess;
void set_last_cuda_error(cudaError_t e) { g_last_cuda_error = e; }
}  // namespace gspn

extern "C" int gspn_version(void) { return 1000; }

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
