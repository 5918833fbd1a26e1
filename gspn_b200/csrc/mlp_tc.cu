// mlp_tc.cu -- tcgen05 / TMEM shared-MLP chain (placeholder entry points until the kernel lands).
#include "common.cuh"
using namespace gspn;
extern "C" size_t gspn_mlp_weight_image_bytes(int cin_padded, int cout) {
    if (cin_padded <= 0 || cout <= 0 || cin_padded % 64) return 0;
    return (size_t)(cin_padded / 64) * (size_t)((cout + 7) / 8 * 8) * 128;
}
extern "C" int gspn_mlp_pack_weights(int, int, int, const float *, const int *, void *, gspn_stream_t) { return GSPN_E_UNSUPPORTED; }
extern "C" int gspn_mlp_chain(long, int, const int *, const void *, const void *const *, const float *const *, const float *const *,
                              const int *, int, float *, void *, gspn_stream_t) { return GSPN_E_UNSUPPORTED; }
extern "C" int gspn_fp_assemble(int, int, int, int, int, const float *, const float *, const int *, const float *, void *, int,
                                gspn_stream_t) { return GSPN_E_UNSUPPORTED; }
