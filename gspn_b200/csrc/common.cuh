// common.cuh -- shared helpers for the sm_100a kernels behind include/gspn_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include "../../include/gspn_b200.h"

#define GSPN_FULL_MASK 0xffffffffu

namespace gspn {

// cudaGetLastError text of the last failed launch on this thread (gspn_last_cuda_error()).
void set_last_cuda_error(cudaError_t e);

inline int check_launch() {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_last_cuda_error(e); return GSPN_E_CUDA; }
    return GSPN_OK;
}

inline cudaStream_t as_stream(gspn_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

#define GSPN_REQUIRE_PTR(p) do { if ((p) == nullptr) return GSPN_E_NULL_PTR; } while (0)
#define GSPN_REQUIRE(cond) do { if (!(cond)) return GSPN_E_BAD_SHAPE; } while (0)
#define GSPN_CUDA_OK(expr) do { cudaError_t e__ = (expr); if (e__ != cudaSuccess) { gspn::set_last_cuda_error(e__); return GSPN_E_CUDA; } } while (0)

// Squared distance with the rounding of the reference's compiled CUDA kernels
// (nvcc contracts dx*dx+dy*dy+dz*dz to  t=dy*dy; t=fma(dx,dx,t); fma(dz,dz,t);
//  tf_sampling_g.cu:142, tf_grouping_g.cu:27, tf_nndistance_g.cu:26).
__device__ __forceinline__ float sqdist_fma(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    float t = __fmul_rn(dy, dy);
    t = __fmaf_rn(dx, dx, t);
    return __fmaf_rn(dz, dz, t);
}

// Squared distance with the rounding of the reference's g++ -O2 CPU loops
// ((xx+yy)+zz, every product rounded; tf_interpolate.cpp:73, tf_nndistance.cpp:30-33).
__device__ __forceinline__ float sqdist_nofma(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = __fsub_rn(ax, bx), dy = __fsub_rn(ay, by), dz = __fsub_rn(az, bz);
    return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
}

// Division by a runtime constant for values below 2^31: q = umulhi(x, mul) >> shr (round-up magic number, exact for
// 0 <= x < 2^31 and 1 <= d < 2^31).  Element-indexed kernels use it instead of a 64-bit divide per element.
struct FastDiv {
    uint32_t d, mul, shr;
    FastDiv() : d(1), mul(0), shr(0) {}
    explicit FastDiv(uint32_t den) : d(den), mul(0), shr(0) {
        if (den > 1) {
            uint32_t lg = 0;
            while ((1ull << lg) < den) ++lg;  // ceil(log2(den))
            const unsigned long long pw = 1ull << (31 + lg);
            mul = (uint32_t)((pw + den - 1) / den);
            shr = lg - 1;
        }
    }
    __host__ __device__ __forceinline__ uint32_t div(uint32_t x) const {
#ifdef __CUDA_ARCH__
        return d == 1 ? x : (__umulhi(x, mul) >> shr);
#else
        return d == 1 ? x : (uint32_t)(((unsigned long long)x * mul) >> 32) >> shr;
#endif
    }
};

__host__ __device__ inline long ceil_div_l(long a, long b) { return (a + b - 1) / b; }
__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// ---- bf16 tile image ("tensor-core-ready layout") --------------------------------
// A (rows x ld) bf16 matrix, ld a multiple of 64, is stored as 128-row x 64-col blocks
// of 16 KiB, block (t, kb) at byte offset (t*(ld/64)+kb)*16384.  Inside a block, row r,
// 16-byte chunk c (8 bf16) lives at  (r/8)*1024 + (r%8)*128 + ((c ^ (r%8))*16)
// -- exactly the K-major SWIZZLE_128B shared-memory layout tcgen05.mma reads, so a block
// is moved HBM->smem by one 16 KiB cp.async.bulk with no tensor map.
// SPLIT image (GSPN_DT_BF16X2, the bf16x3 arithmetic): every block is a PAIR [hi block | lo block] of 32 KiB, where
// hi = bf16(x) and lo = bf16(x - hi); one 32 KiB cp.async.bulk moves both.
constexpr int kTileRows = 128;
constexpr int kTileCols = 64;
constexpr int kTileBytes = kTileRows * kTileCols * 2;

__host__ __device__ inline size_t tile_chunk_offset(long row, int col8 /* column / 8 */, int ld, int split = 0) {
    long t = row >> 7;
    int r = (int)(row & 127);
    int kb = col8 >> 3, c = col8 & 7;
    return ((size_t)t * (size_t)(ld >> 6) + (size_t)kb) * (size_t)(kTileBytes << split) + (size_t)(r >> 3) * 1024 + (size_t)(r & 7) * 128 +
           (size_t)((c ^ (r & 7)) << 4);
}

#ifdef __CUDACC__
// {low half = bf16(a), high half = bf16(b)}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}
// split-bf16 pair of two floats: hi = bf16x2(a, b), lo = bf16x2(a - hi.a, b - hi.b); a == hi + lo to ~2^-17 relative
__device__ __forceinline__ void split_bf16x2(float a, float b, uint32_t &hi, uint32_t &lo) {
    hi = pack_bf16x2(a, b);
    lo = pack_bf16x2(__fsub_rn(a, __uint_as_float(hi << 16)), __fsub_rn(b, __uint_as_float(hi & 0xffff0000u)));
}
#endif

}  // namespace gspn
