// fps_common.cuh -- pieces shared by the farthest-point-sampling kernels (fps.cu, fps_pruned.cu): the reference's tie-break as a
// sortable key, the warp argmax, and the DSMEM all-to-all primitives (st.async + mbarrier complete_tx).
#pragma once
#include <cooperative_groups.h>
#include "common.cuh"

namespace gspn {

__device__ __forceinline__ unsigned fps_key(int k) { return ((unsigned)(k & 511) << 23) | ((unsigned)k >> 9); }
__device__ __forceinline__ int fps_unkey(unsigned key) { return (int)(((key & 0x7FFFFFu) << 9) | (key >> 23)); }

struct Cand {  // a candidate: squared distance bits, tie-break key, coordinates
    int dbits;
    unsigned key;
    float x, y, z;
};

// warp-wide (max dist, then min key); every lane returns the winner's fields.
__device__ __forceinline__ Cand warp_argmax(Cand c) {
    int wm = __reduce_max_sync(GSPN_FULL_MASK, c.dbits);  // non-negative floats order as ints; -1.0f (empty) is negative
    unsigned kk = (c.dbits == wm) ? c.key : 0xFFFFFFFFu;
    unsigned wk = __reduce_min_sync(GSPN_FULL_MASK, kk);
    int src = __ffs(__ballot_sync(GSPN_FULL_MASK, kk == wk)) - 1;
    Cand r;
    r.dbits = wm;
    r.key = wk;
    r.x = __shfl_sync(GSPN_FULL_MASK, c.x, src);
    r.y = __shfl_sync(GSPN_FULL_MASK, c.y, src);
    r.z = __shfl_sync(GSPN_FULL_MASK, c.z, src);
    return r;
}

__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t f_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
// shared::cta address -> shared::cluster address of the same variable in CTA `rank`
__device__ __forceinline__ uint32_t map_to_rank(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
// asynchronous DSMEM stores that signal the destination CTA's mbarrier (complete_tx): no cluster barrier,
// no gpu-scope fence on the round's critical path
__device__ __forceinline__ void st_async_v4(uint32_t raddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(raddr), "r"(a),
                 "r"(b), "r"(c), "r"(d), "r"(rbar)
                 : "memory");
}
__device__ __forceinline__ void st_async_b32(uint32_t raddr, uint32_t a, uint32_t rbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];" ::"r"(raddr), "r"(a), "r"(rbar) : "memory");
}
__device__ __forceinline__ void f_mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void f_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void f_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "FW_%=:\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra FD_%=;\n\t"
        "bra FW_%=;\n\t"
        "FD_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

// packed fp32x2 arithmetic (sm_100 FADD2 / FMUL2 / FFMA2): two points per instruction, each half IEEE round-to-nearest,
// so the per-point result is bit-identical to sqdist_fma
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ unsigned long long f2_sub(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long f2_mul(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

constexpr int kMaxWarps = 32;
constexpr int kMaxCand = 128;  // CLUSTER * warps-per-CTA candidates per round in the all-to-all exchange

struct __align__(16) Slot { float x, y, z; int dbits; };

}  // namespace gspn
