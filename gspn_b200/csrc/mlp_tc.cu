// mlp_tc.cu -- the shared per-point MLP (+ max-pool over nsample) of pointnet_sa_module /
// pointnet_fp_module (utils/pointnet_util.py:109-113,124,167-172; tf_util.conv2d 1x1 + bias + BN + ReLU,
// utils/tf_util.py:170-184,530-534) as ONE tcgen05 kernel per module: the whole layer chain runs on a
// 128-row tile without the activations ever leaving the SM.
//
//   * A operand (activations): 128 rows x K bf16, K-major, SWIZZLE_128B.  Layer 0 reads the tile image
//     written by the fused ball-query+group kernel (or gspn_fp_assemble) -- each 128x64 block is one
//     16 KiB cp.async.bulk (TMA bulk engine), no tensor map -- or, for rows of <= 8 columns, builds the
//     tile in shared memory itself from the ball-query indices (gather mode).  Layers >0 read what the
//     previous layer's epilogue wrote into the ONE activation region, in place (the epilogue of layer l only
//     starts when every MMA of layer l has completed, so it may overwrite that layer's operand).
//   * B operand (weights): W^T as [cout x cin] bf16 K-major blocks, pre-swizzled once by
//     gspn_mlp_pack_weights, streamed through a shared-memory ring by cp.async.bulk with
//     mbarrier complete_tx; the ring runs ahead across layers and tiles.
//   * D accumulates in TMEM (fp32, 128 lanes x cout columns); tcgen05.mma kind::f16, M=128,
//     N<=128 per instruction; tcgen05.commit releases ring stages and signals the epilogue.
//   * warp roles: 4 (two CTAs per SM) or 8 (one CTA per SM) epilogue warps, an input-producer warp, a
//     weight-producer warp (both back off with nanosleep: their waits are not latency-critical) and a CONVERGED
//     MMA-issuer warp (elect.sync around a block's four tcgen05.mma + commit), handing tiles back and forth through
//     the mma_done / epi_done mbarriers.
//   * epilogue: software-pipelined tcgen05.ld 32x32b.x32 -> fp32 scale/shift (bias+BN folded) -> ReLU fused into
//     the bf16 conversion (cvt.rn.relu.bf16x2) -> swizzled st.shared (next layer's A); last layer: max over the
//     nsample rows of a group by recursive-halving shuffles, or (pool == 1) per-warp swizzled staging boxes handed
//     to TMA tensor stores (cp.async.bulk.tensor.2d, full-line writes, rows beyond the tensor clipped by the map).
#include <cstdlib>
#include <cstring>
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link dependency)
#include "common.cuh"

namespace gspn {

constexpr int kMaxLayers = 4;
constexpr int kMaxStages = 4;

struct ChainParams {
    long rows;
    long ntiles;
    int nlayers;
    int K[kMaxLayers];  // padded input width (multiple of 64)
    int N[kMaxLayers];  // output width (multiple of 32)
    const unsigned char *wimg[kMaxLayers];
    const float *scale[kMaxLayers];
    const float *shift[kMaxLayers];
    int relu[kMaxLayers];
    const unsigned char *a;
    int pool;
    float *out_f32;
    int out_f32_vec;  // out_f32 is 32-byte aligned: 256-bit stores
    __nv_bfloat16 *out_bf16;
    int nch;  // weight rows (output channels) per ring stage / per MMA
    int tmem_cols;
    int tm_bufs, slot_cols;  // accumulator buffers in TMEM (2: layer-steps alternate, so a tile's first layer is issued while
                             // the previous tile's last epilogue still drains the other buffer) and columns per buffer
    int epi_warps;           // 4, or 8 (two warps per TMEM lane quadrant, each taking half the columns)
    int a_stages, w_stages;  // ring depths: layer-0 input blocks (16 KiB each) / weight blocks (stage_bytes each)
    uint32_t r_bytes, stage_bytes;  // r_bytes: the activation region (hidden layers are written in place, see below)
    uint32_t affine_off;            // byte offset of the folded scale/shift table
    int tma_out;                    // pool == 1: output rows leave through per-warp swizzled staging + TMA tensor stores
    uint32_t stage_off;             // byte offset of the staging area (kStageWarpBytes per epilogue warp)
    int stage_alias;                // staging aliases the activation region (dead while the last layer's epilogue runs)
    long long *prof;  // optional: per-phase cycle counters of CTA 0 / thread 0 (tools/tc_profile.py)
    // gather mode (a == nullptr): layer 0's operand rows [features(c) | xyz - centre - shift | 0] (c+3 <= 8, K0 = 64) are built
    // in shared memory by the input-producer warp straight from the ball-query indices -- no tile image in HBM at all
    const int *g_idx;      // (b, m, nsample)
    const float *g_xyz;    // (b, n, 3)
    const float *g_ctr;    // (b, m, 3)
    const float *g_shift;  // (b, m, 3) or null
    const float *g_pts;    // (b, n, c) f32 or null
    int g_n, g_m, g_k, g_c;
    FastDiv g_div_k, g_div_m;  // row / nsample, query / m (rows < 2^31)
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t s_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mb_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra D_%=;\n\t"
        "bra W_%=;\n\t"
        "D_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// Producer-side wait (ring slot free): not latency-critical, so back off between polls instead of competing with the
// epilogue warps of the same SM sub-partition for issue slots (try_wait returns within a few cycles when it fails).
__device__ __forceinline__ void mb_wait_relaxed(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    for (;;) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) break;
        __nanosleep(100);
    }
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}
// TMA tensor store of one (32 rows x 32 columns) box from swizzled shared memory; coordinates are (column, row) elements
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *tm, uint32_t src, int c0, int r0) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tm), "r"(src), "r"(c0), "r"(r0)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
constexpr int kStageWarpBytes = 6144;  // per epilogue warp: 32 rows x 128 B (f32, SWIZZLE_128B) + 32 rows x 64 B (bf16, SWIZZLE_64B)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, bf16 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
template <int CW>
__device__ __forceinline__ void tc_ld(uint32_t taddr, uint32_t (&v)[CW]) {
    if constexpr (CW == 32) tc_ld32(taddr, v);
    else tc_ld16(taddr, v);
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14),
// LBO>>4 [16,30) (unused for swizzled K-major, 1), SBO>>4 = 1024>>4 [32,46), version 1 [46,48), layout 2 [61,64).
// Held as two 32-bit words: everything but the 14-bit start address is constant, so the issuer steps through
// ring stages and k-slices (+32 bytes = +2) with 32-bit adds on the low word only (no carry: shared memory ends below 2^18).
constexpr uint32_t kDescHi = 0x40004040u;  // SBO = 64 (bits 32-45), version 1 (bit 46), SWIZZLE_128B (bits 61-63)
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ uint64_t desc64(uint32_t lo) { return ((uint64_t)kDescHi << 32) | lo; }
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1,
// A,B K-major (bits 15,16 = 0), N>>3 [17,23), M>>4 [24,29).
__device__ __forceinline__ uint32_t instr_desc(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}
// {lo = bf16(max(a,0)), hi = bf16(max(b,0))} in one instruction (F2FP.RELU)
__device__ __forceinline__ uint32_t pack2_relu(float a, float b) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
    return d;
}
// one 256-bit global store (STG.E.256): 8 consecutive floats, 32-byte aligned
__device__ __forceinline__ void st_global_v8(float *dst, const float *v) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
                 "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}

struct Cursor {  // position in the per-tile weight block sequence: layer, k-block, n-chunk (fastest)
    int l, kb, nc;
    __device__ __forceinline__ void advance(const ChainParams &p) {
        if (++nc == (p.N[l] + p.nch - 1) / p.nch) {
            nc = 0;
            if (++kb == (p.K[l] >> 6)) {
                kb = 0;
                if (++l == p.nlayers) l = 0;
            }
        }
    }
};

// ---------------------------------------------------------------- epilogue (shared by both chain kernels)
struct EpiCtx {
    const float *sc, *sh;  // this layer's folded scale / shift (shared memory)
    unsigned char *outb;   // activation region the next layer reads (mid layers)
    unsigned char *stg;    // this warp's output staging boxes (TMA output path)
    int Nl, row, lane;     // layer width; row of the tile this thread owns (= TMEM lane); lane
    long grow, row0;       // global row of this thread; global row of the warp's first row
    bool last, relu;
};

// one step: CW accumulator columns of one row per thread -> affine (+ReLU) -> next layer's operand / output / max-pool
template <int CW>
__device__ __forceinline__ void epi_chunk(const ChainParams &p, const CUtensorMap *tm_f32, const CUtensorMap *tm_bf16, const EpiCtx &e,
                                          const uint32_t (&v)[CW], const int c0) {
    const int lane = e.lane, row = e.row, Nl = e.Nl;
    float f[CW];
    {
        const float4 *sc4 = reinterpret_cast<const float4 *>(e.sc + c0), *sh4 = reinterpret_cast<const float4 *>(e.sh + c0);
#pragma unroll
        for (int g = 0; g < CW / 4; ++g) {
            const float4 a4 = sc4[g], b4 = sh4[g];  // same address in every lane: one broadcast LDS.128 each
            f[4 * g] = fmaf(__uint_as_float(v[4 * g]), a4.x, b4.x); f[4 * g + 1] = fmaf(__uint_as_float(v[4 * g + 1]), a4.y, b4.y);
            f[4 * g + 2] = fmaf(__uint_as_float(v[4 * g + 2]), a4.z, b4.z); f[4 * g + 3] = fmaf(__uint_as_float(v[4 * g + 3]), a4.w, b4.w);
        }
    }
    if (!e.last) {
        // ReLU rides on the bf16 conversion (cvt.rn.relu.bf16x2.f32): no separate max per element
        unsigned char *dst = e.outb + (size_t)(c0 >> 6) * kTileBytes + (row >> 3) * 1024 + (row & 7) * 128;
        const int cc0 = (c0 >> 3) & 7;  // first 16-byte chunk of this step inside the 64-column block
        if (e.relu) {
#pragma unroll
            for (int g = 0; g < CW / 8; ++g) {
                uint4 pk = make_uint4(pack2_relu(f[8 * g], f[8 * g + 1]), pack2_relu(f[8 * g + 2], f[8 * g + 3]),
                                      pack2_relu(f[8 * g + 4], f[8 * g + 5]), pack2_relu(f[8 * g + 6], f[8 * g + 7]));
                *reinterpret_cast<uint4 *>(dst + (((cc0 + g) ^ (row & 7)) << 4)) = pk;
            }
        } else {
#pragma unroll
            for (int g = 0; g < CW / 8; ++g) {
                uint4 pk = make_uint4(pack2(f[8 * g], f[8 * g + 1]), pack2(f[8 * g + 2], f[8 * g + 3]), pack2(f[8 * g + 4], f[8 * g + 5]),
                                      pack2(f[8 * g + 6], f[8 * g + 7]));
                *reinterpret_cast<uint4 *>(dst + (((cc0 + g) ^ (row & 7)) << 4)) = pk;
            }
        }
        return;
    }
    if (e.relu) {
#pragma unroll
        for (int i = 0; i < CW; ++i) f[i] = fmaxf(f[i], 0.f);
    }
    if (p.pool == 1) {
        if (CW == 32 && p.tma_out) {
            // coalesced output without LSU pressure: the warp stages its 32 x 32 block in swizzled shared memory
            // (conflict-free 16-byte stores) and one lane hands it to the TMA store engine, which writes full
            // lines and clips rows beyond the tensor.  Direct row-per-lane stores cost one L1 wavefront per lane.
            unsigned char *stg = e.stg;
            if (lane == 0) bulk_wait_read0();  // the previous box of this warp has left shared memory
            __syncwarp();
            if (p.out_f32) {
#pragma unroll
                for (int g = 0; g < CW / 4; ++g)
                    *reinterpret_cast<float4 *>(stg + lane * 128 + ((g ^ (lane & 7)) << 4)) =
                        make_float4(f[4 * g], f[4 * g + 1], f[4 * g + 2], f[4 * g + 3]);
            }
            if (p.out_bf16) {
#pragma unroll
                for (int g = 0; g < CW / 8; ++g)
                    *reinterpret_cast<uint4 *>(stg + 4096 + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4)) =
                        make_uint4(pack2(f[8 * g], f[8 * g + 1]), pack2(f[8 * g + 2], f[8 * g + 3]), pack2(f[8 * g + 4], f[8 * g + 5]),
                                   pack2(f[8 * g + 6], f[8 * g + 7]));
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                if (p.out_f32) tma_store_2d(tm_f32, s_u32(stg), c0, (int)e.row0);
                if (p.out_bf16) tma_store_2d(tm_bf16, s_u32(stg + 4096), c0, (int)e.row0);
                bulk_commit();
            }
        } else if (e.grow < p.rows) {
            // thread = row: its CW columns are contiguous bytes of the output row; 256-bit stores (one 32-byte sector per lane)
            if (p.out_f32) {
                float *o = p.out_f32 + e.grow * Nl + c0;
                if (p.out_f32_vec) {
#pragma unroll
                    for (int g = 0; g < CW / 8; ++g) st_global_v8(o + 8 * g, &f[8 * g]);
                } else {
#pragma unroll
                    for (int i = 0; i < CW; ++i) o[i] = f[i];
                }
            }
            if (p.out_bf16) {
                uint4 *o = reinterpret_cast<uint4 *>(p.out_bf16 + e.grow * Nl + c0);
#pragma unroll
                for (int g = 0; g < CW / 8; ++g)
                    o[g] = make_uint4(pack2(f[8 * g], f[8 * g + 1]), pack2(f[8 * g + 2], f[8 * g + 3]), pack2(f[8 * g + 4], f[8 * g + 5]),
                                      pack2(f[8 * g + 6], f[8 * g + 7]));
            }
        }
    } else {
        // max over the 32 rows this warp holds, for CW columns at once: recursive halving -- at the step of lane
        // bit h a lane keeps the half of its columns selected by that bit and takes the partner's values for them
        // (CW-1 SHFL + FMNMX instead of CW warp-wide redux); lane L ends with the max of column L mod CW
#pragma unroll
        for (int h = CW / 2; h >= 1; h >>= 1) {
            const bool up = lane & h;
#pragma unroll
            for (int i = 0; i < h; ++i) {
                const float send = up ? f[i] : f[i + h], keepv = up ? f[i + h] : f[i];
                f[i] = fmaxf(keepv, __shfl_xor_sync(GSPN_FULL_MASK, send, h));
            }
        }
        float pooled = f[0];
        if (CW == 16) pooled = fmaxf(pooled, __shfl_xor_sync(GSPN_FULL_MASK, pooled, 16));  // the two 16-row halves
        const int keep = __float_as_int(pooled);
        if (e.row0 < p.rows && lane < CW) {
            const long grp = e.row0 / p.pool;
            const int col = c0 + (lane & (CW - 1));
            if (p.pool == 32) {
                if (p.out_f32) p.out_f32[grp * Nl + col] = __int_as_float(keep);
                if (p.out_bf16) p.out_bf16[grp * Nl + col] = __float2bfloat16_rn(__int_as_float(keep));
            } else {
                atomicMax(reinterpret_cast<int *>(p.out_f32) + grp * Nl + col, keep);  // out zeroed by the launcher
            }
        }
    }
}

// columns [c_lo, c_hi) of this warp's 32 rows: software-pipelined TMEM reads (the load of the next CW columns is in flight
// while this one is processed)
template <int CW>
__device__ __forceinline__ void epi_columns(const ChainParams &p, const CUtensorMap *tm_f32, const CUtensorMap *tm_bf16, const EpiCtx &e,
                                            const uint32_t tbase, const int c_lo, const int c_hi) {
    uint32_t va[CW], vb[CW];
    if (c_lo < c_hi) tc_ld<CW>(tbase + c_lo, va);
    for (int c0 = c_lo; c0 < c_hi; c0 += 2 * CW) {
        tc_wait_ld();
        if (c0 + CW < c_hi) tc_ld<CW>(tbase + c0 + CW, vb);
        epi_chunk<CW>(p, tm_f32, tm_bf16, e, va, c0);
        if (c0 + CW < c_hi) {
            tc_wait_ld();
            if (c0 + 2 * CW < c_hi) tc_ld<CW>(tbase + c0 + 2 * CW, va);
            epi_chunk<CW>(p, tm_f32, tm_bf16, e, vb, c0 + CW);
        }
    }
    // a hidden layer whose width is 32 mod 64 leaves the upper half of its last 64-column block to the K padding of
    // the next layer: keep it zero (the region is reused in place, an earlier, wider layer may have written there)
    if (!e.last && (e.Nl & 63) && c_lo == 0) {
        const uint4 z = make_uint4(0, 0, 0, 0);
        const int row = e.row;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const int cc = (e.Nl >> 3) + g;
            *reinterpret_cast<uint4 *>(e.outb + (size_t)(cc >> 3) * kTileBytes + (row >> 3) * 1024 + (row & 7) * 128 +
                                       (((cc & 7) ^ (row & 7)) << 4)) = z;
        }
    }
}

// EPI epilogue warps (4: one per TMEM lane quadrant, 8: two per quadrant, each taking half the columns); MINB CTAs per SM
// (register budget); CW columns per TMEM read (32 in both instantiations; a 16-column, <= 93-register variant was measured no faster)
template <int EPI, int MINB, int CW>
__global__ void __launch_bounds__(EPI * 32 + 96, MINB) mlp_chain_kernel(const ChainParams p, const __grid_constant__ CUtensorMap tm_f32,
                                                                        const __grid_constant__ CUtensorMap tm_bf16) {
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bars[4 * kMaxStages + 4];
    __shared__ int4 ltab[kMaxLayers];  // per layer, for the issuer: k-blocks, n-chunks, instruction descriptors (full / last chunk)

    // warp index through a shuffle: provably warp-uniform, so the role branches below are uniform branches and the
    // MMA issuer's operands can live in uniform registers (no per-instruction R2UR waterfall)
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(GSPN_FULL_MASK, tid >> 5, 0);
    const uint32_t raw = s_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B atoms are 1024-byte aligned
    unsigned char *sm = smem_raw + (base - raw);
    // ONE activation region, written in place: layer l>0 reads it as its A operand, and the epilogue of layer l only
    // starts after every MMA of layer l has completed (mma_done), so it may overwrite that operand with layer l's output.
    const uint32_t Rs = base;
    const uint32_t aring = base + p.r_bytes;
    const uint32_t wring = aring + (uint32_t)p.a_stages * kTileBytes;
    float *affine = reinterpret_cast<float *>(sm + p.affine_off);
    const uint32_t w_full = s_u32(&bars[0]), w_empty = s_u32(&bars[kMaxStages]), a_full = s_u32(&bars[2 * kMaxStages]),
                   a_empty = s_u32(&bars[3 * kMaxStages]), mma_done = s_u32(&bars[4 * kMaxStages]),
                   epi_done = s_u32(&bars[4 * kMaxStages + 2]);  // [2] each: one per TMEM accumulator buffer

    if (tid < p.nlayers) {
        const int nchunks = (p.N[tid] + p.nch - 1) / p.nch;
        ltab[tid] = make_int4(p.K[tid] >> 6, nchunks, (int)instr_desc(128, p.nch), (int)instr_desc(128, p.N[tid] - (nchunks - 1) * p.nch));
    }
    if (tid == 0) {
        for (int i = 0; i < 4 * kMaxStages + 2; ++i) mb_init(s_u32(&bars[i]), 1);
        mb_init(epi_done, p.epi_warps);  // one elected lane per epilogue warp arrives once per layer-step
        mb_init(epi_done + 8, p.epi_warps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // folded bias/BN affine of every layer -> smem ([scale_l | shift_l] per layer)
    {
        int o = 0;
        for (int l = 0; l < p.nlayers; ++l) {
            for (int i = tid; i < p.N[l]; i += blockDim.x) {
                affine[o + i] = __ldg(p.scale[l] + i);
                affine[o + p.N[l] + i] = __ldg(p.shift[l] + i);
            }
            o += 2 * p.N[l];
        }
    }
    // zero both activation regions once: K padding columns must read as 0 for every tile
    {
        uint4 z = make_uint4(0, 0, 0, 0);
        uint4 *r = reinterpret_cast<uint4 *>(sm);
        // gather mode also zeroes the input ring: each row only ever writes its first 16-byte chunk
        const uint32_t zbytes = p.r_bytes + (p.a ? 0u : (uint32_t)p.a_stages * kTileBytes);
        for (uint32_t i = tid; i < zbytes / 16; i += blockDim.x) r[i] = z;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_u32(&tmem_slot)), "r"(p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    int blocks_per_tile = 0;
    for (int l = 0; l < p.nlayers; ++l) blocks_per_tile += ((p.N[l] + p.nch - 1) / p.nch) * (p.K[l] >> 6);
    const int kb0 = p.K[0] >> 6;
    const long my_tiles = (p.ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;
    const int total_w = (int)(my_tiles * blocks_per_tile), total_a = (int)(my_tiles * kb0);

    // ---- producer warps: one lane each keeps a ring full for the whole kernel, independent of the MMA/epilogue
    // timeline, so the loads of the next layers / tiles are in flight while the epilogue warps are busy
    if (warp == p.epi_warps && p.a == nullptr) {
        // gather producer: the whole warp; lane L builds rows L, L+32, L+64, L+96 of the tile (one 16-byte chunk each).
        // Memory-level parallelism is what this warp lives on: the four rows' indices are loaded first (and the NEXT tile's
        // are requested before this tile is built), then all dependent coordinate / feature gathers are issued together,
        // and only then is anything consumed.  Row -> (query, cloud) uses the multiply-shift divider (rows < 2^31).
        int s = 0, par = 0;
        const int c = p.g_c;
        long t = blockIdx.x;
        int ii[4], nxt[4];
        auto load_idx = [&](long tile, int (&dst)[4]) {
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const long row = tile * kTileRows + lane + 32 * rr;
                dst[rr] = (tile < p.ntiles && row < p.rows) ? __ldg(p.g_idx + row) : -1;
            }
        };
        load_idx(t, nxt);
        for (int i = 0; i < total_a; ++i, t += gridDim.x) {
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) ii[rr] = nxt[rr];
            load_idx(t + gridDim.x, nxt);
            float gx[4][3], cx[4][3], sh[4][3], f[4][5];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const long row = t * kTileRows + lane + 32 * rr;
                const bool ok = ii[rr] >= 0;
                const uint32_t q = ok ? p.g_div_k.div((uint32_t)row) : 0u;  // global query index cloud*m + j
                const uint32_t cloud = p.g_div_m.div(q);
                const size_t pt = (size_t)cloud * p.g_n + (ok ? ii[rr] : 0);
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    gx[rr][a] = ok ? __ldg(p.g_xyz + pt * 3 + a) : 0.f;
                    cx[rr][a] = ok ? __ldg(p.g_ctr + (size_t)q * 3 + a) : 0.f;
                    sh[rr][a] = (ok && p.g_shift) ? __ldg(p.g_shift + (size_t)q * 3 + a) : 0.f;
                }
#pragma unroll
                for (int a = 0; a < 5; ++a) f[rr][a] = (ok && a < c) ? __ldg(p.g_pts + pt * c + a) : 0.f;
            }
            if (i >= p.a_stages) mb_wait_relaxed(a_empty + 8 * s, (uint32_t)(par ^ 1));
            unsigned char *stage = sm + p.r_bytes + (size_t)s * kTileBytes;
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const int r = lane + 32 * rr;
                float d[3], v[8];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    d[a] = __fsub_rn(gx[rr][a], cx[rr][a]);             // grouped_xyz -= new_xyz (pointnet_util.py:42)
                    if (p.g_shift) d[a] = __fsub_rn(d[a], sh[rr][a]);   // -= shift_pred (model_rpointnet.py:56-57)
                }
#pragma unroll
                for (int t2 = 0; t2 < 8; ++t2)  // columns [features(c) | dx dy dz | 0]; c is a runtime value <= 5
                    v[t2] = (t2 < c) ? f[rr][t2 < 5 ? t2 : 4] : (t2 == c ? d[0] : (t2 == c + 1 ? d[1] : (t2 == c + 2 ? d[2] : 0.f)));
                if (ii[rr] < 0) {
#pragma unroll
                    for (int t2 = 0; t2 < 8; ++t2) v[t2] = 0.f;  // rows past the end of the problem
                }
                uint4 pk = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
                *reinterpret_cast<uint4 *>(stage + (r >> 3) * 1024 + (r & 7) * 128 + ((r & 7) << 4)) = pk;  // chunk 0 ^ (r & 7)
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(a_full + 8 * s) : "memory");
            if (++s == p.a_stages) { s = 0; par ^= 1; }
        }
    } else if (warp == p.epi_warps) {
        if (lane == 0) {
            int s = 0, par = 0, kb = 0;
            long t = blockIdx.x;
            for (int i = 0; i < total_a; ++i) {
                if (i >= p.a_stages) mb_wait_relaxed(a_empty + 8 * s, (uint32_t)(par ^ 1));
                mb_expect_tx(a_full + 8 * s, kTileBytes);
                bulk_load(aring + s * kTileBytes, p.a + ((size_t)t * kb0 + kb) * kTileBytes, kTileBytes, a_full + 8 * s);
                if (++kb == kb0) { kb = 0; t += gridDim.x; }
                if (++s == p.a_stages) { s = 0; par ^= 1; }
            }
        }
    } else if (warp == p.epi_warps + 1) {
        if (lane == 0) {
            int s = 0, par = 0;
            Cursor wc = {0, 0, 0};
            for (int i = 0; i < total_w; ++i) {
                if (i >= p.w_stages) mb_wait_relaxed(w_empty + 8 * s, (uint32_t)(par ^ 1));
                const int rows_i = min(p.nch, p.N[wc.l] - wc.nc * p.nch);
                const uint32_t bytes = (uint32_t)rows_i * 128u;
                mb_expect_tx(w_full + 8 * s, bytes);
                bulk_load(wring + s * p.stage_bytes, p.wimg[wc.l] + (size_t)wc.kb * p.N[wc.l] * 128 + (size_t)wc.nc * p.nch * 128, bytes,
                          w_full + 8 * s);
                wc.advance(p);
                if (++s == p.w_stages) { s = 0; par ^= 1; }
            }
        }
    } else if (warp == p.epi_warps + 2) {
        // ---- MMA issuer warp.  The whole warp runs the loop CONVERGED (barrier waits, ring bookkeeping and descriptor
        // arithmetic stay warp-uniform, so they live in uniform registers next to the UTCHMMA operands); only the
        // tcgen05.mma / tcgen05.commit instructions themselves are executed by one lane.  Operand waits run ahead of the
        // epilogue warps; the layer-to-layer critical path is  wait(epi_done) -> tcgen05.mma ... -> commit(mma_done)
        {
            int wu_s = 0, wu_par = 0, au_s = 0, au_par = 0;  // consumer-side stage index and round parity
            // Layer-steps alternate between the TMEM accumulator buffers (tm_bufs == 2).  Step s may be issued when
            //   (1) its buffer is free: the epilogue of step s - tm_bufs has drained it, and
            //   (2) for layers > 0, its A operand is written: the epilogue of step s - 1 is done.
            // A tile's FIRST layer reads the input ring, so with two buffers it is issued while the previous tile's last
            // epilogue is still running: that layer's MMA time disappears from the critical path.
            const int NB = p.tm_bufs;
            int seen0 = 0, seen1 = 0;  // completed epi_done phases already observed, per buffer
            int step = 0;
            auto wait_epi = [&](int b, int phase) {  // phases of one barrier complete, and are waited for, in order
                if ((b ? seen1 : seen0) > phase) return;
                mb_wait(epi_done + 8 * b, (uint32_t)(phase & 1));
                if (b) seen1 = phase + 1; else seen0 = phase + 1;
            };
            const uint32_t a_lo0 = desc_lo(aring), w_lo0 = desc_lo(wring), r_lo0 = desc_lo(Rs);
            const uint32_t w_step = p.stage_bytes >> 4;
            for (long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
                for (int l = 0; l < p.nlayers; ++l, ++step) {
                    const int buf = NB == 2 ? (step & 1) : 0;
                    const int ph = NB == 2 ? (step >> 1) : step;  // this step's phase on its buffer's barriers
                    const int4 L = ltab[l];
                    const int KBl = L.x, nchunks = L.y;
                    // every operand wait that can be satisfied from what the rings already hold is done BEFORE the
                    // epilogue hand-off, so that after it the loop is  tcgen05.mma x4 + commit  per block
                    long long mw0 = 0;
                    if (p.prof) mw0 = clock64();
                    const int nblk = KBl * nchunks;
                    const int pre_w = nblk < p.w_stages ? nblk : p.w_stages;
                    const int pre_a = (l == 0) ? (KBl < p.a_stages ? KBl : p.a_stages) : 0;
                    for (int i = 0, st = au_s, pr = au_par; i < pre_a; ++i) {
                        mb_wait(a_full + 8 * st, (uint32_t)pr);
                        if (++st == p.a_stages) { st = 0; pr ^= 1; }
                    }
                    for (int i = 0, st = wu_s, pr = wu_par; i < pre_w; ++i) {
                        mb_wait(w_full + 8 * st, (uint32_t)pr);
                        if (++st == p.w_stages) { st = 0; pr ^= 1; }
                    }
                    const uint32_t tm = tmem + buf * p.slot_cols;
                    if (ph >= 1) wait_epi(buf, ph - 1);                               // (1) TMEM buffer drained
                    if (l > 0 && NB == 2) wait_epi(buf ^ 1, (step - 1) >> 1);         // (2) A operand written
                    long long mt0 = 0;
                    if (p.prof) {
                        mt0 = clock64();  // slot 7: operand pre-waits + epilogue hand-off, as seen by the issuer
                        if (blockIdx.x == 0 && lane == 0) atomicAdd((unsigned long long *)p.prof + 7, (unsigned long long)(mt0 - mw0));
                    }
                    tc_fence_after();
                    int blk = 0;
                    for (int kb = 0; kb < KBl; ++kb) {
                        uint32_t a_lo;
                        if (l == 0) {
                            if (kb >= pre_a) { mb_wait(a_full + 8 * au_s, (uint32_t)au_par); tc_fence_after(); }
                            a_lo = a_lo0 + (uint32_t)au_s * (kTileBytes >> 4);
                        } else {
                            a_lo = r_lo0 + (uint32_t)kb * (kTileBytes >> 4);
                        }
                        for (int nc = 0; nc < nchunks; ++nc, ++blk) {
                            const int s = wu_s;
                            if (blk >= pre_w) { mb_wait(w_full + 8 * s, (uint32_t)wu_par); tc_fence_after(); }
                            const uint32_t idesc = (uint32_t)((nc == nchunks - 1) ? L.w : L.z);
                            const uint32_t b_lo = w_lo0 + (uint32_t)s * w_step;
                            if (elect_one()) {
#pragma unroll
                                for (int k = 0; k < 4; ++k)  // 4 x UMMA_K(16) per 64-wide block: +32 bytes = +2 in the descriptor
                                    tc_mma(tm + nc * p.nch, desc64(a_lo + 2 * k), desc64(b_lo + 2 * k), idesc, (kb | k) != 0);
                                tc_commit(w_empty + 8 * s);  // stage free once these MMAs have read it
                            }
                            __syncwarp();
                            if (++wu_s == p.w_stages) { wu_s = 0; wu_par ^= 1; }
                        }
                        if (l == 0) {
                            if (elect_one()) tc_commit(a_empty + 8 * au_s);
                            if (++au_s == p.a_stages) { au_s = 0; au_par ^= 1; }
                        }
                    }
                    if (elect_one()) tc_commit(mma_done + 8 * buf);
                    __syncwarp();
                    if (p.prof && blockIdx.x == 0 && lane == 0) {  // profiling only: how long issuing takes, and how long the MMAs take to drain
                        long long mt1 = clock64();
                        mb_wait(mma_done + 8 * buf, (uint32_t)(ph & 1));
                        long long mt2 = clock64();
                        atomicAdd((unsigned long long *)p.prof + 5, (unsigned long long)(mt1 - mt0));
                        atomicAdd((unsigned long long *)p.prof + 6, (unsigned long long)(mt2 - mt1));
                    }
                }
            }
        }
    } else {
    // ---- epilogue warps
    long estep = 0;  // layer-steps alternate between the TMEM accumulator buffers exactly as the issuer's do

    for (long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x) {
        int ao = 0;  // running offset of layer l's [scale | shift] in the affine table
        for (int l = 0; l < p.nlayers; ++l, ++estep) {
            const int Nl = p.N[l];
            const int buf = p.tm_bufs == 2 ? (int)(estep & 1) : 0;
            const long eph = p.tm_bufs == 2 ? (estep >> 1) : estep;
            long long pt0 = 0, pt1 = 0, pt2 = 0, pt3 = 0;
            if (p.prof) pt0 = pt1 = clock64();
            mb_wait(mma_done + 8 * buf, (uint32_t)(eph & 1));
            tc_fence_after();
            if (p.prof) pt2 = clock64();

            // ---- epilogue: thread = row (TMEM lane), CW columns at a time
            const bool last = (l == p.nlayers - 1);
            if (p.tma_out && p.stage_alias && l == 0 && !last) {
                // the staging boxes of the previous tile's output live in the activation region this epilogue is about to
                // write: every warp's pending TMA stores must have finished READING shared memory first
                if (lane == 0) bulk_wait_read0();
                asm volatile("bar.sync 1, %0;" ::"r"(EPI * 32) : "memory");
            }
            if (p.tma_out && p.stage_alias && last && l > 0) {
                // the other direction: this warp's staging boxes overlay operand bytes OTHER epilogue warps wrote one layer
                // earlier.  That is already ordered (their epi_done arrive -> issuer -> tcgen05.commit -> this wait), but only
                // through the tensor core's asynchronous arrive; a named barrier makes it a plain CTA-level ordering as well
                // (compute-sanitizer racecheck cannot follow the former).  ~60 cycles per tile.
                asm volatile("bar.sync 1, %0;" ::"r"(EPI * 32) : "memory");
            }
            const float *sc = affine + ao, *sh = sc + Nl;
            ao += 2 * Nl;
            const int quad = warp & 3, half = warp >> 2, nhalf = EPI >> 2;
            const int row = quad * 32 + lane;
            const long grow = tile * kTileRows + row;
            const int c_lo = ((Nl / CW) * half / nhalf) * CW, c_hi = ((Nl / CW) * (half + 1) / nhalf) * CW;  // this warp's columns
            unsigned char *outb = sm;
            const uint32_t tbase = tmem + buf * p.slot_cols + ((uint32_t)(quad * 32) << 16);
            EpiCtx ec;
            ec.sc = sc; ec.sh = sh; ec.outb = outb; ec.stg = sm + p.stage_off + warp * kStageWarpBytes;
            ec.Nl = Nl; ec.row = row; ec.lane = lane; ec.grow = grow; ec.row0 = tile * kTileRows + quad * 32;
            ec.last = last; ec.relu = p.relu[l] != 0;
            epi_columns<CW>(p, &tm_f32, &tm_bf16, ec, tbase, c_lo, c_hi);
            if (p.prof) pt3 = clock64();
            tc_fence_before();
            // epilogue st.shared -> visible to the tensor core's async-proxy reads.  The last layer wrote no operand, and its
            // fence (MEMBAR + proxy fence) would only wait for the output stores to drain before the tile is handed back
            if (!last) fence_proxy_async();
            __syncwarp();  // orders every lane's stores + proxy fence before the elected lane's arrive
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(epi_done + 8 * buf) : "memory");  // buffer back to the issuer
            if (p.prof && blockIdx.x == 0 && tid == 0) {
                long long pt4 = clock64();
                atomicAdd((unsigned long long *)p.prof + 0, (unsigned long long)(pt1 - pt0));  // issue (loads + MMAs)
                atomicAdd((unsigned long long *)p.prof + 1, (unsigned long long)(pt2 - pt1));  // wait for MMA completion
                atomicAdd((unsigned long long *)p.prof + 2, (unsigned long long)(pt3 - pt2));  // epilogue
                atomicAdd((unsigned long long *)p.prof + 3, (unsigned long long)(pt4 - pt3));  // fences + CTA barrier
                atomicAdd((unsigned long long *)p.prof + 4, 1ull);                              // layer-steps
            }
        }
    }
    if (p.tma_out && lane == 0) bulk_wait0();  // this warp's output boxes have been written
    }  // roles
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(p.tmem_cols) : "memory");
}

// ---- weights (cin,cout) f32 row-major -> [cin_padded/64] blocks of (cout x 64) bf16, K-major, 128B-swizzled
__global__ void pack_weights_kernel(int cin, int cin_padded, int cout, const float *__restrict__ w, const int *__restrict__ row_perm,
                                    unsigned char *__restrict__ img) {
    long total = (long)cin_padded * cout;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        int k = (int)(e / cout), n = (int)(e - (long)k * cout);
        int src = row_perm ? row_perm[k] : (k < cin ? k : -1);
        float v = (src >= 0 && src < cin) ? w[(size_t)src * cout + n] : 0.f;
        int kb = k >> 6, kk = k & 63;
        size_t off = (size_t)kb * cout * 128 + (size_t)(n >> 3) * 1024 + (n & 7) * 128 + ((((kk >> 3) ^ (n & 7))) << 4) + (kk & 7) * 2;
        *reinterpret_cast<__nv_bfloat16 *>(img + off) = __float2bfloat16_rn(v);
    }
}

// ---- FP front end: [three_interpolate(points2) | points1 | 0] -> bf16 tile image (one 16-byte chunk per thread)
__global__ void __launch_bounds__(256) fp_assemble_kernel(long rows, int n, int m, int c1, int c2, const float *__restrict__ points1,
                                                          const float *__restrict__ points2, const int *__restrict__ idx,
                                                          const float *__restrict__ weight, unsigned char *__restrict__ img, int ld,
                                                          const FastDiv div_chunks, const FastDiv div_n) {
    const int chunks = ld >> 3;
    const long total = rows * chunks;
    const bool small = total < (1L << 31);  // element indices fit the multiply-shift divider
    const bool vec2 = (c2 % 8 == 0) && ((reinterpret_cast<uintptr_t>(points2) & 15u) == 0);
    const bool vec1 = (c2 % 8 == 0) && (c1 % 4 == 0) && points1 && ((reinterpret_cast<uintptr_t>(points1) & 15u) == 0);
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        long row, bi;
        if (small) {
            row = div_chunks.div((uint32_t)e);
            bi = div_n.div((uint32_t)row);
        } else {
            row = e / chunks;
            bi = row / n;
        }
        int ch = (int)(e - row * chunks);
        float v[8];
        const int col0 = ch * 8;
        if (col0 + 8 <= c2 && vec2) {
            // whole chunk interpolated: 6 x 128-bit gathers; (p1*w1+p2*w2)+p3*w3, no FMA: tf_interpolate.cpp:107-127
            const int *ip = idx + row * 3;
            const float *wp = weight + row * 3;
            const float w1 = __ldg(wp), w2 = __ldg(wp + 1), w3 = __ldg(wp + 2);
            const float4 *b1 = reinterpret_cast<const float4 *>(points2 + (bi * m + __ldg(ip)) * c2 + col0);
            const float4 *b2 = reinterpret_cast<const float4 *>(points2 + (bi * m + __ldg(ip + 1)) * c2 + col0);
            const float4 *b3 = reinterpret_cast<const float4 *>(points2 + (bi * m + __ldg(ip + 2)) * c2 + col0);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float4 a = __ldg(b1 + h), b = __ldg(b2 + h), c = __ldg(b3 + h);
                v[4 * h] = __fadd_rn(__fadd_rn(__fmul_rn(a.x, w1), __fmul_rn(b.x, w2)), __fmul_rn(c.x, w3));
                v[4 * h + 1] = __fadd_rn(__fadd_rn(__fmul_rn(a.y, w1), __fmul_rn(b.y, w2)), __fmul_rn(c.y, w3));
                v[4 * h + 2] = __fadd_rn(__fadd_rn(__fmul_rn(a.z, w1), __fmul_rn(b.z, w2)), __fmul_rn(c.z, w3));
                v[4 * h + 3] = __fadd_rn(__fadd_rn(__fmul_rn(a.w, w1), __fmul_rn(b.w, w2)), __fmul_rn(c.w, w3));
            }
        } else if (col0 >= c2 && col0 - c2 + 8 <= c1 && vec1) {
            const float4 *q = reinterpret_cast<const float4 *>(points1 + row * c1 + (col0 - c2));
            const float4 a = __ldg(q), b = __ldg(q + 1);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else if (col0 < c2) {
            const int *ip = idx + row * 3;
            const float *wp = weight + row * 3;
            int i1 = __ldg(ip), i2 = __ldg(ip + 1), i3 = __ldg(ip + 2);
            float w1 = __ldg(wp), w2 = __ldg(wp + 1), w3 = __ldg(wp + 2);
            const float *b1 = points2 + (bi * m + i1) * c2, *b2 = points2 + (bi * m + i2) * c2, *b3 = points2 + (bi * m + i3) * c2;
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                int col = col0 + t;
                if (col < c2) {
                    v[t] = __fadd_rn(__fadd_rn(__fmul_rn(__ldg(b1 + col), w1), __fmul_rn(__ldg(b2 + col), w2)), __fmul_rn(__ldg(b3 + col), w3));
                } else {
                    int q = col - c2;
                    v[t] = q < c1 ? __ldg(points1 + row * c1 + q) : 0.f;
                }
            }
        } else {
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                int q = col0 + t - c2;
                v[t] = q < c1 ? __ldg(points1 + row * c1 + q) : 0.f;
            }
        }
        uint4 pk = make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
        *reinterpret_cast<uint4 *>(img + tile_chunk_offset(row, ch, ld)) = pk;
    }
}

}  // namespace gspn

using namespace gspn;

extern "C" size_t gspn_mlp_weight_image_bytes(int cin_padded, int cout) {
    if (cin_padded <= 0 || cout <= 0 || cin_padded % 64 || cout % 8) return 0;
    return (size_t)(cin_padded / 64) * (size_t)cout * 128;
}

extern "C" int gspn_mlp_pack_weights(int cin, int cin_padded, int cout, const float *w_f32, const int *row_perm, void *wimg,
                                     gspn_stream_t stream) {
    GSPN_REQUIRE(cin > 0 && cin_padded >= cin && cin_padded % 64 == 0 && cout > 0 && cout % 8 == 0);
    GSPN_REQUIRE_PTR(w_f32); GSPN_REQUIRE_PTR(wimg);
    long total = (long)cin_padded * cout;
    pack_weights_kernel<<<(unsigned)ceil_div_l(total, 256), 256, 0, as_stream(stream)>>>(cin, cin_padded, cout, w_f32, row_perm,
                                                                                        (unsigned char *)wimg);
    return check_launch();
}

// cuTensorMapEncodeTiled through the runtime's driver entry point query: no link-time dependency on libcuda, so the library still
// loads (and exports its symbols) on a machine without a driver
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = nullptr;
    static int tried = 0;  // benign race: the query is idempotent
    if (!tried) {
        void *sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
        else
            (void)cudaGetLastError();
        tried = 1;
    }
    return fn;
}
// (rows x n) row-major output as a 2-D tensor map with a 32 x 32 box: 128-byte (f32) or 64-byte (bf16) swizzled box rows
static bool encode_out_map(CUtensorMap *tm, void *base, long rows, int n, bool bf16) {
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc || (reinterpret_cast<uintptr_t>(base) & 15u) || rows <= 0 || rows > 0x7fffffffL) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)n * (bf16 ? 2u : 4u)};
    const cuuint32_t box[2] = {32u, 32u}, estr[2] = {1u, 1u};
    return enc(tm, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, bf16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static long long *g_chain_prof = nullptr;  // tuning door: set by gspn_mlp_chain_set_profile, read at launch
extern "C" void gspn_mlp_chain_set_profile(long long *prof5) { g_chain_prof = prof5; }

static int chain_launch(ChainParams p, long rows, int nlayers, const int *dims, const void *a, const void *const *wimg,
                        const float *const *scale, const float *const *shift, const int *relu, int pool, float *out_f32, void *out_bf16,
                        gspn_stream_t stream) {
    GSPN_REQUIRE(rows >= 0 && nlayers >= 1 && nlayers <= kMaxLayers && pool >= 1);
    GSPN_REQUIRE_PTR(dims); GSPN_REQUIRE_PTR(wimg); GSPN_REQUIRE_PTR(scale); GSPN_REQUIRE_PTR(shift); GSPN_REQUIRE_PTR(relu);
    if (rows == 0) return GSPN_OK;
    if (a == nullptr && p.g_idx == nullptr) return GSPN_E_NULL_PTR;
    if (out_f32 == nullptr && out_bf16 == nullptr) return GSPN_E_NULL_PTR;
    p.rows = rows;
    p.ntiles = ceil_div_l(rows, kTileRows);
    p.nlayers = nlayers;
    int maxn = 0;
    size_t affine_floats = 0;
    for (int l = 0; l < nlayers; ++l) {
        p.K[l] = l == 0 ? dims[0] : ((dims[l] + 63) / 64) * 64;
        p.N[l] = dims[l + 1];
        GSPN_REQUIRE(p.K[l] > 0 && p.K[l] % 64 == 0 && p.N[l] > 0);
        if (p.N[l] % 32 != 0 || p.N[l] > 512) return GSPN_E_UNSUPPORTED;  // use the fp32 path
        GSPN_REQUIRE_PTR(wimg[l]); GSPN_REQUIRE_PTR(scale[l]); GSPN_REQUIRE_PTR(shift[l]);
        p.wimg[l] = (const unsigned char *)wimg[l];
        p.scale[l] = scale[l];
        p.shift[l] = shift[l];
        p.relu[l] = relu[l];
        maxn = p.N[l] > maxn ? p.N[l] : maxn;
        affine_floats += 2 * (size_t)p.N[l];
    }
    if (pool > 1) {
        if (pool % 32 != 0 || rows % pool != 0 || !relu[nlayers - 1]) return GSPN_E_UNSUPPORTED;
        if (pool != 32 && (out_f32 == nullptr || out_bf16 != nullptr)) return GSPN_E_UNSUPPORTED;
    }
    p.a = (const unsigned char *)a;
    p.prof = g_chain_prof;
    p.pool = pool;
    p.out_f32 = out_f32;
    p.out_f32_vec = (reinterpret_cast<uintptr_t>(out_f32) & 31u) == 0;
    p.out_bf16 = (__nv_bfloat16 *)out_bf16;
    p.nch = maxn < 128 ? maxn : 128;
    p.tmem_cols = 32;
    while (p.tmem_cols < maxn) p.tmem_cols <<= 1;
    int sms = 148;
    {
        static int sms_cached = 0;
        if (sms_cached == 0) {
            int dev = 0, v = 148;
            GSPN_CUDA_OK(cudaGetDevice(&dev));
            GSPN_CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
            sms_cached = v;
        }
        sms = sms_cached;
    }
    // the activation region holds the widest hidden layer (hidden layers are written in place)
    int rk = 0;
    for (int l = 0; l + 1 < nlayers; ++l) {
        int kp = ((p.N[l] + 63) / 64) * 64;
        rk = kp > rk ? kp : rk;
    }
    p.r_bytes = (uint32_t)(rk / 64) * kTileBytes;
    p.stage_bytes = (uint32_t)p.nch * 128u;
    // pool == 1 outputs leave through TMA tensor stores from per-warp staging boxes (door: GSPN_TC_TMA_OUT=0 -> direct stores)
    CUtensorMap tm_f32, tm_bf16;
    memset(&tm_f32, 0, sizeof(tm_f32));
    memset(&tm_bf16, 0, sizeof(tm_bf16));
    p.tma_out = pool == 1;
    if (const char *e = getenv("GSPN_TC_TMA_OUT")) p.tma_out = p.tma_out && atoi(e) != 0;
    if (p.tma_out && out_f32) p.tma_out = encode_out_map(&tm_f32, out_f32, rows, p.N[nlayers - 1], false);
    if (p.tma_out && out_bf16) p.tma_out = encode_out_map(&tm_bf16, out_bf16, rows, p.N[nlayers - 1], true);
    cudaStream_t s = as_stream(stream);
    if (pool > 1 && pool != 32)
        GSPN_CUDA_OK(cudaMemsetAsync(out_f32, 0, sizeof(float) * (size_t)(rows / pool) * p.N[nlayers - 1], s));

    // Shared-memory plan: [activation region | input ring | weight ring | affine table].
    // Pick the deepest rings that still give the best CTA co-residency (a second CTA on the SM overlaps its MMAs with
    // this one's epilogue); TMEM (512 columns/SM) and the register file bound co-residency too.
    const int occ_tmem = 512 / p.tmem_cols;
    int occ_cap = 2;  // 128 registers x 224 threads: two CTAs (plus an FPS CTA of another lane) fit the register file
    if (const char *e = getenv("GSPN_TC_OCC")) {  // tuning door
        int v = atoi(e);
        if (v == 1 || v == 2) occ_cap = v;
    }
    const int tries[4][2] = {{3, 4}, {2, 4}, {2, 3}, {2, 2}};
    size_t smem = 0;
    int occ = 0;
    // first plan for two co-resident CTAs with 4 epilogue warps each, then for one CTA with 8 (staging is per epilogue warp)
    for (int epi = 4; epi <= 8 && occ < 2; epi += 4) {
        const size_t staging = p.tma_out ? (size_t)epi * kStageWarpBytes : 0;
        const bool alias = p.tma_out && nlayers > 1 && staging <= p.r_bytes;
        const size_t stage_extra = alias ? 0 : staging;
        for (int t = 0; t < 4; ++t) {
            const size_t rings = (size_t)p.r_bytes + (size_t)tries[t][0] * kTileBytes + (size_t)tries[t][1] * p.stage_bytes;
            const size_t sz = 1024 + rings + stage_extra + affine_floats * sizeof(float);
            if (sz > 226 * 1024) continue;
            int o = (int)((228 * 1024) / (sz + 1024));
            o = o > occ_tmem ? occ_tmem : o;
            o = o > occ_cap ? occ_cap : o;
            if (epi == 8) o = o > 1 ? 1 : o;
            if (epi == 4 && o < 2) continue;  // 4 epilogue warps only pay off with a second CTA on the SM
            if (o > occ) {
                occ = o; smem = sz; p.a_stages = tries[t][0]; p.w_stages = tries[t][1];
                p.stage_alias = alias;
                p.stage_off = alias ? 0u : (uint32_t)rings;  // rings end on a 1 KiB boundary (swizzle atoms)
                p.affine_off = (uint32_t)(rings + stage_extra);
            }
        }
    }
    // two CTAs per SM: 4 epilogue warps each; one CTA per SM: 8 epilogue warps (two per TMEM lane quadrant)
    p.epi_warps = occ >= 2 ? 4 : 8;
    if (occ < 1) return GSPN_E_UNSUPPORTED;
    // a second accumulator buffer when the SM's 512 TMEM columns allow it for every co-resident CTA (door: GSPN_TC_BUFS=1)
    p.slot_cols = p.tmem_cols;
    p.tm_bufs = (2 * p.slot_cols * occ <= 512) ? 2 : 1;
    if (const char *e = getenv("GSPN_TC_BUFS")) { if (atoi(e) == 1) p.tm_bufs = 1; }
    p.tmem_cols = p.tm_bufs * p.slot_cols;
    // never let more CTAs co-reside than TMEM can serve: inflate the request if shared memory alone would allow it
    const size_t min_smem = (size_t)(228 * 1024) / (occ + 1) - 1024 + 1;
    if (smem < min_smem) smem = min_smem;
    static int smem_attr_set = 0;  // launch attribute already raised to at least this (benign race: set is idempotent)
    if ((int)smem > smem_attr_set) {
        // 227 KiB is the per-CTA limit for static + dynamic together; leave 1 KiB for the kernel's static __shared__
        GSPN_CUDA_OK(cudaFuncSetAttribute(mlp_chain_kernel<4, 2, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        GSPN_CUDA_OK(cudaFuncSetAttribute(mlp_chain_kernel<8, 1, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
        smem_attr_set = 226 * 1024;
    }
    long grid = (long)sms * occ;
    if (grid > p.ntiles) grid = p.ntiles;
    if (p.epi_warps == 4) mlp_chain_kernel<4, 2, 32><<<(unsigned)grid, 4 * 32 + 96, smem, s>>>(p, tm_f32, tm_bf16);
    else mlp_chain_kernel<8, 1, 32><<<(unsigned)grid, 8 * 32 + 96, smem, s>>>(p, tm_f32, tm_bf16);
    return check_launch();
}

extern "C" int gspn_mlp_chain(long rows, int nlayers, const int *dims, const void *a, const void *const *wimg, const float *const *scale,
                              const float *const *shift, const int *relu, int pool, float *out_f32, void *out_bf16, gspn_stream_t stream) {
    ChainParams p = {};
    return chain_launch(p, rows, nlayers, dims, a, wimg, scale, shift, relu, pool, out_f32, out_bf16, stream);
}

extern "C" int gspn_mlp_chain_gather(int b, int n, int m, int nsample, int c, const float *xyz, const float *new_xyz, const float *shift_pred,
                                     const float *points, const int *idx, int nlayers, const int *dims, const void *const *wimg,
                                     const float *const *scale, const float *const *shift, const int *relu, int pool, float *out_f32,
                                     void *out_bf16, gspn_stream_t stream) {
    GSPN_REQUIRE(b >= 0 && n > 0 && m >= 0 && nsample > 0 && c >= 0);
    if (c + 3 > 8) return GSPN_E_UNSUPPORTED;  // one 16-byte chunk per row; wider rows go through the tile image
    if (b == 0 || m == 0) return GSPN_OK;
    GSPN_REQUIRE_PTR(xyz); GSPN_REQUIRE_PTR(new_xyz); GSPN_REQUIRE_PTR(idx); GSPN_REQUIRE_PTR(dims);
    if (c > 0) GSPN_REQUIRE_PTR(points);
    GSPN_REQUIRE(dims[0] == 64);
    ChainParams p = {};
    p.g_idx = idx; p.g_xyz = xyz; p.g_ctr = new_xyz; p.g_shift = shift_pred; p.g_pts = points;
    p.g_n = n; p.g_m = m; p.g_k = nsample; p.g_c = c;
    if ((long)b * m * nsample >= (1L << 31)) return GSPN_E_UNSUPPORTED;  // the in-kernel row -> (query, cloud) divider is 32-bit
    p.g_div_k = FastDiv((uint32_t)nsample);
    p.g_div_m = FastDiv((uint32_t)m);
    return chain_launch(p, (long)b * m * nsample, nlayers, dims, nullptr, wimg, scale, shift, relu, pool, out_f32, out_bf16, stream);
}

extern "C" int gspn_fp_assemble(int b, int n, int m, int c1, int c2, const float *points1, const float *points2, const int *idx,
                                const float *weight, void *a_img, int ld, gspn_stream_t stream) {
    GSPN_REQUIRE(b >= 0 && n > 0 && m > 0 && c1 >= 0 && c2 > 0 && ld % 64 == 0 && ld >= c1 + c2);
    if (b == 0) return GSPN_OK;
    GSPN_REQUIRE_PTR(points2); GSPN_REQUIRE_PTR(idx); GSPN_REQUIRE_PTR(weight); GSPN_REQUIRE_PTR(a_img);
    if (c1 > 0) GSPN_REQUIRE_PTR(points1);
    long rows = (long)b * n;
    long total = rows * (ld / 8);
    long blk = ceil_div_l(total, 256);
    if (blk > 148L * 64) blk = 148L * 64;
    fp_assemble_kernel<<<(unsigned)blk, 256, 0, as_stream(stream)>>>(rows, n, m, c1, c2, points1, points2, idx, weight, (unsigned char *)a_img, ld,
                                                                     FastDiv((uint32_t)(ld >> 3)), FastDiv((uint32_t)n));
    return check_launch();
}
