// mlp_tc.cu -- the shared per-point MLP (+ max-pool over nsample) of pointnet_sa_module /
// pointnet_fp_module (utils/pointnet_util.py:109-113,124,167-172; tf_util.conv2d 1x1 + bias + BN + ReLU,
// utils/tf_util.py:170-184,530-534) as ONE tcgen05 kernel per module: the whole layer chain runs on a
// 128-row tile without the activations ever leaving the SM -- from the second layer on they do not even leave
// TENSOR MEMORY: the epilogue writes a layer's output back into TMEM with tcgen05.st and the next layer's
// tcgen05.mma takes its A operand from there.
//
//   * arithmetic: GSPN_MLP_BF16 (one bf16 product per term, unit round-off 2^-8) or GSPN_MLP_BF16X3, the default of the
//     Python modules: every fp32 operand x is carried as the pair hi = bf16(x), lo = bf16(x - hi) and a product a*b is
//     accumulated as  a_hi*b_hi + a_lo*b_hi + a_hi*b_lo  into the same fp32 TMEM accumulator (three tcgen05.mma per k-slice;
//     the dropped lo*lo term is below 2^-16 relative) -- the reference computes these layers in fp32, and this is what keeps
//     the tensor-core path inside the 1e-3 parity bound with two decimal orders to spare.
//   * layer 0's A operand (128 rows x K, K-major, SWIZZLE_128B; in split mode a [hi | lo] pair of blocks) comes through a
//     shared-memory ring, filled one 64-column block per stage by
//       - one cp.async.bulk per block from the tile image written by the fused ball-query+group kernel / gspn_fp_assemble, or
//       - four producer warps that build it in place: the neighbourhood rows [features | xyz - centre | 0] gathered straight
//         from the ball-query indices (narrow rows: SA1, the context encoder; no grouped tensor in HBM), or the OUTPUT of a
//         feature-propagation module's first layer computed on the CUDA cores from the pre-multiplied coarse features
//         (gspn_mlp_chain_fp: three_interpolate is linear, so interp(points2) @ W = interp(points2 @ W); the products for the
//         m known points are computed once instead of once per interpolated point and no interpolated map is ever written).
//   * layers >= 1: A operand in TMEM (row = lane, two bf16 per 32-bit column, 8 columns per K=16 slice), written by the
//     epilogue of the previous layer IN PLACE (that epilogue only starts when every MMA of its layer has completed).
//   * B operand (weights): W^T as [cout x cin] bf16 K-major blocks (hi and lo images in split mode), pre-swizzled once by
//     gspn_mlp_pack_weights, streamed through a shared-memory ring by cp.async.bulk with mbarrier complete_tx; the ring
//     runs ahead across layers and tiles.  Only the ceil(K/16) k-slices that hold real columns are issued.
//   * D accumulates in TMEM (fp32, 128 lanes x up to 256 columns per buffer; a wider last layer runs in passes);
//     tcgen05.commit releases ring stages and signals the epilogue.
//   * warp roles: 4 (two CTAs per SM) or 8 (one CTA per SM) epilogue warps, 1 or 4 operand-producer warps, a weight-producer
//     warp and a CONVERGED MMA-issuer warp (elect.sync around a block's tcgen05.mma + commit).
//   * epilogue: software-pipelined tcgen05.ld 32x32b.x32 -> fp32 scale/shift (bias+BN folded) -> ReLU -> bf16 (pair) ->
//     tcgen05.st (next layer's A); last layer: max over the nsample rows of a group by recursive-halving shuffles, or
//     (pool == 1) per-warp swizzled staging boxes handed to TMA tensor stores (fp32 and/or a 16-bit copy).
#include <cstdlib>
#include <cstring>
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link dependency)
#include <cuda_fp16.h>
#include "common.cuh"

namespace gspn {

constexpr int kMaxLayers = 4;
constexpr int kMaxStages = 4;
enum { kModeBulk = 0, kModeGatherSA = 1, kModeFP = 2 };

struct ChainParams {
    long rows;
    long ntiles;
    int nlayers;         // layers run as MMAs (gspn_mlp_chain_fp: without the producer-computed first layer)
    int KB[kMaxLayers];  // 64-column weight blocks of layer l
    int NS[kMaxLayers];  // 16-column k-slices issued (the real input width rounded up to 16; <= 4 * KB)
    int N[kMaxLayers];   // output width (multiple of 32)
    const unsigned char *wimg[kMaxLayers];
    const float *scale[kMaxLayers];
    const float *shift[kMaxLayers];
    int relu[kMaxLayers];
    int mode;
    const unsigned char *a;  // kModeBulk: the tile image
    int pool;
    float *out_f32;
    int out_f32_vec;  // out_f32 is 32-byte aligned: 256-bit stores
    void *out_h;      // optional 16-bit copy of the output
    int out_h_f16;    // 1: IEEE half (saturating), 0: bf16
    int nch;          // weight rows per ring stage = accumulator columns per MMA
    int dcols;        // accumulator buffer width; the last layer runs in passes of dcols columns
    int tm_bufs;      // accumulator buffers (2: a tile's first layer is issued while the previous tile's last epilogue still drains)
    int tmem_cols;    // allocation (power of two)
    int a_col, a_lo_off;  // TMEM column of the activation operand (hi), distance to the lo half (split mode)
    int epi_warps;
    int a_stages, w_stages;
    uint32_t a_stage_bytes, w_stage_bytes;
    uint32_t stage_off, stg_bytes, stg_h_off;  // per-epilogue-warp output staging boxes (TMA output path)
    uint32_t affine_off;
    int tma_out;
    int dyn_tiles;    // 1: the grid has one CTA per tile and running CTAs take over not-yet-launched ones (walk_next); 0: static stride
    long long *prof;  // optional: per-phase cycle counters of CTA 0 (tools/tc_profile.py)
    // kModeGatherSA: layer 0's operand rows [features(c) | xyz - centre - shift | 0] (c+3 <= 8) built from the ball-query indices
    const int *g_idx;      // (b, m, nsample)
    const float *g_xyz;    // (b, n, 3)
    const float *g_ctr;    // (b, m, 3)
    const float *g_shift;  // (b, m, 3) or null
    const float *g_pts;    // (b, n, c) f32 or null
    int g_n, g_m, g_k, g_c;
    FastDiv g_div_k, g_div_m;  // row / nsample, query / m (rows < 2^31)
    // kModeFP: the first MMA layer's operand = act(scale0 * (interp3(y2) + points1 @ w0b) + shift0)
    const float *f_y2;   // (b, m, n0): points2 @ W0[:c2]
    const int *f_idx;    // (b, n, 3)
    const float *f_w;    // (b, n, 3)
    const float *f_p1;   // (b, n, c1) or null
    const float *f_w0b;  // (c1, n0): W0[c2:]
    const float *f_scale, *f_shift;
    int f_relu, f_n, f_m, f_c1, f_n0;
    FastDiv f_div_n;
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t s_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mb_init(uint32_t bar, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mb_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mb_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
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
// ---- which tile next.  Static (the default): tile += gridDim.x of a grid sized to the SMs.  Dynamic (gspn_mlp_chain_tune_sched(1)): the
// grid has one CTA per tile, and a CTA that finishes a tile cancels a CTA the hardware has not launched yet and does that CTA's tile
// instead (clusterlaunchcontrol.try_cancel, the work-stealing form of a persistent kernel; tools/probes/clc_probe.cu: every tile
// exactly once, ~760 cycles per query).  Built for the pipelined step, where FPS clusters hold part of the GPU for a millisecond and
// a static grid's CTAs that found no SM run as a second wave; measured there it is a wash (0.735 vs 0.722 ms/step: the step is bound
// by register-file occupancy, not by wave quantisation) and 7 us slower on the 256-tile chains, so it stays a door.
// One lane (operand-producer warp 0) sends the queries, one in flight beyond the last answer read; every warp of the CTA walks
// the same answers through a ring of kSched (full, empty) mbarrier pairs.  A query is only sent after a successful one.
constexpr int kSched = 8;
struct __align__(16) SchedRing {
    uint4 resp[kSched];      // the 16-byte answers
    uint64_t full[kSched];   // answer landed (complete_tx)
    uint64_t empty[kSched];  // every warp of the CTA has read it
};
__device__ __forceinline__ void walk_issue(uint32_t ring, int k) {
    const int s = k & (kSched - 1), u = k / kSched;
    const uint32_t full = ring + (uint32_t)offsetof(SchedRing, full) + 8 * s, empty = ring + (uint32_t)offsetof(SchedRing, empty) + 8 * s;
    if (u >= 1) mb_wait_relaxed(empty, (uint32_t)((u - 1) & 1));  // every warp has read the slot's previous answer
    mb_expect_tx(full, 16);
    asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.b128 [%0], [%1];" ::"r"(ring + 16 * s), "r"(full)
                 : "memory");
}
// The tile after `tile`, or -1.  WARP: called by all 32 lanes (one arrival per warp); otherwise by a single thread.  q: the caller's
// count of answers read so far.  issuer: this thread sends the queries.
template <bool WARP>
__device__ __forceinline__ int walk_next(const ChainParams &p, uint32_t ring, int &q, int tile, int lane, bool issuer) {
    if (!p.dyn_tiles) {
        tile += gridDim.x;
        return tile < (int)p.ntiles ? tile : -1;
    }
    const int s = q & (kSched - 1), u = q / kSched;
    mb_wait(ring + (uint32_t)offsetof(SchedRing, full) + 8 * s, (uint32_t)(u & 1));
    uint32_t valid, x;
    asm volatile(
        "{\n\t.reg .pred p1;\n\t.reg .b128 r;\n\t"
        "ld.shared.b128 r, [%2];\n\t"
        "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, r;\n\t"
        "selp.u32 %1, 1, 0, p1;\n\t"
        "mov.u32 %0, 0;\n\t"
        "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid.v4.b32.b128 {%0, _, _, _}, r;\n\t}"
        : "=r"(x), "=r"(valid)
        : "r"(ring + 16 * s)
        : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the slot is rewritten through the async proxy
    if (WARP) __syncwarp();
    if (!WARP || lane == 0) mb_arrive(ring + (uint32_t)offsetof(SchedRing, empty) + 8 * s);
    ++q;
    if (issuer && valid) walk_issue(ring, q);
    return valid ? (int)x : -1;
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
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]^T, bf16 inputs, fp32 accumulate
__device__ __forceinline__ void tc_mma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the same with the A operand in tensor memory (row = lane, elements 2j / 2j+1 of the row = low / high half of column j; checked on
// a B200 by tools/probes/ts_mma_probe.cu)
__device__ __forceinline__ void tc_mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
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
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
        "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

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

// {lo = bf16(max(a,0)), hi = bf16(max(b,0))} in one instruction (F2FP.RELU)
__device__ __forceinline__ uint32_t pack2_relu(float a, float b) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
    return d;
}
// 16-bit copy of the output: bf16, or IEEE half saturated to the finite range
__device__ __forceinline__ uint32_t pack2_out(float a, float b, int f16) {
    if (f16) {
        __half2 h = __floats2half2_rn(fminf(fmaxf(a, -65504.f), 65504.f), fminf(fmaxf(b, -65504.f), 65504.f));
        return *reinterpret_cast<uint32_t *>(&h);
    }
    return pack_bf16x2(a, b);
}
// one 256-bit global store (STG.E.256): 8 consecutive floats, 32-byte aligned
__device__ __forceinline__ void st_global_v8(float *dst, const float *v) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]),
                 "f"(v[5]), "f"(v[6]), "f"(v[7])
                 : "memory");
}

// ---------------------------------------------------------------- epilogue
struct EpiCtx {
    const float *sc, *sh;  // this layer's folded scale / shift (shared memory), indexed by the layer's column
    unsigned char *stg;    // this warp's output staging boxes (TMA output path)
    uint32_t ta_hi;        // TMEM address (this warp's lane quadrant) of the activation operand's hi half
    int Nl, col0, lane;    // layer width; first layer column of this pass; lane
    long grow, row0;       // global row of this thread; global row of the warp's first row
    bool last, relu;
};

// one step: 32 accumulator columns [c0, c0+32) of the pass (layer columns col0 + c0 ...) of one row per thread
template <bool SPLIT>
__device__ __forceinline__ void epi_chunk(const ChainParams &p, const CUtensorMap *tm_f32, const CUtensorMap *tm_h, const EpiCtx &e,
                                          const uint32_t (&v)[32], const int c0) {
    constexpr int CW = 32;
    const int lane = e.lane, Nl = e.Nl, col = e.col0 + c0;
    float f[CW];
    {
        const float4 *sc4 = reinterpret_cast<const float4 *>(e.sc + col), *sh4 = reinterpret_cast<const float4 *>(e.sh + col);
#pragma unroll
        for (int g = 0; g < CW / 4; ++g) {
            const float4 a4 = sc4[g], b4 = sh4[g];  // same address in every lane: one broadcast LDS.128 each
            f[4 * g] = fmaf(__uint_as_float(v[4 * g]), a4.x, b4.x); f[4 * g + 1] = fmaf(__uint_as_float(v[4 * g + 1]), a4.y, b4.y);
            f[4 * g + 2] = fmaf(__uint_as_float(v[4 * g + 2]), a4.z, b4.z); f[4 * g + 3] = fmaf(__uint_as_float(v[4 * g + 3]), a4.w, b4.w);
        }
    }
    if (!e.last) {
        // the next layer's A operand, back into tensor memory: thread = row = lane, two bf16 per 32-bit column
        uint32_t hi[CW / 2];
        if constexpr (SPLIT) {
            uint32_t lo[CW / 2];
            if (e.relu) {
#pragma unroll
                for (int i = 0; i < CW; ++i) f[i] = fmaxf(f[i], 0.f);
            }
#pragma unroll
            for (int g = 0; g < CW / 2; ++g) split_bf16x2(f[2 * g], f[2 * g + 1], hi[g], lo[g]);
            tc_st16(e.ta_hi + (uint32_t)(col >> 1), hi);
            tc_st16(e.ta_hi + (uint32_t)(p.a_lo_off + (col >> 1)), lo);
        } else {
            if (e.relu) {  // ReLU rides on the bf16 conversion (cvt.rn.relu.bf16x2.f32): no separate max per element
#pragma unroll
                for (int g = 0; g < CW / 2; ++g) hi[g] = pack2_relu(f[2 * g], f[2 * g + 1]);
            } else {
#pragma unroll
                for (int g = 0; g < CW / 2; ++g) hi[g] = pack_bf16x2(f[2 * g], f[2 * g + 1]);
            }
            tc_st16(e.ta_hi + (uint32_t)(col >> 1), hi);
        }
        return;
    }
    if (e.relu) {
#pragma unroll
        for (int i = 0; i < CW; ++i) f[i] = fmaxf(f[i], 0.f);
    }
    if (p.pool == 1) {
        if (p.tma_out) {
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
            if (p.out_h) {
#pragma unroll
                for (int g = 0; g < CW / 8; ++g)
                    *reinterpret_cast<uint4 *>(stg + p.stg_h_off + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4)) =
                        make_uint4(pack2_out(f[8 * g], f[8 * g + 1], p.out_h_f16), pack2_out(f[8 * g + 2], f[8 * g + 3], p.out_h_f16),
                                   pack2_out(f[8 * g + 4], f[8 * g + 5], p.out_h_f16), pack2_out(f[8 * g + 6], f[8 * g + 7], p.out_h_f16));
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
                if (p.out_f32) tma_store_2d(tm_f32, s_u32(stg), col, (int)e.row0);
                if (p.out_h) tma_store_2d(tm_h, s_u32(stg + p.stg_h_off), col, (int)e.row0);
                bulk_commit();
            }
        } else if (e.grow < p.rows) {
            // thread = row: its CW columns are contiguous bytes of the output row; 256-bit stores (one 32-byte sector per lane)
            if (p.out_f32) {
                float *o = p.out_f32 + e.grow * Nl + col;
                if (p.out_f32_vec) {
#pragma unroll
                    for (int g = 0; g < CW / 8; ++g) st_global_v8(o + 8 * g, &f[8 * g]);
                } else {
#pragma unroll
                    for (int i = 0; i < CW; ++i) o[i] = f[i];
                }
            }
            if (p.out_h) {
                uint4 *o = reinterpret_cast<uint4 *>(reinterpret_cast<uint16_t *>(p.out_h) + e.grow * Nl + col);
#pragma unroll
                for (int g = 0; g < CW / 8; ++g)
                    o[g] = make_uint4(pack2_out(f[8 * g], f[8 * g + 1], p.out_h_f16), pack2_out(f[8 * g + 2], f[8 * g + 3], p.out_h_f16),
                                      pack2_out(f[8 * g + 4], f[8 * g + 5], p.out_h_f16), pack2_out(f[8 * g + 6], f[8 * g + 7], p.out_h_f16));
            }
        }
    } else {
        // max over the 32 rows this warp holds, for CW columns at once: recursive halving -- at the step of lane
        // bit h a lane keeps the half of its columns selected by that bit and takes the partner's values for them
        // (CW-1 SHFL + FMNMX instead of CW warp-wide redux); lane L ends with the max of column L
#pragma unroll
        for (int h = CW / 2; h >= 1; h >>= 1) {
            const bool up = lane & h;
#pragma unroll
            for (int i = 0; i < h; ++i) {
                const float send = up ? f[i] : f[i + h], keepv = up ? f[i + h] : f[i];
                f[i] = fmaxf(keepv, __shfl_xor_sync(GSPN_FULL_MASK, send, h));
            }
        }
        const int keep = __float_as_int(f[0]);
        if (e.row0 < p.rows) {
            const long grp = e.row0 / p.pool;
            const int oc = col + lane;
            if (p.pool == 32) {
                if (p.out_f32) p.out_f32[grp * Nl + oc] = __int_as_float(keep);
                if (p.out_h) {
                    if (p.out_h_f16) reinterpret_cast<__half *>(p.out_h)[grp * Nl + oc] = __float2half_rn(fminf(__int_as_float(keep), 65504.f));
                    else reinterpret_cast<__nv_bfloat16 *>(p.out_h)[grp * Nl + oc] = __float2bfloat16_rn(__int_as_float(keep));
                }
            } else {
                atomicMax(reinterpret_cast<int *>(p.out_f32) + grp * Nl + oc, keep);  // post-ReLU values: int order = float order; out zeroed by the launcher
            }
        }
    }
}

// columns [c_lo, c_hi) of this warp's 32 rows: software-pipelined TMEM reads (the load of the next 32 columns is in flight
// while this one is processed)
template <bool SPLIT>
__device__ __forceinline__ void epi_columns(const ChainParams &p, const CUtensorMap *tm_f32, const CUtensorMap *tm_h, const EpiCtx &e,
                                            const uint32_t tbase, const int c_lo, const int c_hi) {
    constexpr int CW = 32;
    uint32_t va[CW], vb[CW];
    if (c_lo < c_hi) tc_ld32(tbase + c_lo, va);
    for (int c0 = c_lo; c0 < c_hi; c0 += 2 * CW) {
        tc_wait_ld();
        if (c0 + CW < c_hi) tc_ld32(tbase + c0 + CW, vb);
        epi_chunk<SPLIT>(p, tm_f32, tm_h, e, va, c0);
        if (c0 + CW < c_hi) {
            tc_wait_ld();
            if (c0 + 2 * CW < c_hi) tc_ld32(tbase + c0 + 2 * CW, va);
            epi_chunk<SPLIT>(p, tm_f32, tm_h, e, vb, c0 + CW);
        }
    }
}

// EPI epilogue warps (4: one per TMEM lane quadrant, 8: two per quadrant, each taking half the columns); MINB CTAs per SM
// (register budget); SPLIT: bf16x3 arithmetic; NPW operand-producer warps: 1 (one lane issuing bulk copies of the tile image),
// 2 (neighbourhood rows gathered from the ball-query indices) or 8 (feature-propagation first layer: an L2 gather of 1.5 KB per row
// that lives on the number of loads in flight: measured 0.116 ms with 4 epilogue + 8 gather warps, 0.145 ms with 8 + 6)
template <int EPI, int MINB, bool SPLIT, int NPW>
__global__ void __launch_bounds__((EPI + NPW + 2) * 32, MINB)
    mlp_chain_kernel(const ChainParams p, const __grid_constant__ CUtensorMap tm_f32, const __grid_constant__ CUtensorMap tm_h) {
    constexpr int kSplitMul = SPLIT ? 2 : 1;
    extern __shared__ unsigned char smem_raw[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bars[4 * kMaxStages + 4];
    __shared__ SchedRing sched;

    // warp index through a shuffle: provably warp-uniform, so the role branches below are uniform branches and the
    // MMA issuer's operands can live in uniform registers (no per-instruction R2UR waterfall)
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(GSPN_FULL_MASK, tid >> 5, 0);
    const uint32_t raw = s_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;  // SWIZZLE_128B atoms are 1024-byte aligned
    unsigned char *sm = smem_raw + (base - raw);
    const uint32_t aring = base;
    const uint32_t wring = aring + (uint32_t)p.a_stages * p.a_stage_bytes;
    float *affine = reinterpret_cast<float *>(sm + p.affine_off);
    const uint32_t w_full = s_u32(&bars[0]), w_empty = s_u32(&bars[kMaxStages]), a_full = s_u32(&bars[2 * kMaxStages]),
                   a_empty = s_u32(&bars[3 * kMaxStages]), mma_done = s_u32(&bars[4 * kMaxStages]),
                   epi_done = s_u32(&bars[4 * kMaxStages + 2]);  // [2] each: one per TMEM accumulator buffer

    if (tid == 0) {
        for (int i = 0; i < 4 * kMaxStages + 2; ++i)  // a_full: every producer warp arrives once per stage (bulk mode: one expect_tx)
            mb_init(s_u32(&bars[i]), (i >= 2 * kMaxStages && i < 3 * kMaxStages) ? NPW : 1);
        mb_init(epi_done, p.epi_warps);  // one elected lane per epilogue warp arrives once per step
        mb_init(epi_done + 8, p.epi_warps);
        for (int i = 0; i < kSched; ++i) {
            mb_init(s_u32(&sched.full[i]), 1);
            mb_init(s_u32(&sched.empty[i]), EPI + NPW + 2);
        }
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
    if (p.mode == kModeGatherSA) {
        // each row only ever writes its first 16-byte chunk; the k-slice the MMA reads spans two chunks: keep the rest zero
        uint4 z = make_uint4(0, 0, 0, 0);
        uint4 *r = reinterpret_cast<uint4 *>(sm);
        for (uint32_t i = tid; i < (uint32_t)p.a_stages * p.a_stage_bytes / 16; i += blockDim.x) r[i] = z;
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

    const int last_l = p.nlayers - 1;
    const int npass_last = (p.N[last_l] + p.dcols - 1) / p.dcols;
    const int kb0 = p.KB[0];
    const uint32_t ring = s_u32(&sched);
    int wq = 0;  // answers of the tile ring this warp has read

    if (warp >= EPI && warp < EPI + NPW) {
        // ================= operand producers: keep the layer-0 ring full for the whole kernel, independent of the MMA/epilogue
        // timeline, so the operand of the next tile is in flight while the epilogue warps are busy
        if (p.mode == kModeBulk) {
            if (warp == EPI && lane == 0) {
                int s = 0, par = 0;
                long i = 0;
                if (p.dyn_tiles) walk_issue(ring, 0);
                for (int t = blockIdx.x; t >= 0; t = walk_next<false>(p, ring, wq, t, 0, true)) {
                    for (int kb = 0; kb < kb0; ++kb, ++i) {
                        if (i >= p.a_stages) mb_wait_relaxed(a_empty + 8 * s, (uint32_t)(par ^ 1));
                        mb_expect_tx(a_full + 8 * s, p.a_stage_bytes);
                        bulk_load(aring + s * p.a_stage_bytes, p.a + ((size_t)t * kb0 + kb) * p.a_stage_bytes, p.a_stage_bytes, a_full + 8 * s);
                        if (++s == p.a_stages) { s = 0; par ^= 1; }
                    }
                }
            }
        } else if (NPW == 2 && p.mode == kModeGatherSA) {
            // warp pw builds rows 64*pw + lane and + 32 of the tile: one 16-byte chunk (hi) [+ one (lo)] per row.  Memory-level
            // parallelism is what these warps live on: the NEXT tile's indices are requested before this tile is built, and a tile's
            // dependent coordinate / feature gathers are all issued before any is consumed.
            const int pw = warp - EPI;
            const int c = p.g_c;
            int s = 0, par = 0;
            int t = blockIdx.x;
            const bool sender = pw == 0 && lane == 0;
            int ii[2], nxt[2];
            auto load_idx = [&](int tile, int (&dst)[2]) {
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const long row = (long)tile * kTileRows + pw * 64 + rr * 32 + lane;
                    dst[rr] = (tile >= 0 && tile < p.ntiles && row < p.rows) ? __ldg(p.g_idx + row) : -1;
                }
            };
            if (sender && p.dyn_tiles) walk_issue(ring, 0);
            int tn = walk_next<true>(p, ring, wq, t, lane, sender);  // one tile ahead: its indices are requested while this one is built
            load_idx(t, nxt);
            for (long i = 0; t >= 0; ++i) {
                ii[0] = nxt[0]; ii[1] = nxt[1];
                load_idx(tn, nxt);
                float gx[2][3], cx[2][3], sh[2][3], f[2][5];
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const long row = (long)t * kTileRows + pw * 64 + rr * 32 + lane;
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
                unsigned char *stage = sm + (size_t)s * p.a_stage_bytes;
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int r = pw * 64 + rr * 32 + lane;
                    float d[3], v[8];
#pragma unroll
                    for (int a = 0; a < 3; ++a) {
                        d[a] = __fsub_rn(gx[rr][a], cx[rr][a]);             // grouped_xyz -= new_xyz (pointnet_util.py:42)
                        if (p.g_shift) d[a] = __fsub_rn(d[a], sh[rr][a]);   // -= shift_pred (model_rpointnet.py:56-57)
                    }
#pragma unroll
                    for (int t2 = 0; t2 < 8; ++t2)  // columns [features(c) | dx dy dz | 0]; c is a runtime value <= 5
                        v[t2] = (ii[rr] < 0) ? 0.f
                                             : ((t2 < c) ? f[rr][t2 < 5 ? t2 : 4] : (t2 == c ? d[0] : (t2 == c + 1 ? d[1] : (t2 == c + 2 ? d[2] : 0.f))));
                    uint4 pk, pl;
                    if constexpr (SPLIT) {
                        split_bf16x2(v[0], v[1], pk.x, pl.x); split_bf16x2(v[2], v[3], pk.y, pl.y);
                        split_bf16x2(v[4], v[5], pk.z, pl.z); split_bf16x2(v[6], v[7], pk.w, pl.w);
                    } else {
                        pk = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
                    }
                    unsigned char *dst = stage + (r >> 3) * 1024 + (r & 7) * 128 + ((r & 7) << 4);  // chunk 0 ^ (r & 7)
                    *reinterpret_cast<uint4 *>(dst) = pk;
                    if constexpr (SPLIT) *reinterpret_cast<uint4 *>(dst + kTileBytes) = pl;
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mb_arrive(a_full + 8 * s);
                if (++s == p.a_stages) { s = 0; par ^= 1; }
                t = tn;
                if (t >= 0) tn = walk_next<true>(p, ring, wq, t, lane, sender);
            }
        } else if ((NPW == 8 || NPW == 4) && p.mode == kModeFP) {
            // the feature-propagation module's first layer on the CUDA cores, from the pre-multiplied coarse features:
            //   y[row, :] = act(scale0 * (w1*y2[i1,:] + w2*y2[i2,:] + w3*y2[i3,:] + points1[row,:] @ w0b) + shift0)
            // written as the (split) bf16 operand blocks of the first MMA layer (n0 = 128: two 64-column blocks = two ring stages,
            // filled together).  A warp owns a row (32 lanes x float4), so a row's indices, weights and skip-link values are loaded once;
            // the tile's 128 rows are dealt round-robin to the eight warps, four rows (12 x 128-bit gathers per lane) in flight.
            const int pw = warp - EPI;
            const int h = lane & 15, kb = lane >> 4;
            const int n0 = p.f_n0, c1 = p.f_c1;
            const int col = 4 * lane;
            const float4 sc = __ldg(reinterpret_cast<const float4 *>(p.f_scale + col)), sf = __ldg(reinterpret_cast<const float4 *>(p.f_shift + col));
            float4 wb[4];
#pragma unroll
            for (int a = 0; a < 4; ++a)
                wb[a] = a < c1 ? __ldg(reinterpret_cast<const float4 *>(p.f_w0b + (size_t)a * n0 + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
            int s = 0, par = 0;  // s: first of the tile's two stages (a_stages is even in this mode)
            long i = 0;
            const bool sender = pw == 0 && lane == 0;
            if (sender && p.dyn_tiles) walk_issue(ring, 0);
            for (int t = blockIdx.x; t >= 0; t = walk_next<true>(p, ring, wq, t, lane, sender), ++i) {
                if (2 * i >= p.a_stages) {
                    mb_wait_relaxed(a_empty + 8 * s, (uint32_t)(par ^ 1));
                    mb_wait_relaxed(a_empty + 8 * (s + 1), (uint32_t)(par ^ 1));
                }
                unsigned char *stage = sm + (size_t)(s + kb) * p.a_stage_bytes;
                for (int st0 = pw; st0 < kTileRows; st0 += 4 * NPW) {
                    float4 y1[4], y2[4], y3[4];
                    float w1[4], w2[4], w3[4], q1[4][4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = st0 + u * NPW;
                        const long grow = (long)t * kTileRows + r;
                        const bool ok = grow < p.rows;
                        const long gr = ok ? grow : 0;
                        const uint32_t cloud = p.f_div_n.div((uint32_t)gr);
                        const int *ip = p.f_idx + gr * 3;
                        const float *wp = p.f_w + gr * 3;
                        const float *yb = p.f_y2 + (size_t)cloud * p.f_m * n0 + col;
                        w1[u] = ok ? __ldg(wp) : 0.f; w2[u] = ok ? __ldg(wp + 1) : 0.f; w3[u] = ok ? __ldg(wp + 2) : 0.f;
                        y1[u] = __ldg(reinterpret_cast<const float4 *>(yb + (size_t)__ldg(ip) * n0));
                        y2[u] = __ldg(reinterpret_cast<const float4 *>(yb + (size_t)__ldg(ip + 1) * n0));
                        y3[u] = __ldg(reinterpret_cast<const float4 *>(yb + (size_t)__ldg(ip + 2) * n0));
#pragma unroll
                        for (int a = 0; a < 4; ++a) q1[u][a] = (ok && a < c1) ? __ldg(p.f_p1 + gr * c1 + a) : 0.f;
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int r = st0 + u * NPW;
                        float v[4];
                        v[0] = fmaf(y3[u].x, w3[u], fmaf(y2[u].x, w2[u], y1[u].x * w1[u]));
                        v[1] = fmaf(y3[u].y, w3[u], fmaf(y2[u].y, w2[u], y1[u].y * w1[u]));
                        v[2] = fmaf(y3[u].z, w3[u], fmaf(y2[u].z, w2[u], y1[u].z * w1[u]));
                        v[3] = fmaf(y3[u].w, w3[u], fmaf(y2[u].w, w2[u], y1[u].w * w1[u]));
#pragma unroll
                        for (int a = 0; a < 4; ++a) {
                            v[0] = fmaf(q1[u][a], wb[a].x, v[0]); v[1] = fmaf(q1[u][a], wb[a].y, v[1]);
                            v[2] = fmaf(q1[u][a], wb[a].z, v[2]); v[3] = fmaf(q1[u][a], wb[a].w, v[3]);
                        }
                        v[0] = fmaf(v[0], sc.x, sf.x); v[1] = fmaf(v[1], sc.y, sf.y); v[2] = fmaf(v[2], sc.z, sf.z); v[3] = fmaf(v[3], sc.w, sf.w);
                        if (p.f_relu) {
#pragma unroll
                            for (int a = 0; a < 4; ++a) v[a] = fmaxf(v[a], 0.f);
                        }
                        unsigned char *dst = stage + (r >> 3) * 1024 + (r & 7) * 128 + (((h >> 1) ^ (r & 7)) << 4) + (h & 1) * 8;
                        uint2 pk, pl;
                        if constexpr (SPLIT) {
                            split_bf16x2(v[0], v[1], pk.x, pl.x); split_bf16x2(v[2], v[3], pk.y, pl.y);
                            *reinterpret_cast<uint2 *>(dst + kTileBytes) = pl;
                        } else {
                            pk = make_uint2(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]));
                        }
                        *reinterpret_cast<uint2 *>(dst) = pk;
                    }
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) { mb_arrive(a_full + 8 * s); mb_arrive(a_full + 8 * (s + 1)); }
                s += 2;
                if (s == p.a_stages) { s = 0; par ^= 1; }
            }
        }
    } else if (warp == EPI + NPW) {
        // ================= weight producer: one lane streams every layer's blocks in the order the issuer consumes them
        if (lane == 0) {
            int s = 0, par = 0;
            long cnt = 0;
            for (int t = blockIdx.x; t >= 0; t = walk_next<false>(p, ring, wq, t, 0, false)) {
                for (int l = 0; l < p.nlayers; ++l) {
                    const int Nl = p.N[l];
                    const int npass = (l == last_l) ? npass_last : 1;
                    const unsigned char *wl = p.wimg[l];
                    for (int ps = 0; ps < npass; ++ps) {
                        const int cols = min(p.dcols, Nl - ps * p.dcols);
                        const int nchunks = (cols + p.nch - 1) / p.nch;
                        for (int kb = 0; kb < p.KB[l]; ++kb) {
                            for (int nc = 0; nc < nchunks; ++nc, ++cnt) {
                                if (cnt >= p.w_stages) mb_wait_relaxed(w_empty + 8 * s, (uint32_t)(par ^ 1));
                                const int rows_i = min(p.nch, cols - nc * p.nch);
                                const uint32_t bytes = (uint32_t)rows_i * 128u;
                                const unsigned char *src = wl + (size_t)kb * Nl * 128 * kSplitMul + (size_t)(ps * p.dcols + nc * p.nch) * 128;
                                const uint32_t dst = wring + s * p.w_stage_bytes;
                                mb_expect_tx(w_full + 8 * s, bytes * kSplitMul);
                                bulk_load(dst, src, bytes, w_full + 8 * s);
                                if constexpr (SPLIT) bulk_load(dst + (uint32_t)p.nch * 128u, src + (size_t)Nl * 128, bytes, w_full + 8 * s);
                                if (++s == p.w_stages) { s = 0; par ^= 1; }
                            }
                        }
                    }
                }
            }
        }
    } else if (warp == EPI + NPW + 1) {
        // ================= MMA issuer warp.  The whole warp runs the loop CONVERGED (barrier waits, ring bookkeeping and descriptor
        // arithmetic stay warp-uniform, so they live in uniform registers next to the UTCHMMA operands); only the
        // tcgen05.mma / tcgen05.commit instructions themselves are executed by one lane.  The layer-to-layer critical path is
        //   wait(epi_done) -> tcgen05.mma ... -> commit(mma_done)
        int wu_s = 0, wu_par = 0, au_s = 0, au_par = 0;  // consumer-side stage index and round parity
        // Steps (layer, pass) alternate between the TMEM accumulator buffers (tm_bufs == 2).  Step s may be issued when
        //   (1) its buffer is free: the epilogue of step s - tm_bufs has drained it, and
        //   (2) for the first pass of a layer > 0, its A operand is written: the epilogue of step s - 1 is done.
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
        const uint32_t a_lo0 = desc_lo(aring), w_lo0 = desc_lo(wring);
        const uint32_t a_step = p.a_stage_bytes >> 4, w_step = p.w_stage_bytes >> 4;
        const uint32_t b_lo_off = ((uint32_t)p.nch * 128u) >> 4;   // lo weight rows inside a stage
        const uint32_t a_lo_blk = (uint32_t)kTileBytes >> 4;        // lo operand block inside a layer-0 stage
        const uint32_t ta_hi = tmem + (uint32_t)p.a_col, ta_lo = ta_hi + (uint32_t)p.a_lo_off;
        for (int tile = blockIdx.x; tile >= 0; tile = walk_next<true>(p, ring, wq, tile, lane, false)) {
            for (int l = 0; l < p.nlayers; ++l) {
                const int Nl = p.N[l], KBl = p.KB[l], NSl = p.NS[l];
                const int npass = (l == last_l) ? npass_last : 1;
                for (int ps = 0; ps < npass; ++ps, ++step) {
                    const int buf = NB == 2 ? (step & 1) : 0;
                    const int ph = NB == 2 ? (step >> 1) : step;  // this step's phase on its buffer's barriers
                    const int cols = min(p.dcols, Nl - ps * p.dcols);
                    const int nchunks = (cols + p.nch - 1) / p.nch;
                    const uint32_t idesc_full = instr_desc(128, p.nch), idesc_last = instr_desc(128, cols - (nchunks - 1) * p.nch);
                    long long mw0 = 0;
                    if (p.prof) mw0 = clock64();
                    const uint32_t tm = tmem + buf * p.dcols;
                    if (ph >= 1) wait_epi(buf, ph - 1);                                  // (1) TMEM buffer drained
                    if (l > 0 && ps == 0 && NB == 2) wait_epi(buf ^ 1, (step - 1) >> 1);  // (2) A operand written
                    long long mt0 = 0;
                    if (p.prof) {
                        mt0 = clock64();  // slot 7: epilogue hand-off, as seen by the issuer
                        if (blockIdx.x == 0 && lane == 0) atomicAdd((unsigned long long *)p.prof + 7, (unsigned long long)(mt0 - mw0));
                    }
                    tc_fence_after();
                    for (int kb = 0; kb < KBl; ++kb) {
                        const int ns = min(4, NSl - 4 * kb);  // k-slices of this block that hold real columns
                        uint32_t a_lo = 0;
                        if (l == 0) {
                            mb_wait(a_full + 8 * au_s, (uint32_t)au_par);
                            tc_fence_after();
                            a_lo = a_lo0 + (uint32_t)au_s * a_step;
                        }
                        for (int nc = 0; nc < nchunks; ++nc) {
                            const int s = wu_s;
                            mb_wait(w_full + 8 * s, (uint32_t)wu_par);
                            tc_fence_after();
                            const uint32_t idesc = (nc == nchunks - 1) ? idesc_last : idesc_full;
                            const uint32_t b_lo = w_lo0 + (uint32_t)s * w_step;
                            const uint32_t td = tm + nc * p.nch;
                            if (elect_one()) {
                                if (l == 0) {
                                    for (int k = 0; k < ns; ++k) {  // UMMA_K = 16 columns: +32 bytes = +2 in the descriptor
                                        tc_mma_ss(td, desc64(a_lo + 2 * k), desc64(b_lo + 2 * k), idesc, (kb | k) != 0);
                                        if constexpr (SPLIT) {
                                            tc_mma_ss(td, desc64(a_lo + a_lo_blk + 2 * k), desc64(b_lo + 2 * k), idesc, 1);
                                            tc_mma_ss(td, desc64(a_lo + 2 * k), desc64(b_lo + b_lo_off + 2 * k), idesc, 1);
                                        }
                                    }
                                } else {
                                    for (int k = 0; k < ns; ++k) {  // 8 TMEM columns per k-slice
                                        const uint32_t ka = (uint32_t)(8 * (4 * kb + k));
                                        tc_mma_ts(td, ta_hi + ka, desc64(b_lo + 2 * k), idesc, (kb | k) != 0);
                                        if constexpr (SPLIT) {
                                            tc_mma_ts(td, ta_lo + ka, desc64(b_lo + 2 * k), idesc, 1);
                                            tc_mma_ts(td, ta_hi + ka, desc64(b_lo + b_lo_off + 2 * k), idesc, 1);
                                        }
                                    }
                                }
                                tc_commit(w_empty + 8 * s);  // stage free once these MMAs have read it
                            }
                            __syncwarp();
                            if (++wu_s == p.w_stages) { wu_s = 0; wu_par ^= 1; }
                        }
                        if (l == 0) {
                            // a wider first layer running in passes would re-read the ring; the launcher never plans that (nlayers == 1
                            // chains have N <= dcols)
                            if (elect_one()) tc_commit(a_empty + 8 * au_s);
                            __syncwarp();
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
    } else if (warp < EPI) {
        // ================= epilogue warps
        long estep = 0;  // steps alternate between the TMEM accumulator buffers exactly as the issuer's do
        const int quad = warp & 3, half = warp >> 2, nhalf = EPI >> 2;
        const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
        for (int tile = blockIdx.x; tile >= 0; tile = walk_next<true>(p, ring, wq, tile, lane, false)) {
            int ao = 0;  // running offset of layer l's [scale | shift] in the affine table
            for (int l = 0; l < p.nlayers; ++l) {
                const int Nl = p.N[l];
                const bool last = (l == last_l);
                const int npass = last ? npass_last : 1;
                const float *sc = affine + ao, *sh = sc + Nl;
                ao += 2 * Nl;
                for (int ps = 0; ps < npass; ++ps, ++estep) {
                    const int buf = p.tm_bufs == 2 ? (int)(estep & 1) : 0;
                    const long eph = p.tm_bufs == 2 ? (estep >> 1) : estep;
                    long long pt1 = 0, pt2 = 0, pt3 = 0;
                    if (p.prof) pt1 = clock64();
                    mb_wait(mma_done + 8 * buf, (uint32_t)(eph & 1));
                    tc_fence_after();
                    if (p.prof) pt2 = clock64();
                    const int cols = min(p.dcols, Nl - ps * p.dcols);
                    const int c_lo = ((cols / 32) * half / nhalf) * 32, c_hi = ((cols / 32) * (half + 1) / nhalf) * 32;  // this warp's columns
                    EpiCtx ec;
                    ec.sc = sc; ec.sh = sh; ec.stg = sm + p.stage_off + warp * p.stg_bytes;
                    ec.ta_hi = tmem + lane_base + (uint32_t)p.a_col;
                    ec.Nl = Nl; ec.col0 = ps * p.dcols; ec.lane = lane;
                    ec.grow = (long)tile * kTileRows + quad * 32 + lane; ec.row0 = (long)tile * kTileRows + quad * 32;
                    ec.last = last; ec.relu = p.relu[l] != 0;
                    epi_columns<SPLIT>(p, &tm_f32, &tm_h, ec, tmem + lane_base + buf * p.dcols, c_lo, c_hi);
                    if (p.prof) pt3 = clock64();
                    if (!last) tc_wait_st();  // the operand this warp wrote is in tensor memory before the issuer is told
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mb_arrive(epi_done + 8 * buf);  // buffer back to the issuer
                    if (p.prof && blockIdx.x == 0 && tid == 0) {
                        long long pt4 = clock64();
                        atomicAdd((unsigned long long *)p.prof + 1, (unsigned long long)(pt2 - pt1));  // wait for MMA completion
                        atomicAdd((unsigned long long *)p.prof + 2, (unsigned long long)(pt3 - pt2));  // epilogue
                        atomicAdd((unsigned long long *)p.prof + 3, (unsigned long long)(pt4 - pt3));  // fences + hand-off
                        atomicAdd((unsigned long long *)p.prof + 4, 1ull);                              // steps
                    }
                }
            }
        }
        if (p.tma_out && lane == 0) bulk_wait0();  // this warp's output boxes have been written
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(p.tmem_cols) : "memory");
}

// ---- weights (cin,cout) f32 row-major -> [cin_padded/64] blocks of (cout x 64) bf16, K-major, 128B-swizzled; split: every block is
// the pair [hi rows | lo rows]
__global__ void pack_weights_kernel(int cin, int cin_padded, int cout, const float *__restrict__ w, const int *__restrict__ row_perm,
                                    unsigned char *__restrict__ img, int split) {
    long total = (long)cin_padded * cout;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        int k = (int)(e / cout), n = (int)(e - (long)k * cout);
        int src = row_perm ? row_perm[k] : (k < cin ? k : -1);
        float v = (src >= 0 && src < cin) ? w[(size_t)src * cout + n] : 0.f;
        int kb = k >> 6, kk = k & 63;
        size_t off = (size_t)kb * cout * (128 << split) + (size_t)(n >> 3) * 1024 + (n & 7) * 128 + ((((kk >> 3) ^ (n & 7))) << 4) + (kk & 7) * 2;
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        *reinterpret_cast<__nv_bfloat16 *>(img + off) = hi;
        if (split) *reinterpret_cast<__nv_bfloat16 *>(img + off + (size_t)cout * 128) = __float2bfloat16_rn(__fsub_rn(v, __bfloat162float(hi)));
    }
}

// ---- FP front end: [three_interpolate(points2) | points1 | 0] -> (split) bf16 tile image (one 16-byte chunk per thread).
// c2 == 0 (points2 null): plain rows -> tile image.
__global__ void __launch_bounds__(256) fp_assemble_kernel(long rows, int n, int m, int c1, int c2, const float *__restrict__ points1,
                                                          const float *__restrict__ points2, const int *__restrict__ idx,
                                                          const float *__restrict__ weight, unsigned char *__restrict__ img, int ld,
                                                          const FastDiv div_chunks, const FastDiv div_n, int split) {
    const int chunks = ld >> 3;
    const long total = rows * chunks;
    const bool small = total < (1L << 31);  // element indices fit the multiply-shift divider
    const bool vec2 = c2 > 0 && (c2 % 8 == 0) && ((reinterpret_cast<uintptr_t>(points2) & 15u) == 0);
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
        const size_t off = tile_chunk_offset(row, ch, ld, split);
        if (split) {
            uint4 pk, pl;
            split_bf16x2(v[0], v[1], pk.x, pl.x); split_bf16x2(v[2], v[3], pk.y, pl.y);
            split_bf16x2(v[4], v[5], pk.z, pl.z); split_bf16x2(v[6], v[7], pk.w, pl.w);
            *reinterpret_cast<uint4 *>(img + off) = pk;
            *reinterpret_cast<uint4 *>(img + off + kTileBytes) = pl;
        } else {
            *reinterpret_cast<uint4 *>(img + off) =
                make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]), pack_bf16x2(v[6], v[7]));
        }
    }
}

}  // namespace gspn

using namespace gspn;

static inline bool arith_ok(int arith) { return arith == GSPN_MLP_BF16 || arith == GSPN_MLP_BF16X3; }

extern "C" size_t gspn_mlp_weight_image_bytes(int cin_padded, int cout, int arith) {
    if (cin_padded <= 0 || cout <= 0 || cin_padded % 64 || cout % 8 || !arith_ok(arith)) return 0;
    return (size_t)(cin_padded / 64) * (size_t)cout * 128 * (arith == GSPN_MLP_BF16X3 ? 2 : 1);
}

extern "C" int gspn_mlp_pack_weights(int cin, int cin_padded, int cout, const float *w_f32, const int *row_perm, void *wimg, int arith,
                                     gspn_stream_t stream) {
    GSPN_REQUIRE(cin > 0 && cin_padded >= cin && cin_padded % 64 == 0 && cout > 0 && cout % 8 == 0);
    if (!arith_ok(arith)) return GSPN_E_BAD_DTYPE;
    GSPN_REQUIRE_PTR(w_f32); GSPN_REQUIRE_PTR(wimg);
    long total = (long)cin_padded * cout;
    pack_weights_kernel<<<(unsigned)ceil_div_l(total, 256), 256, 0, as_stream(stream)>>>(cin, cin_padded, cout, w_f32, row_perm,
                                                                                        (unsigned char *)wimg, arith == GSPN_MLP_BF16X3);
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
// (rows x n) row-major output as a 2-D tensor map with a 32 x 32 box: 128-byte (f32) or 64-byte (16-bit) swizzled box rows
static bool encode_out_map(CUtensorMap *tm, void *base, long rows, int n, int kind /* 0 f32, 1 bf16, 2 f16 */) {
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc || (reinterpret_cast<uintptr_t>(base) & 15u) || rows <= 0 || rows > 0x7fffffffL) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)n * (kind ? 2u : 4u)};
    const cuuint32_t box[2] = {32u, 32u}, estr[2] = {1u, 1u};
    const CUtensorMapDataType dt = kind == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : (kind == 1 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16);
    return enc(tm, dt, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, kind ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
               CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---- tuning doors (benchmark A/B runs only): process-wide, set explicitly through the ABI -- nothing is read from the environment
static struct ChainTune { int occ_cap, bufs_cap, tma_out, fp_warps, dyn_tiles; long long *prof; } g_tune = {2, 2, 1, 8, 0, nullptr};
extern "C" void gspn_mlp_chain_set_profile(long long *prof) { g_tune.prof = prof; }
extern "C" void gspn_mlp_chain_tune(int occ_cap, int bufs_cap, int tma_out) {
    g_tune.occ_cap = (occ_cap == 1) ? 1 : 2;
    g_tune.bufs_cap = (bufs_cap == 1) ? 1 : 2;
    g_tune.tma_out = tma_out != 0;
}
extern "C" void gspn_mlp_chain_tune_fp(int gather_warps) { g_tune.fp_warps = gather_warps == 4 ? 4 : 8; }
extern "C" void gspn_mlp_chain_tune_sched(int dynamic_tiles) { g_tune.dyn_tiles = dynamic_tiles != 0; }

// per-device facts the launcher needs (SM count; the opt-in shared-memory attribute is per device too)
struct DevInfo { int sms; };
static int device_info(DevInfo **out) {
    static DevInfo info[64] = {};
    int dev = 0;
    GSPN_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return GSPN_E_UNSUPPORTED;
    if (info[dev].sms == 0) {
        int v = 148;
        GSPN_CUDA_OK(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev));
        info[dev].sms = v;  // benign race: idempotent
    }
    *out = &info[dev];
    return GSPN_OK;
}

template <int EPI, int MINB, bool SPLIT, int NPW>
static cudaError_t launch_variant(const ChainParams &p, const CUtensorMap &tm_f32, const CUtensorMap &tm_h, unsigned grid, size_t smem,
                                  bool set_attr, cudaStream_t s) {
    auto k = mlp_chain_kernel<EPI, MINB, SPLIT, NPW>;
    if (set_attr) {
        // 227 KiB is the per-CTA limit for static + dynamic together; leave 1 KiB for the kernel's static __shared__
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
        if (e != cudaSuccess) return e;
    }
    k<<<grid, (EPI + NPW + 2) * 32, smem, s>>>(p, tm_f32, tm_h);
    return cudaSuccess;
}

// dims[0] = K0 (multiple of 64: the operand blocks the ring carries), dims[l+1] = cout of layer l; k0_used = real columns of layer 0's
// operand (0: all of K0).
// plan_out != nullptr: plan only (gspn_mlp_chain_plan) -- no device, no pointers but the two output sentinels, nothing launched;
// plan_out[0..11] = CTAs per SM, epilogue warps, threads per CTA, weight rows per ring stage, accumulator columns per buffer,
// accumulator buffers, TMEM columns allocated, TMEM column of the activation operand, operand ring stages, weight ring stages,
// dynamic shared memory bytes, passes of the last layer.
static int chain_launch(ChainParams p, long rows, int nlayers, const int *dims, int k0_used, const void *const *wimg, const float *const *scale,
                        const float *const *shift, const int *relu, int pool, float *out_f32, void *out_h, int out_h_dtype, int arith,
                        gspn_stream_t stream, int *plan_out = nullptr) {
    const bool plan_only = plan_out != nullptr;
    GSPN_REQUIRE(rows >= 0 && nlayers >= 1 && nlayers <= kMaxLayers && pool >= 1);
    GSPN_REQUIRE_PTR(dims);
    if (!plan_only) { GSPN_REQUIRE_PTR(wimg); GSPN_REQUIRE_PTR(scale); GSPN_REQUIRE_PTR(shift); GSPN_REQUIRE_PTR(relu); }
    if (!arith_ok(arith)) return GSPN_E_BAD_DTYPE;
    if (out_h && out_h_dtype != GSPN_DT_BF16 && out_h_dtype != GSPN_DT_F16) return GSPN_E_BAD_DTYPE;
    if (rows == 0) return plan_only ? GSPN_E_BAD_SHAPE : GSPN_OK;
    if (out_f32 == nullptr && out_h == nullptr) return GSPN_E_NULL_PTR;
    const bool split = arith == GSPN_MLP_BF16X3;
    p.rows = rows;
    p.ntiles = ceil_div_l(rows, kTileRows);
    p.nlayers = nlayers;
    int maxn = 0, dmid = 0, awords = 0;
    size_t affine_floats = 0;
    GSPN_REQUIRE(dims[0] > 0 && dims[0] % 64 == 0 && k0_used >= 0 && k0_used <= dims[0]);
    for (int l = 0; l < nlayers; ++l) {
        const int kin = l == 0 ? (k0_used ? k0_used : dims[0]) : dims[l];
        p.KB[l] = l == 0 ? dims[0] / 64 : (dims[l] + 63) / 64;
        p.NS[l] = (kin + 15) / 16;
        p.N[l] = dims[l + 1];
        GSPN_REQUIRE(p.N[l] > 0);
        if (p.N[l] % 32 != 0 || p.N[l] > 512) return GSPN_E_UNSUPPORTED;  // use the fp32 path
        if (!plan_only) {
            GSPN_REQUIRE_PTR(wimg[l]); GSPN_REQUIRE_PTR(scale[l]); GSPN_REQUIRE_PTR(shift[l]);
            p.wimg[l] = (const unsigned char *)wimg[l];
            p.scale[l] = scale[l];
            p.shift[l] = shift[l];
        }
        p.relu[l] = relu ? relu[l] : 1;
        maxn = p.N[l] > maxn ? p.N[l] : maxn;
        if (l + 1 < nlayers) {
            dmid = p.N[l] > dmid ? p.N[l] : dmid;
            awords = p.N[l] / 2 > awords ? p.N[l] / 2 : awords;  // layer l+1's operand: two bf16 per TMEM column
        }
        affine_floats += 2 * (size_t)p.N[l];
    }
    if (nlayers == 1) dmid = p.N[0];  // a single layer never runs in passes (that would re-read the operand ring)
    if (pool > 1) {
        if (pool % 32 != 0 || rows % pool != 0 || !p.relu[nlayers - 1]) return GSPN_E_UNSUPPORTED;
        if (pool != 32 && (out_f32 == nullptr || out_h != nullptr)) return GSPN_E_UNSUPPORTED;
    }
    p.prof = g_tune.prof;
    p.pool = pool;
    p.out_f32 = out_f32;
    p.out_f32_vec = (reinterpret_cast<uintptr_t>(out_f32) & 31u) == 0;
    p.out_h = out_h;
    p.out_h_f16 = out_h_dtype == GSPN_DT_F16;
    static DevInfo b200 = {148};
    DevInfo *di = &b200;
    if (!plan_only) { int rc = device_info(&di); if (rc != GSPN_OK) return rc; }
    // pool == 1 outputs leave through TMA tensor stores from per-warp staging boxes
    CUtensorMap tm_f32, tm_h;
    memset(&tm_f32, 0, sizeof(tm_f32));
    memset(&tm_h, 0, sizeof(tm_h));
    p.tma_out = pool == 1 && g_tune.tma_out;
    if (!plan_only) {
        if (p.tma_out && out_f32) p.tma_out = encode_out_map(&tm_f32, out_f32, rows, p.N[nlayers - 1], 0);
        if (p.tma_out && out_h) p.tma_out = encode_out_map(&tm_h, out_h, rows, p.N[nlayers - 1], p.out_h_f16 ? 2 : 1);
    }
    p.stg_h_off = out_f32 ? 4096u : 0u;
    p.stg_bytes = p.tma_out ? (uint32_t)((out_f32 ? 4096 : 0) + (out_h ? 2048 : 0)) : 0u;
    cudaStream_t s = as_stream(stream);
    if (!plan_only && pool > 1 && pool != 32)
        GSPN_CUDA_OK(cudaMemsetAsync(out_f32, 0, sizeof(float) * (size_t)(rows / pool) * p.N[nlayers - 1], s));

    // Plan.  TMEM (512 columns per SM): [accumulator buffer(s) of dcols columns | activation operand hi | lo]; mid layers need their
    // whole width in one buffer (their epilogue overwrites the operand in place), the last layer runs in passes of dcols columns.
    // Shared memory: [layer-0 operand ring | weight ring | output staging | affine table].  Prefer two co-resident CTAs (a second
    // CTA's MMAs overlap this one's epilogue), and among ring depths the deepest that fits.
    const int acols = awords * (split ? 2 : 1);
    p.a_lo_off = awords;
    p.a_stage_bytes = (uint32_t)kTileBytes * (split ? 2u : 1u);
    int occ = 0;
    size_t smem = 0;
    auto plan = [&](int o, int nch) -> bool {
        const int cap = 512 / o;
        int d = dmid > nch ? dmid : nch;
        d = ((d + nch - 1) / nch) * nch;
        if (acols + d > cap) return false;
        const int epi = (o == 2 || p.mode == kModeFP) ? 4 : 8;  // FP mode: 4 epilogue + 8 gather warps share the register file
        const size_t staging = (size_t)epi * p.stg_bytes;
        const uint32_t wsb = (uint32_t)nch * 128u * (split ? 2u : 1u);
        const int tries_any[5][2] = {{3, 4}, {2, 4}, {2, 3}, {2, 2}, {1, 2}};
        const int tries_fp[5][2] = {{4, 3}, {4, 2}, {2, 4}, {2, 3}, {2, 2}};  // FP mode fills a tile's two stages together: even depths
        const int (*tries)[2] = p.mode == kModeFP ? tries_fp : tries_any;
        for (int t = 0; t < 5; ++t) {
            const size_t rings = (size_t)tries[t][0] * p.a_stage_bytes + (size_t)tries[t][1] * wsb;
            const size_t sz = 1024 + rings + staging + affine_floats * sizeof(float);
            if (sz > 226 * 1024 || (int)((228 * 1024) / (sz + 1024)) < o) continue;
            occ = o; smem = sz;
            p.nch = nch; p.dcols = d;
            p.a_stages = tries[t][0]; p.w_stages = tries[t][1]; p.w_stage_bytes = wsb;
            p.stage_off = (uint32_t)rings;  // rings end on a 1 KiB boundary (swizzle atoms)
            p.affine_off = (uint32_t)(rings + staging);
            p.epi_warps = epi;
            p.tm_bufs = (acols + 2 * d <= cap && g_tune.bufs_cap == 2) ? 2 : 1;
            return true;
        }
        return false;
    };
    const int nch0 = maxn < 128 ? maxn : 128;
    // the feature-propagation producer needs its eight gather warps' registers: one CTA per SM
    for (int o = (p.mode == kModeFP ? 1 : g_tune.occ_cap); o >= 1; --o) {
        // (a plan with ONE operand stage -- split arithmetic at two CTAs per SM: SA2, SA3, config 3 -- was A/B-ed against 64-row weight
        // stages + two operand stages: 0.070 vs 0.070 ms (SA2), 0.060 vs 0.062 ms (SA3): the co-resident CTA already covers the refill)
        if (plan(o, nch0)) break;
        if (nch0 > 64 && plan(o, 64)) break;
    }
    if (occ < 1) return GSPN_E_UNSUPPORTED;
    p.a_col = p.tm_bufs * p.dcols;
    p.tmem_cols = 32;
    while (p.tmem_cols < p.a_col + acols) p.tmem_cols <<= 1;
    // never let more CTAs co-reside than TMEM can serve: inflate the request if shared memory alone would allow it
    const size_t min_smem = (size_t)(228 * 1024) / (occ + 1) - 1024 + 1;
    if (smem < min_smem) smem = min_smem;
    // dynamic tiles: one CTA per tile; the ones the hardware has not launched yet are taken over by running CTAs (walk_next)
    p.dyn_tiles = g_tune.dyn_tiles;
    long grid = (long)di->sms * occ;
    if (grid > p.ntiles || p.dyn_tiles) grid = p.ntiles;
    if (grid > 0x7FFFFFFFL) return GSPN_E_UNSUPPORTED;
    if (plan_only) {
        const int npw = p.mode == kModeBulk ? 1 : (p.mode == kModeGatherSA ? 2 : g_tune.fp_warps);
        const int vals[12] = {occ, p.epi_warps, (p.epi_warps + npw + 2) * 32, p.nch, p.dcols, p.tm_bufs, p.tmem_cols, p.a_col, p.a_stages,
                              p.w_stages, (int)smem, (p.N[nlayers - 1] + p.dcols - 1) / p.dcols};
        for (int i = 0; i < 12; ++i) plan_out[i] = vals[i];
        return GSPN_OK;
    }
    int dev = 0;
    GSPN_CUDA_OK(cudaGetDevice(&dev));
    cudaError_t e = cudaErrorInvalidValue;
    // the opt-in shared-memory attribute is per device AND per kernel variant: raised once for each pair
    static unsigned char attr_done[64][16];  // benign race: setting it twice is harmless
    const int variant = (p.epi_warps == 8 ? 8 : 0) + (split ? 4 : 0) +
                        (p.mode == kModeBulk ? 0 : (p.mode == kModeGatherSA ? 1 : (g_tune.fp_warps == 4 ? 3 : 2)));
    const bool set_attr = !attr_done[dev][variant];
#define GSPN_LAUNCH(EPI, MINB, SP, NP) e = launch_variant<EPI, MINB, SP, NP>(p, tm_f32, tm_h, (unsigned)grid, smem, set_attr, s)
    if (p.mode == kModeFP) {
        if (g_tune.fp_warps == 4) { if (split) GSPN_LAUNCH(4, 1, true, 4); else GSPN_LAUNCH(4, 1, false, 4); }
        else { if (split) GSPN_LAUNCH(4, 1, true, 8); else GSPN_LAUNCH(4, 1, false, 8); }
    } else if (p.epi_warps == 4) {
        if (p.mode == kModeBulk) { if (split) GSPN_LAUNCH(4, 2, true, 1); else GSPN_LAUNCH(4, 2, false, 1); }
        else { if (split) GSPN_LAUNCH(4, 2, true, 2); else GSPN_LAUNCH(4, 2, false, 2); }
    } else {
        if (p.mode == kModeBulk) { if (split) GSPN_LAUNCH(8, 1, true, 1); else GSPN_LAUNCH(8, 1, false, 1); }
        else { if (split) GSPN_LAUNCH(8, 1, true, 2); else GSPN_LAUNCH(8, 1, false, 2); }
    }
#undef GSPN_LAUNCH
    if (e == cudaSuccess) attr_done[dev][variant] = 1;
    if (e != cudaSuccess) { set_last_cuda_error(e); return GSPN_E_CUDA; }
    return check_launch();
}

// What the launcher would do for a chain of this shape, without a device: mode 0 = tile image (gspn_mlp_chain), 1 = in-chain gather
// (gspn_mlp_chain_gather), 2 = feature propagation (gspn_mlp_chain_fp: dims start at the first MMA layer's input, n0).
extern "C" int gspn_mlp_chain_plan(int mode, long rows, int nlayers, const int *dims, int k0_used, int pool, int want_f32, int want_h, int arith,
                                   int *plan12) {
    GSPN_REQUIRE(mode >= 0 && mode <= 2);
    GSPN_REQUIRE_PTR(plan12);
    ChainParams p = {};
    p.mode = mode == 0 ? kModeBulk : (mode == 1 ? kModeGatherSA : kModeFP);
    return chain_launch(p, rows, nlayers, dims, k0_used, nullptr, nullptr, nullptr, nullptr, pool, want_f32 ? reinterpret_cast<float *>(64) : nullptr,
                        want_h ? reinterpret_cast<void *>(64) : nullptr, GSPN_DT_F16, arith, nullptr, plan12);
}

extern "C" int gspn_mlp_chain(long rows, int nlayers, const int *dims, int k0_used, const void *a, const void *const *wimg,
                              const float *const *scale, const float *const *shift, const int *relu, int pool, float *out_f32, void *out_h,
                              int out_h_dtype, int arith, gspn_stream_t stream) {
    if (rows > 0) GSPN_REQUIRE_PTR(a);
    ChainParams p = {};
    p.mode = kModeBulk;
    p.a = (const unsigned char *)a;
    return chain_launch(p, rows, nlayers, dims, k0_used, wimg, scale, shift, relu, pool, out_f32, out_h, out_h_dtype, arith, stream);
}

extern "C" int gspn_mlp_chain_gather(int b, int n, int m, int nsample, int c, const float *xyz, const float *new_xyz, const float *shift_pred,
                                     const float *points, const int *idx, int nlayers, const int *dims, const void *const *wimg,
                                     const float *const *scale, const float *const *shift, const int *relu, int pool, float *out_f32,
                                     void *out_h, int out_h_dtype, int arith, gspn_stream_t stream) {
    GSPN_REQUIRE(b >= 0 && n > 0 && m >= 0 && nsample > 0 && c >= 0);
    if (c + 3 > 8) return GSPN_E_UNSUPPORTED;  // one 16-byte chunk per row; wider rows go through the tile image
    if (b == 0 || m == 0) return GSPN_OK;
    GSPN_REQUIRE_PTR(xyz); GSPN_REQUIRE_PTR(new_xyz); GSPN_REQUIRE_PTR(idx); GSPN_REQUIRE_PTR(dims);
    if (c > 0) GSPN_REQUIRE_PTR(points);
    GSPN_REQUIRE(dims[0] == 64);
    ChainParams p = {};
    p.mode = kModeGatherSA;
    p.g_idx = idx; p.g_xyz = xyz; p.g_ctr = new_xyz; p.g_shift = shift_pred; p.g_pts = points;
    p.g_n = n; p.g_m = m; p.g_k = nsample; p.g_c = c;
    if ((long)b * m * nsample >= (1L << 31)) return GSPN_E_UNSUPPORTED;  // the in-kernel row -> (query, cloud) divider is 32-bit
    p.g_div_k = FastDiv((uint32_t)nsample);
    p.g_div_m = FastDiv((uint32_t)m);
    return chain_launch(p, (long)b * m * nsample, nlayers, dims, c + 3, wimg, scale, shift, relu, pool, out_f32, out_h, out_h_dtype, arith, stream);
}

extern "C" int gspn_mlp_chain_fp(int b, int n, int m, int c1, const float *y2, const int *idx, const float *weight, const float *points1,
                                 const float *w0b, int nlayers, const int *dims, const void *const *wimg, const float *const *scale,
                                 const float *const *shift, const int *relu, float *out_f32, void *out_h, int out_h_dtype, int arith,
                                 gspn_stream_t stream) {
    GSPN_REQUIRE(b >= 0 && n > 0 && m > 0 && c1 >= 0 && nlayers >= 2 && nlayers <= kMaxLayers + 1);
    if (c1 > 4) return GSPN_E_UNSUPPORTED;  // the producer keeps W0[c2:] in registers
    if (b == 0) return GSPN_OK;
    GSPN_REQUIRE_PTR(y2); GSPN_REQUIRE_PTR(idx); GSPN_REQUIRE_PTR(weight); GSPN_REQUIRE_PTR(dims); GSPN_REQUIRE_PTR(scale); GSPN_REQUIRE_PTR(shift);
    GSPN_REQUIRE_PTR(relu); GSPN_REQUIRE_PTR(wimg);
    if (c1 > 0) { GSPN_REQUIRE_PTR(points1); GSPN_REQUIRE_PTR(w0b); }
    GSPN_REQUIRE_PTR(scale[0]); GSPN_REQUIRE_PTR(shift[0]);
    const int n0 = dims[1];  // width of the producer-computed first layer = K of the first MMA layer
    if (n0 != 128) return GSPN_E_UNSUPPORTED;  // the producer maps one row of 128 columns onto a warp (fa_layer4 of the model)
    if ((long)b * n >= (1L << 31)) return GSPN_E_UNSUPPORTED;
    if ((reinterpret_cast<uintptr_t>(y2) & 15u) || (c1 > 0 && (reinterpret_cast<uintptr_t>(w0b) & 15u)) || (reinterpret_cast<uintptr_t>(scale[0]) & 15u) ||
        (reinterpret_cast<uintptr_t>(shift[0]) & 15u))
        return GSPN_E_UNSUPPORTED;
    ChainParams p = {};
    p.mode = kModeFP;
    p.f_y2 = y2; p.f_idx = idx; p.f_w = weight; p.f_p1 = points1; p.f_w0b = w0b;
    p.f_scale = scale[0]; p.f_shift = shift[0]; p.f_relu = relu[0];
    p.f_n = n; p.f_m = m; p.f_c1 = c1; p.f_n0 = n0;
    p.f_div_n = FastDiv((uint32_t)n);
    return chain_launch(p, (long)b * n, nlayers - 1, dims + 1, 0, wimg + 1, scale + 1, shift + 1, relu + 1, 1, out_f32, out_h, out_h_dtype, arith,
                        stream);
}

extern "C" int gspn_fp_assemble(int b, int n, int m, int c1, int c2, const float *points1, const float *points2, const int *idx,
                                const float *weight, void *a_img, int ld, int image_dtype, gspn_stream_t stream) {
    GSPN_REQUIRE(b >= 0 && n > 0 && c1 >= 0 && c2 >= 0 && c1 + c2 > 0 && ld % 64 == 0 && ld >= c1 + c2);
    if (image_dtype != GSPN_DT_BF16 && image_dtype != GSPN_DT_BF16X2) return GSPN_E_BAD_DTYPE;
    if (b == 0) return GSPN_OK;
    GSPN_REQUIRE_PTR(a_img);
    if (c2 > 0) { GSPN_REQUIRE(m > 0); GSPN_REQUIRE_PTR(points2); GSPN_REQUIRE_PTR(idx); GSPN_REQUIRE_PTR(weight); }
    if (c1 > 0) GSPN_REQUIRE_PTR(points1);
    long rows = (long)b * n;
    long total = rows * (ld / 8);
    long blk = ceil_div_l(total, 256);
    if (blk > 148L * 64) blk = 148L * 64;
    fp_assemble_kernel<<<(unsigned)blk, 256, 0, as_stream(stream)>>>(rows, n, m, c1, c2, points1, points2, idx, weight, (unsigned char *)a_img, ld,
                                                                     FastDiv((uint32_t)(ld >> 3)), FastDiv((uint32_t)n),
                                                                     image_dtype == GSPN_DT_BF16X2);
    return check_launch();
}
