// fps_bucket.cu -- exact bucket-pruned farthest point sampling for clouds of 8193 .. 32768 points (SA1 of the model:
// 32768 -> 2048), ONE CTA per cloud.  Same results, bit for bit, as farthestpointsamplingKernel (tf_sampling_g.cu:105-170) and as
// the cluster kernels in fps.cu; what changes is the work per round and the footprint.
//
// Every round of FPS lowers the running min-distance of the points NEAR the new sample and leaves everything else untouched.
// The cloud is therefore sorted once into spatial buckets of 32 points (Hilbert-curve order of a 32^3 grid over the bounding box, a
// counting sort inside the kernel's prologue), every bucket keeps its bounding box and its current (max distance, tie-break key,
// coordinates of that point), and a round only recomputes the buckets that can change:
//     bound = the reference's own float distance expression evaluated on the per-axis gaps between the sample and the box
//           <= the float distance of the sample to every point in the box   (each operation is monotone in |operand|)
//     bound >= max distance in the bucket  =>  min(d, d_new) = d for all its points: skip, bit-exactly.
// About 39 000 bucket updates replace the 2047 x 1024 of the full scan (config 2, measured), so one SM does the work of the 8-CTA
// cluster -- and without the cluster there is no DSMEM exchange on the round's critical path: one __syncthreads and two warp
// reductions.  Buckets are dealt to the 12 warps round-robin in curve order: the buckets one sample touches are neighbours in
// that order and must not queue up in one warp.
//
// A 32768-point cloud needs 512 KB on chip (x, y, z, distance).  One SM has it, in three places:
//     x and the original index (uint16) : shared memory        (128 KB + 64 KB)
//     y, z                              : TENSOR MEMORY, all 512 columns, used as plain scratch through tcgen05.st / tcgen05.ld
//                                         (33 cycles round trip, tools/probes/ts_mma_probe.cu) -- lane = point, column = bucket
//     running min-distances             : registers, 86 per thread (thread (w, l) owns point l of each of warp w's 86 buckets;
//                                         12 warps x 170 registers: 32768 of the SM's 65536 registers hold distances)
// A bucket is processed by a warp with lane = point, so the dynamic part of the address (which bucket) is a shared-memory row /
// a TMEM column, and the register array is indexed statically by an unrolled, hierarchically skipped loop over the warp's slots.
// Taking every TMEM column (and 192 KB of shared memory) also keeps the tensor-core chain kernels off this SM, by construction.
//
// Tie-break: the reference's winner among equal maxima is the lowest (k mod 512, k) of the ORIGINAL index k (fps.cu); the key is
// computed from the index each sorted point carries, so the bucket order is irrelevant to the result -- also when the counting
// sort's atomics place equal-cell points in a different order from run to run.
#include "common.cuh"

namespace gspn {

#ifndef GSPN_FB_WARPS
#define GSPN_FB_WARPS 16
#endif
// 16 warps x 64 slots (128 registers per thread) or 12 warps x 86 slots (170 registers): 32768 of the SM's 65536 registers hold
// distances either way; what is left per thread decides whether the round loop compiles without local-memory spills, which with
// 220 KB of shared memory carved out have no L1 to land in.
constexpr int kFbWarps = GSPN_FB_WARPS, kFbThreads = kFbWarps * 32, kFbSlots = (1024 + kFbWarps - 1) / kFbWarps;  // slots = buckets per warp
constexpr int kFbQuadWarps = kFbWarps / 4;                               // warps sharing a TMEM lane quadrant
constexpr int kFbBuckets = 1024;                                         // 12 x 86 = 1032 slots: the last 8 stay empty
constexpr int kFbMaxPoints = kFbBuckets * 32;                            // 32768
constexpr int kFbSlotsAll = kFbWarps * kFbSlots;                         // 1032
constexpr int kFbCells = 32 * 32 * 32;
constexpr int kFbCellsPerThread = (kFbCells + kFbThreads - 1) / kFbThreads;
// shared memory: x (fp32) + original index (uint16) of every sorted point, then per slot: box (6 floats) and key of the maximum
constexpr size_t kFbSmem = (size_t)kFbMaxPoints * 6 + (size_t)kFbSlotsAll * 7 * 4;

__device__ __forceinline__ unsigned fb_key(int k) { return ((unsigned)(k & 511) << 23) | ((unsigned)k >> 9); }
__device__ __forceinline__ int fb_unkey(unsigned key) { return (int)(((key & 0x7FFFFFu) << 9) | (key >> 23)); }
// float <-> int with the same order (for redux.sync min / max on coordinates of either sign)
__device__ __forceinline__ int fb_ord(float f) { int b = __float_as_int(f); return b ^ ((b >> 31) & 0x7FFFFFFF); }
__device__ __forceinline__ float fb_unord(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7FFFFFFF)); }
__device__ __forceinline__ unsigned fb_spread5(unsigned v) {  // bit i -> bit 3i
    return (v & 1u) | ((v & 2u) << 2) | ((v & 4u) << 4) | ((v & 8u) << 6) | ((v & 16u) << 8);
}
__device__ __forceinline__ uint32_t fb_s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fb_tld1(uint32_t taddr, uint32_t &v) { asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr)); }
__device__ __forceinline__ void fb_tst1(uint32_t taddr, uint32_t v) { asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v) : "memory"); }
__device__ __forceinline__ void fb_twait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void fb_twait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

struct __align__(16) FbSlot { float x, y, z; int dbits; };

// the reference's distance expression on the gaps between a sample and a box: a lower bound, in the same float arithmetic, of
// the distance from the sample to any point inside the box
__device__ __forceinline__ float fb_box_bound(float lx, float ly, float lz, float hx, float hy, float hz, float sx, float sy, float sz) {
    const float gx = fmaxf(fmaxf(__fsub_rn(lx, sx), __fsub_rn(sx, hx)), 0.f);
    const float gy = fmaxf(fmaxf(__fsub_rn(ly, sy), __fsub_rn(sy, hy)), 0.f);
    const float gz = fmaxf(fmaxf(__fsub_rn(lz, sz), __fsub_rn(sz, hz)), 0.f);
    float t = __fmul_rn(gy, gy);
    t = __fmaf_rn(gx, gx, t);
    return __fmaf_rn(gz, gz, t);
}

// slot r of warp w holds bucket j = r * 12 + w of the curve order (sorted points 32 j .. 32 j + 31): the buckets a sample touches
// are neighbours in that order, so they spread over all warps instead of queueing up in one (measured: 2.7x per round)
template <bool PROF>
__global__ void __launch_bounds__(kFbThreads, 1) fps_bucket_kernel(int n, int m, const float *__restrict__ xyz, int *__restrict__ out,
                                                                   float4 *__restrict__ sorted, long long *__restrict__ prof) {
    extern __shared__ __align__(16) unsigned char fb_smem[];
    float *X = reinterpret_cast<float *>(fb_smem);
    unsigned short *IDX = reinterpret_cast<unsigned short *>(fb_smem + (size_t)kFbMaxPoints * 4);
    float *BOX = reinterpret_cast<float *>(fb_smem + (size_t)kFbMaxPoints * 6);  // [6][kFbSlotsAll]: lo x,y,z, hi x,y,z of every slot's bucket
    unsigned *BKEY = reinterpret_cast<unsigned *>(BOX + 6 * kFbSlotsAll);        // [kFbSlotsAll]: tie-break key of every bucket's maximum
    int *cellcnt = reinterpret_cast<int *>(fb_smem);  // prologue only: aliases X (the sort is finished before X is filled)
    __shared__ uint32_t tmem_slot;
    __shared__ float red[kFbWarps][6];
    __shared__ int wsum[kFbWarps];
    __shared__ FbSlot tslot[2][16];
    __shared__ unsigned tkey[2][16];
    __shared__ FbSlot cand[kFbWarps];   // each warp's current candidate (rewritten only when one of its bucket maxima moved)
    __shared__ unsigned candkey[kFbWarps];

    const int cloud = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *p = xyz + (size_t)cloud * n * 3;
    float4 *srt = sorted + (size_t)cloud * kFbMaxPoints;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(fb_s32(&tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid < 2 * 16) {  // table entries no warp owns never win
        (&tslot[0][0])[tid] = FbSlot{0.f, 0.f, 0.f, __float_as_int(-1.0f)};
        (&tkey[0][0])[tid] = 0xFFFFFFFFu;
    }
    // ---- P1: bounding box of the cloud
    float lo0 = 3.4e38f, lo1 = 3.4e38f, lo2 = 3.4e38f, hi0 = -3.4e38f, hi1 = -3.4e38f, hi2 = -3.4e38f;
    for (int k = tid; k < n; k += kFbThreads) {
        const float x = __ldg(p + 3 * k), y = __ldg(p + 3 * k + 1), z = __ldg(p + 3 * k + 2);
        lo0 = fminf(lo0, x); hi0 = fmaxf(hi0, x); lo1 = fminf(lo1, y); hi1 = fmaxf(hi1, y); lo2 = fminf(lo2, z); hi2 = fmaxf(hi2, z);
    }
    lo0 = fb_unord(__reduce_min_sync(GSPN_FULL_MASK, fb_ord(lo0))); hi0 = fb_unord(__reduce_max_sync(GSPN_FULL_MASK, fb_ord(hi0)));
    lo1 = fb_unord(__reduce_min_sync(GSPN_FULL_MASK, fb_ord(lo1))); hi1 = fb_unord(__reduce_max_sync(GSPN_FULL_MASK, fb_ord(hi1)));
    lo2 = fb_unord(__reduce_min_sync(GSPN_FULL_MASK, fb_ord(lo2))); hi2 = fb_unord(__reduce_max_sync(GSPN_FULL_MASK, fb_ord(hi2)));
    if (lane == 0) { red[warp][0] = lo0; red[warp][1] = lo1; red[warp][2] = lo2; red[warp][3] = hi0; red[warp][4] = hi1; red[warp][5] = hi2; }
    for (int i = tid; i < kFbCells; i += kFbThreads) cellcnt[i] = 0;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    for (int w = 0; w < kFbWarps; ++w) {
        lo0 = fminf(lo0, red[w][0]); lo1 = fminf(lo1, red[w][1]); lo2 = fminf(lo2, red[w][2]);
        hi0 = fmaxf(hi0, red[w][3]); hi1 = fmaxf(hi1, red[w][4]); hi2 = fmaxf(hi2, red[w][5]);
    }
    const float e0 = hi0 - lo0, e1 = hi1 - lo1, e2 = hi2 - lo2;
    const float inv0 = (e0 > 0.f && e0 < 3e38f) ? 32.f / e0 : 0.f, inv1 = (e1 > 0.f && e1 < 3e38f) ? 32.f / e1 : 0.f,
                inv2 = (e2 > 0.f && e2 < 3e38f) ? 32.f / e2 : 0.f;
    // cell of a point on a 32^3 grid, numbered along a Hilbert curve (Skilling's transpose form): consecutive cells are always
    // face neighbours, so 32 consecutive points of the sorted cloud form a compact bucket (a Morton order jumps across the cloud at
    // every power-of-two boundary, and buckets that straddle a jump have boxes that every sample touches)
    auto cell_of = [=](float x, float y, float z) -> unsigned {
        unsigned X0 = (unsigned)min(31, max(0, (int)((x - lo0) * inv0))), X1 = (unsigned)min(31, max(0, (int)((y - lo1) * inv1))),
                 X2 = (unsigned)min(31, max(0, (int)((z - lo2) * inv2)));
#pragma unroll
        for (unsigned Q = 16; Q > 1; Q >>= 1) {
            const unsigned P = Q - 1;
            if (X0 & Q) X0 ^= P;  // i = 0: invert (the exchange with itself is a no-op)
            if (X1 & Q) X0 ^= P; else { const unsigned t = (X0 ^ X1) & P; X0 ^= t; X1 ^= t; }
            if (X2 & Q) X0 ^= P; else { const unsigned t = (X0 ^ X2) & P; X0 ^= t; X2 ^= t; }
        }
        X1 ^= X0; X2 ^= X1;
        unsigned t = 0;
#pragma unroll
        for (unsigned Q = 16; Q > 1; Q >>= 1)
            if (X2 & Q) t ^= Q - 1;
        X0 ^= t; X1 ^= t; X2 ^= t;
        return (fb_spread5(X0) << 2) | (fb_spread5(X1) << 1) | fb_spread5(X2);
    };
    // ---- P2: histogram of the cells
    for (int k = tid; k < n; k += kFbThreads) atomicAdd(&cellcnt[cell_of(__ldg(p + 3 * k), __ldg(p + 3 * k + 1), __ldg(p + 3 * k + 2))], 1);
    __syncthreads();
    // ---- P3: exclusive scan (kFbCellsPerThread consecutive cells per thread)
    {
        const int c_lo = min(kFbCells, tid * kFbCellsPerThread), c_hi = min(kFbCells, c_lo + kFbCellsPerThread);
        int s = 0;
        for (int i = c_lo; i < c_hi; ++i) s += cellcnt[i];
        int incl = s;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(GSPN_FULL_MASK, incl, d);
            if (lane >= d) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int base = incl - s;
        for (int w = 0; w < warp; ++w) base += wsum[w];
        for (int i = c_lo; i < c_hi; ++i) { const int c = cellcnt[i]; cellcnt[i] = base; base += c; }
    }
    __syncthreads();
    // ---- P4: scatter (x, y, z, original index) into curve order
    for (int k = tid; k < n; k += kFbThreads) {
        const float x = __ldg(p + 3 * k), y = __ldg(p + 3 * k + 1), z = __ldg(p + 3 * k + 2);
        const int pos = atomicAdd(&cellcnt[cell_of(x, y, z)], 1);
        srt[pos] = make_float4(x, y, z, __int_as_float(k));
    }
    __syncthreads();  // the sort is complete and visible to the CTA; cellcnt is dead: X may be written
    // ---- P5: resident copy.  Thread (warp, lane) owns point `lane` of the bucket in each of its warp's slots.
    // TMEM column of slot r: kFbQuadWarps * r + (warp / 4) for y, + 256 for z -- the warps sharing a lane quadrant interleave
    // their slots (256 columns per array).
    const uint32_t tY = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(warp >> 2);
    float d[kFbSlots];
#pragma unroll
    for (int r = 0; r < kFbSlots; ++r) {
        const int j = r * kFbWarps + warp;  // bucket
        const int pos = j * 32 + lane;
        const bool ok = j < kFbBuckets && pos < n;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ok) v = __ldcg(srt + pos);
        if (j < kFbBuckets) {
            X[pos] = v.x;
            IDX[pos] = (unsigned short)__float_as_int(v.w);
            fb_tst1(tY + kFbQuadWarps * r, __float_as_uint(v.y));
            fb_tst1(tY + 256 + kFbQuadWarps * r, __float_as_uint(v.z));
        }
        d[r] = ok ? 1e38f : -1.0f;  // tf_sampling_g.cu:118; padding never wins (distances are >= 0)
        const int ilx = __reduce_min_sync(GSPN_FULL_MASK, ok ? fb_ord(v.x) : 0x7FFFFFFF), ihx = __reduce_max_sync(GSPN_FULL_MASK, ok ? fb_ord(v.x) : (int)0x80000000);
        const int ily = __reduce_min_sync(GSPN_FULL_MASK, ok ? fb_ord(v.y) : 0x7FFFFFFF), ihy = __reduce_max_sync(GSPN_FULL_MASK, ok ? fb_ord(v.y) : (int)0x80000000);
        const int ilz = __reduce_min_sync(GSPN_FULL_MASK, ok ? fb_ord(v.z) : 0x7FFFFFFF), ihz = __reduce_max_sync(GSPN_FULL_MASK, ok ? fb_ord(v.z) : (int)0x80000000);
        if (lane == 0) {  // an empty bucket keeps an inverted box (never tested: its max distance is -1)
            float *bb = BOX + warp * kFbSlots + r;
            bb[0] = fb_unord(ilx); bb[kFbSlotsAll] = fb_unord(ily); bb[2 * kFbSlotsAll] = fb_unord(ilz);
            bb[3 * kFbSlotsAll] = fb_unord(ihx); bb[4 * kFbSlotsAll] = fb_unord(ihy); bb[5 * kFbSlotsAll] = fb_unord(ihz);
        }
    }
    fb_twait_st();
    __syncwarp();
    // ---- P6: bucket state.  Lane l owns slots l, 32 + l, 64 + l (< 86) of its warp: the max distance lives in a register, the key
    // and the box in shared memory.
    int bmax0, bmax1, bmax2;  // float bits of each owned bucket's max distance (-1.0f: empty bucket, never touched)
    unsigned iswin0 = 0, iswin1 = 0, iswin2 = 0;  // bit L of iswinS: THIS lane's point holds the maximum of slot 32 S + L
    const float *mybox = BOX + warp * kFbSlots + lane;
    unsigned *mykey = BKEY + warp * kFbSlots + lane;
    {
        auto nonempty = [&](int r) { return r < kFbSlots && (r * kFbWarps + warp) < kFbBuckets && (r * kFbWarps + warp) * 32 < n; };
        bmax0 = __float_as_int(nonempty(lane) ? 1e38f : -1.0f);
        bmax1 = __float_as_int(nonempty(32 + lane) ? 1e38f : -1.0f);
        bmax2 = __float_as_int(nonempty(64 + lane) ? 1e38f : -1.0f);
        mykey[0] = 0xFFFFFFFFu; mykey[32] = 0xFFFFFFFFu;
        if (64 + lane < kFbSlots) mykey[64] = 0xFFFFFFFFu;
        if (kFbSlots <= 64) bmax2 = __float_as_int(-1.0f);
    }
    // the warp's cached candidate (uniform across lanes): max distance bits, key, coordinates
    if (lane == 0) {
        cand[warp] = FbSlot{0.f, 0.f, 0.f, __float_as_int((warp * 32 < n) ? 1e38f : -1.0f)};
        candkey[warp] = 0xFFFFFFFFu;
    }
    __syncwarp();
    float sx = __ldg(p), sy = __ldg(p + 1), sz = __ldg(p + 2);  // old = 0 (:114)
    int *o = out + (size_t)cloud * m;
    if (tid == 0) o[0] = 0;
    unsigned acc_touched = 0, acc_full = 0, t_start = 0, ph0 = 0, ph1 = 0, ph2 = 0, ph3 = 0, ph4 = 0, c0 = 0, c1 = 0, c2 = 0, c3 = 0, c4 = 0;  // PROF only
    if (PROF) t_start = (unsigned)clock();
    const bool own2 = kFbSlots > 64 && 64 + lane < kFbSlots;

    for (int r = 1; r < m; ++r) {
        const int par = r & 1;
        if (PROF) c0 = (unsigned)clock();
        {
            // one box at a time (the empty asm keeps the compiler from hoisting all 18 loads: registers are what this kernel is short of)
            const bool t0 = fb_box_bound(mybox[0], mybox[kFbSlotsAll], mybox[2 * kFbSlotsAll], mybox[3 * kFbSlotsAll], mybox[4 * kFbSlotsAll],
                                         mybox[5 * kFbSlotsAll], sx, sy, sz) < __int_as_float(bmax0);
            asm volatile("" ::: "memory");
            const bool t1 = fb_box_bound(mybox[32], mybox[kFbSlotsAll + 32], mybox[2 * kFbSlotsAll + 32], mybox[3 * kFbSlotsAll + 32],
                                         mybox[4 * kFbSlotsAll + 32], mybox[5 * kFbSlotsAll + 32], sx, sy, sz) < __int_as_float(bmax1);
            bool t2 = false;
            if constexpr (kFbSlots > 64) {
                asm volatile("" ::: "memory");
                if (own2)
                    t2 = fb_box_bound(mybox[64], mybox[kFbSlotsAll + 64], mybox[2 * kFbSlotsAll + 64], mybox[3 * kFbSlotsAll + 64],
                                      mybox[4 * kFbSlotsAll + 64], mybox[5 * kFbSlotsAll + 64], sx, sy, sz) < __int_as_float(bmax2);
            }
            // NOT an array: a runtime-indexed mask[S] lives in local memory, and with 220 KB of shared memory no L1 is left to catch it
            const unsigned mask0 = __ballot_sync(GSPN_FULL_MASK, t0), mask1 = __ballot_sync(GSPN_FULL_MASK, t1),
                           mask2 = kFbSlots > 64 ? __ballot_sync(GSPN_FULL_MASK, t2) : 0u;
            if (PROF) { acc_touched += __popc(mask0) + __popc(mask1) + __popc(mask2); c1 = (unsigned)clock(); c2 = c1; }
            if (mask0 | mask1 | mask2) {
                bool dirty = false;
                // ONE copy of the bucket update, looped over the touched slots: everything that depends on the slot is an address
                // (TMEM column, shared-memory row) except the distance register, which a jump table of one-instruction cases
                // picks (R is warp-uniform).  Unrolling the body per slot instead costs 100 KB of code and an instruction-cache
                // miss chain per touched bucket.
#pragma unroll 1
                for (int S = 0; S < (kFbSlots > 64 ? 3 : 2); ++S) {
                    unsigned mk = S == 0 ? mask0 : (S == 1 ? mask1 : mask2);
#pragma unroll 1
                    while (mk) {
                        const int L = __ffs(mk) - 1;  // owner lane of this bucket
                        mk &= mk - 1;
                        const int R = 32 * S + L;
                        uint32_t yb, zb;
                        fb_tld1(tY + kFbQuadWarps * R, yb);
                        fb_tld1(tY + 256 + kFbQuadWarps * R, zb);
                        const int pos = (R * kFbWarps + warp) * 32 + lane;
                        const float x = X[pos];
                        const int oi = IDX[pos];
                        fb_twait_ld();
                        const float dn = sqdist_fma(x, __uint_as_float(yb), __uint_as_float(zb), sx, sy, sz);
                        float dd, dold;
                        switch (R) {
#define FB_CASE(i) case i: dold = d[i]; dd = fminf(dn, dold); d[i] = dd; break;
#define FB_CASE8(i) FB_CASE(i) FB_CASE(i + 1) FB_CASE(i + 2) FB_CASE(i + 3) FB_CASE(i + 4) FB_CASE(i + 5) FB_CASE(i + 6) FB_CASE(i + 7)
                            FB_CASE8(0) FB_CASE8(8) FB_CASE8(16) FB_CASE8(24) FB_CASE8(32) FB_CASE8(40) FB_CASE8(48) FB_CASE8(56)
#if GSPN_FB_WARPS == 12
                            FB_CASE8(64) FB_CASE8(72) FB_CASE(80) FB_CASE(81) FB_CASE(82) FB_CASE(83) FB_CASE(84) FB_CASE(85)
#endif
#undef FB_CASE8
#undef FB_CASE
                            default: dd = dold = dn; break;
                        }
                        // Distances only fall, so no other point can newly reach the bucket's maximum: if the point that holds it
                        // did not move, (max, key, holder) are what they were -- one ballot instead of the argmax below.
                        const unsigned iw = S == 0 ? iswin0 : (S == 1 ? iswin1 : iswin2);
                        const bool holder = (iw >> L) & 1u;
                        if (__ballot_sync(GSPN_FULL_MASK, holder && dd != dold) == 0u && r > 1) continue;
                        if (PROF) ++acc_full;
                        const int db = __float_as_int(dd);
                        const int bm = __reduce_max_sync(GSPN_FULL_MASK, db);  // non-negative floats order as ints; -1.0f is negative
                        const unsigned kk = (db == bm) ? fb_key(oi) : 0xFFFFFFFFu;
                        const unsigned bk = __reduce_min_sync(GSPN_FULL_MASK, kk);
                        const unsigned src = (unsigned)__ffs(__ballot_sync(GSPN_FULL_MASK, kk == bk)) - 1u;
                        const unsigned nw = (iw & ~(1u << L)) | ((lane == (int)src ? 1u : 0u) << L);
                        if (S == 0) iswin0 = nw; else if (S == 1) iswin1 = nw; else iswin2 = nw;
                        if (lane == L) {
                            if (S == 0) bmax0 = bm; else if (S == 1) bmax1 = bm; else bmax2 = bm;
                            mykey[32 * S] = bk;
                        }
                        dirty = true;
                    }
                }
                if (PROF) c2 = (unsigned)clock();
                if (dirty) {
                    // ---- the warp's candidate again: best of its bucket maxima
                    __syncwarp();
                    int cb = bmax0, cs = 0;
                    unsigned ck = mykey[0];
                    {
                        const unsigned k1 = mykey[32];
                        if (bmax1 > cb || (bmax1 == cb && k1 < ck)) { cb = bmax1; ck = k1; cs = 32; }
                        if (own2) {
                            const unsigned k2 = mykey[64];
                            if (bmax2 > cb || (bmax2 == cb && k2 < ck)) { cb = bmax2; ck = k2; cs = 64; }
                        }
                    }
                    const int wmax = __reduce_max_sync(GSPN_FULL_MASK, cb);
                    const unsigned kk = (cb == wmax) ? ck : 0xFFFFFFFFu;
                    const unsigned wkey = __reduce_min_sync(GSPN_FULL_MASK, kk);
                    const int src = __ffs(__ballot_sync(GSPN_FULL_MASK, kk == wkey)) - 1;
                    // its coordinates: the slot from the owning lane, the holder from the flags, then one shared-memory / two TMEM reads
                    const int rs = __shfl_sync(GSPN_FULL_MASK, cs + lane, src);
                    const unsigned iw = rs < 32 ? iswin0 : (rs < 64 ? iswin1 : iswin2);
                    const int ls = __ffs(__ballot_sync(GSPN_FULL_MASK, (iw >> (rs & 31)) & 1u)) - 1;
                    uint32_t yb, zb;
                    fb_tld1(tY + kFbQuadWarps * rs, yb);
                    fb_tld1(tY + 256 + kFbQuadWarps * rs, zb);
                    const float wx = X[(rs * kFbWarps + warp) * 32 + ls];
                    fb_twait_ld();
                    if (lane == ls) {
                        cand[warp] = FbSlot{wx, __uint_as_float(yb), __uint_as_float(zb), wmax};
                        candkey[warp] = wkey;
                    }
                    __syncwarp();
                }
            }
        }
        if (PROF) c3 = (unsigned)clock();
        if (lane == 0) {
            tslot[par][warp] = cand[warp];
            tkey[par][warp] = candkey[warp];
        }
        __syncthreads();
        if (PROF) c4 = (unsigned)clock();
        // ---- every warp reduces the candidates (entries 12..15 are empty)
        int cb = __float_as_int(-1.0f);
        unsigned ck = 0xFFFFFFFFu;
        float cx = 0.f, cy = 0.f, cz = 0.f;
        if (lane < 16) {
            const FbSlot s = tslot[par][lane];
            cb = s.dbits; ck = tkey[par][lane]; cx = s.x; cy = s.y; cz = s.z;
        }
        const int gm = __reduce_max_sync(GSPN_FULL_MASK, cb);
        const unsigned kk = (cb == gm) ? ck : 0xFFFFFFFFu;
        const unsigned gk = __reduce_min_sync(GSPN_FULL_MASK, kk);
        const int src = __ffs(__ballot_sync(GSPN_FULL_MASK, kk == gk)) - 1;
        sx = __shfl_sync(GSPN_FULL_MASK, cx, src);
        sy = __shfl_sync(GSPN_FULL_MASK, cy, src);
        sz = __shfl_sync(GSPN_FULL_MASK, cz, src);
        if (tid == 0) o[r] = fb_unkey(gk);
        if (PROF) {
            const unsigned c5 = (unsigned)clock();
            ph0 += c1 - c0; ph1 += c2 - c1; ph2 += c3 - c2; ph3 += c4 - c3; ph4 += c5 - c4;
        }
    }
    if (PROF) {
        // per cloud 0: [0] cycles of the round loop, [1] rounds, [2] bucket updates (all warps), [3..7] warp 0's cycles in box tests,
        // bucket updates, warp argmax, barrier wait, table reduce, [8] bucket updates that needed the full argmax (all warps)
        const unsigned t_end = (unsigned)clock();
        if (cloud == 0) {
            if (tid == 0) {
                prof[0] = t_end - t_start; prof[1] = m - 1;
                prof[3] = ph0; prof[4] = ph1; prof[5] = ph2; prof[6] = ph3; prof[7] = ph4;
            }
            if (lane == 0) {
                atomicAdd((unsigned long long *)prof + 2, (unsigned long long)acc_touched);
                atomicAdd((unsigned long long *)prof + 8, (unsigned long long)acc_full);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
}

}  // namespace gspn

using namespace gspn;

// used by fps.cu's entry points
size_t gspn_fps_bucket_workspace_bytes(int b, int n) {
    if (b <= 0 || n <= 8192 || n > kFbMaxPoints) return 0;
    return (size_t)b * kFbMaxPoints * sizeof(float4);
}

int gspn_fps_bucket_launch(int b, int n, int m, const float *inp, int *out, void *workspace, long long *prof, cudaStream_t s) {
    static unsigned char attr_done[64];  // per device; benign race
    int dev = 0;
    GSPN_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return GSPN_E_UNSUPPORTED;
    if (!attr_done[dev]) {
        GSPN_CUDA_OK(cudaFuncSetAttribute(fps_bucket_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFbSmem));
        GSPN_CUDA_OK(cudaFuncSetAttribute(fps_bucket_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFbSmem));
        attr_done[dev] = 1;
    }
    if (prof) fps_bucket_kernel<true><<<b, kFbThreads, kFbSmem, s>>>(n, m, inp, out, (float4 *)workspace, prof);
    else fps_bucket_kernel<false><<<b, kFbThreads, kFbSmem, s>>>(n, m, inp, out, (float4 *)workspace, prof);
    return check_launch();
}
