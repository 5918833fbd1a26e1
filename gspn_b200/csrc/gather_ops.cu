// gather_ops.cu -- the HBM-bound row movers of the SA/FP path: gather_point, group_point,
// three_interpolate and the scatter-add backward ops.  One thread per 16-byte vector of the
// output where channels allow it (coalesced 128-bit stores, gathered 128-bit loads); grids are
// sized from the element count, not a fixed <<<b,256>>> like the reference.
#include "common.cuh"

namespace gspn {

constexpr int kThreads = 256;

static inline int blocks_for(long work) {
    long blk = ceil_div_l(work, kThreads);
    // grid-stride beyond ~32 waves of 148 SMs x 8 CTAs
    const long cap = 148L * 8 * 32;
    return (int)(blk < cap ? (blk > 0 ? blk : 1) : cap);
}

// ---- gather_point: out[b,j,:] = inp[b,idx[b,j],:]   (tf_sampling_g.cu:172-181) ----------
template <typename V>
__global__ void __launch_bounds__(kThreads) gather_rows_kernel(long total, int n, int m, int cv /* vectors per row */,
                                                               const V *__restrict__ inp, const int *__restrict__ idx, V *__restrict__ out) {
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        long row = e / cv;  // b*m + j
        int l = (int)(e - row * cv);
        long bi = row / m;
        int a = __ldg(idx + row);
        out[e] = __ldg(inp + (bi * n + a) * cv + l);
    }
}

// ---- scatter-add of rows: dst[b,idx[b,j],:] += src[b,j,:]   (tf_sampling_g.cu:183-192) ----
__global__ void __launch_bounds__(kThreads) scatter_rows_kernel(long total, int n, int m, int c, const float *__restrict__ src,
                                                                const int *__restrict__ idx, float *__restrict__ dst) {
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        long row = e / c;
        int l = (int)(e - row * c);
        long bi = row / m;
        int a = __ldg(idx + row);
        atomicAdd(dst + (bi * n + a) * c + l, __ldg(src + e));
    }
}
__global__ void __launch_bounds__(kThreads) scatter_rows_v4_kernel(long total, int n, int m, int cv, const float4 *__restrict__ src,
                                                                   const int *__restrict__ idx, float4 *__restrict__ dst) {
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        long row = e / cv;
        int l = (int)(e - row * cv);
        long bi = row / m;
        int a = __ldg(idx + row);
        atomicAdd(dst + (bi * n + a) * cv + l, __ldg(src + e));  // red.global.add.v4.f32 (sm_90+)
    }
}

// ---- three_interpolate: out[j,l] = (p[i1,l]*w1 + p[i2,l]*w2) + p[i3,l]*w3, no FMA
// (tf_interpolate.cpp:107-127) -------------------------------------------------------------
__device__ __forceinline__ float interp3(float a, float b, float c, float w1, float w2, float w3) {
    return __fadd_rn(__fadd_rn(__fmul_rn(a, w1), __fmul_rn(b, w2)), __fmul_rn(c, w3));
}

__global__ void __launch_bounds__(kThreads) three_interpolate_v4_kernel(long total, int m, int n, int cv, const float4 *__restrict__ points,
                                                                        const int *__restrict__ idx, const float *__restrict__ weight,
                                                                        float4 *__restrict__ out) {
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        long row = e / cv;  // b*n + j
        int l = (int)(e - row * cv);
        long bi = row / n;
        const int *ip = idx + row * 3;
        const float *wp = weight + row * 3;
        int i1 = __ldg(ip), i2 = __ldg(ip + 1), i3 = __ldg(ip + 2);
        float w1 = __ldg(wp), w2 = __ldg(wp + 1), w3 = __ldg(wp + 2);
        const float4 *base = points + bi * m * cv + l;
        float4 a = __ldg(base + (long)i1 * cv), b = __ldg(base + (long)i2 * cv), c = __ldg(base + (long)i3 * cv);
        float4 o;
        o.x = interp3(a.x, b.x, c.x, w1, w2, w3);
        o.y = interp3(a.y, b.y, c.y, w1, w2, w3);
        o.z = interp3(a.z, b.z, c.z, w1, w2, w3);
        o.w = interp3(a.w, b.w, c.w, w1, w2, w3);
        __stcs(out + e, o);  // streaming store: the interpolated map is read once by the MLP
    }
}

// The streaming form: a group of G = c/8 lanes (8, 16 or 32) owns an output row; ONE lane of the group loads the row's three
// indices and weights and the group shares them through shuffles (the v4 kernel re-loads all six values in every thread), every lane
// moves 32 bytes per access (LDG.E.256 gathers, a streaming STG.E.256), and each group keeps two rows (six 256-bit gathers per lane)
// in flight.  Rounding as the CPU op: (p1*w1 + p2*w2) + p3*w3 without FMA.
struct __align__(32) F8 { float v[8]; };
__device__ __forceinline__ F8 ldg_f8(const float *p) {
    F8 r;
    asm volatile("ld.global.nc.L1::evict_last.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stcs_f8(float *p, const F8 &r) {
    asm volatile("st.global.cs.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]),
                 "f"(r.v[4]), "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7])
                 : "memory");
}
template <int G>
__global__ void __launch_bounds__(kThreads) three_interpolate_rows_kernel(long rows, int m, int n, const float *__restrict__ points,
                                                                          const int *__restrict__ idx, const float *__restrict__ weight,
                                                                          float *__restrict__ out, const FastDiv div_n) {
    constexpr int C = G * 8;
    constexpr int RPW = 32 / G;  // rows per warp-step
    const int lane = threadIdx.x & 31, g = lane / G, l = lane % G;
    const long warp = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5, nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long r0 = warp * (2 * RPW); r0 < rows; r0 += nwarps * (2 * RPW)) {
        F8 a[2], b[2], c[2];
        float w1[2], w2[2], w3[2];
        long row[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            row[u] = r0 + u * RPW + g;
            const bool ok = row[u] < rows;
            const long rr = ok ? row[u] : 0;
            int iv = 0;
            float wv = 0.f;
            if (l < 3) { iv = __ldg(idx + rr * 3 + l); wv = __ldg(weight + rr * 3 + l); }  // lanes 0..2 of the group: one value each
            const int i1 = __shfl_sync(GSPN_FULL_MASK, iv, g * G), i2 = __shfl_sync(GSPN_FULL_MASK, iv, g * G + 1),
                      i3 = __shfl_sync(GSPN_FULL_MASK, iv, g * G + 2);
            w1[u] = __shfl_sync(GSPN_FULL_MASK, wv, g * G); w2[u] = __shfl_sync(GSPN_FULL_MASK, wv, g * G + 1);
            w3[u] = __shfl_sync(GSPN_FULL_MASK, wv, g * G + 2);
            const float *base = points + (size_t)div_n.div((uint32_t)rr) * m * C + l * 8;
            a[u] = ldg_f8(base + (size_t)i1 * C); b[u] = ldg_f8(base + (size_t)i2 * C); c[u] = ldg_f8(base + (size_t)i3 * C);
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (row[u] >= rows) continue;
            F8 o;
#pragma unroll
            for (int k = 0; k < 8; ++k) o.v[k] = interp3(a[u].v[k], b[u].v[k], c[u].v[k], w1[u], w2[u], w3[u]);
            stcs_f8(out + row[u] * C + l * 8, o);
        }
    }
}

__global__ void __launch_bounds__(kThreads) three_interpolate_kernel(long total, int m, int n, int c, const float *__restrict__ points,
                                                                     const int *__restrict__ idx, const float *__restrict__ weight,
                                                                     float *__restrict__ out) {
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        long row = e / c;
        int l = (int)(e - row * c);
        long bi = row / n;
        const int *ip = idx + row * 3;
        const float *wp = weight + row * 3;
        const float *base = points + bi * m * c + l;
        out[e] = interp3(__ldg(base + (long)__ldg(ip) * c), __ldg(base + (long)__ldg(ip + 1) * c), __ldg(base + (long)__ldg(ip + 2) * c),
                         __ldg(wp), __ldg(wp + 1), __ldg(wp + 2));
    }
}

// grad_points[i_t,l] += grad_out[j,l]*w_t   (tf_interpolate.cpp:131-153)
__global__ void __launch_bounds__(kThreads) three_interpolate_grad_kernel(long total, int m, int n, int c, const float *__restrict__ grad_out,
                                                                          const int *__restrict__ idx, const float *__restrict__ weight,
                                                                          float *__restrict__ grad_points) {
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        long row = e / c;
        int l = (int)(e - row * c);
        long bi = row / n;
        const int *ip = idx + row * 3;
        const float *wp = weight + row * 3;
        float g = __ldg(grad_out + e);
        float *base = grad_points + bi * m * c + l;
        atomicAdd(base + (long)__ldg(ip) * c, __fmul_rn(g, __ldg(wp)));
        atomicAdd(base + (long)__ldg(ip + 1) * c, __fmul_rn(g, __ldg(wp + 1)));
        atomicAdd(base + (long)__ldg(ip + 2) * c, __fmul_rn(g, __ldg(wp + 2)));
    }
}

// NnDistanceGrad, one direction (tf_nndistance_g.cu:132-151): g=2*grad*(p-q); p += g, q -= g.
__global__ void __launch_bounds__(kThreads) nn_distance_grad_kernel(long total, int n, int m, const float *__restrict__ xyz1,
                                                                    const float *__restrict__ xyz2, const float *__restrict__ grad_dist1,
                                                                    const int *__restrict__ idx1, float *__restrict__ grad_xyz1,
                                                                    float *__restrict__ grad_xyz2) {
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        long bi = e / n;
        int j2 = __ldg(idx1 + e);
        const float *p = xyz1 + e * 3;
        const float *q = xyz2 + (bi * m + j2) * 3;
        float g = __fmul_rn(__ldg(grad_dist1 + e), 2.0f);
#pragma unroll
        for (int l = 0; l < 3; ++l) {
            float v = __fmul_rn(g, __fsub_rn(__ldg(p + l), __ldg(q + l)));
            atomicAdd(grad_xyz1 + e * 3 + l, v);
            atomicAdd(grad_xyz2 + (bi * m + j2) * 3 + l, -v);
        }
    }
}

// ---- deterministic scatter-add (the reference's backward ops add with atomics in whatever order the hardware picks:
// tf_sampling_g.cu:183-192, tf_grouping_g.cu:66-83, and its CPU ThreeInterpolateGrad sums in index order).  Float addition does
// not commute in rounding, integer addition does: every contribution is scaled by a power of two chosen from the largest
// |contribution| and the number of entries per cloud, rounded to a 64-bit integer and added with integer atomics; the sum is
// converted back at the end.  Bit-reproducible from run to run and order-independent; the quantisation step is 2^-(61 - log2(E))
// of the largest contribution, far below fp32 round-off.  One generic kernel set serves gather / group / three_interpolate:
//   dst[b, idx[b, e], :] += weight[b, e] * src[b, e / rep, :]      e = 0 .. E-1 per cloud (rep = 3 for three_interpolate)
__global__ void __launch_bounds__(kThreads) det_absmax_kernel(long total, int c, int rep, const float *__restrict__ src,
                                                              const float *__restrict__ weight, unsigned *__restrict__ amax_bits) {
    float mx = 0.f;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const long ent = e / c;  // global entry index b*E + e
        const int l = (int)(e - ent * c);
        const float v = __ldg(src + (ent / rep) * c + l) * (weight ? __ldg(weight + ent) : 1.f);
        mx = fmaxf(mx, fabsf(v));
    }
    mx = __uint_as_float(__reduce_max_sync(GSPN_FULL_MASK, __float_as_uint(mx)));  // non-negative floats order as unsigned ints
    if ((threadIdx.x & 31) == 0 && mx > 0.f) atomicMax(amax_bits, __float_as_uint(mx));
}
__device__ __forceinline__ float det_scale(unsigned amax_bits, long entries_per_cloud) {
    // 2^k with  amax * entries * 2^k < 2^62 ; amax == 0 (all-zero gradient) -> any scale works
    const float amax = __uint_as_float(amax_bits);
    int ea = 0;
    if (amax > 0.f && amax < 3e38f) frexpf(amax, &ea);  // amax < 2^ea
    int ee = 0;
    while ((1L << ee) < entries_per_cloud) ++ee;
    int k = 61 - ea - ee;
    k = k > 126 ? 126 : (k < -126 ? -126 : k);
    return ldexpf(1.0f, k);
}
__global__ void __launch_bounds__(kThreads) det_scatter_kernel(long total, int n_dst, long E, int c, int rep, const float *__restrict__ src,
                                                               const int *__restrict__ idx, const float *__restrict__ weight,
                                                               const unsigned *__restrict__ amax_bits, long long *__restrict__ acc) {
    const float scale = det_scale(*amax_bits, E);
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const long ent = e / c;
        const int l = (int)(e - ent * c);
        const long bi = ent / E;
        const float v = __fmul_rn(__ldg(src + (ent / rep) * c + l), weight ? __ldg(weight + ent) : 1.f);  // the contribution, rounded as the op rounds it
        const long long q = __float2ll_rn(v * scale);  // exact scaling by a power of two, then ONE rounding to the integer grid
        atomicAdd(reinterpret_cast<unsigned long long *>(acc) + (bi * n_dst + __ldg(idx + ent)) * c + l, (unsigned long long)q);
    }
}
__global__ void __launch_bounds__(kThreads) det_finish_kernel(long total, long E, const unsigned *__restrict__ amax_bits,
                                                              const long long *__restrict__ acc, float *__restrict__ dst) {
    const float inv = 1.0f / det_scale(*amax_bits, E);
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x)
        dst[e] = (float)((double)acc[e] * (double)inv);
}

// box_shrink (models/model_rpointnet.py:529-551): tighten every box to the points it contains.  One warp per box; the
// reference's "large number" trick is kept verbatim (outside points are shifted by -/+ gamma = 1e4 before the max / min over ALL
// points), so an empty box yields max - min < 0 and is zeroed exactly as there.
__global__ void __launch_bounds__(kThreads) box_shrink_kernel(int nbox_per_cloud, int n, long nbox, const float *__restrict__ box,
                                                              const float *__restrict__ pc, float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long w = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
    if (w >= nbox) return;
    const float *bx = box + w * 6;
    const float *p = pc + (w / nbox_per_cloud) * (long)n * 3;
    float lo[3], hi[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float half = __fdiv_rn(__ldg(bx + 3 + a), 2.0f);  // box[...,3:]/2
        lo[a] = __fsub_rn(__ldg(bx + a), half);
        hi[a] = __fadd_rn(__ldg(bx + a), half);
    }
    const float gamma = 1e4f;
    float mx[3] = {-3.402823466e38f, -3.402823466e38f, -3.402823466e38f}, mn[3] = {3.402823466e38f, 3.402823466e38f, 3.402823466e38f};
    for (int k = lane; k < n; k += 32) {
        const float x = __ldg(p + 3 * k), y = __ldg(p + 3 * k + 1), z = __ldg(p + 3 * k + 2);
        const bool in = x >= lo[0] && x <= hi[0] && y >= lo[1] && y <= hi[1] && z >= lo[2] && z <= hi[2];
        const float shift = in ? 0.0f : gamma;  // gamma * (1 - mask)
        mx[0] = fmaxf(mx[0], __fsub_rn(x, shift)); mx[1] = fmaxf(mx[1], __fsub_rn(y, shift)); mx[2] = fmaxf(mx[2], __fsub_rn(z, shift));
        mn[0] = fminf(mn[0], __fadd_rn(x, shift)); mn[1] = fminf(mn[1], __fadd_rn(y, shift)); mn[2] = fminf(mn[2], __fadd_rn(z, shift));
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
        for (int o = 16; o > 0; o >>= 1) {
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(GSPN_FULL_MASK, mx[a], o));
            mn[a] = fminf(mn[a], __shfl_xor_sync(GSPN_FULL_MASK, mn[a], o));
        }
    if (lane == 0) {
        bool keep = true;
        float c[3], e[3];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float ext = __fsub_rn(mx[a], mn[a]);
            keep = keep && ext > 0.0f;                                 // tf.greater(box_max - box_min, 0), all three axes
            c[a] = __fdiv_rn(__fadd_rn(mx[a], mn[a]), 2.0f);           // (box_max + box_min) / 2
            e[a] = __fadd_rn(ext, 1e-3f);                              // box_max - box_min + 1e-3
        }
        const float k = keep ? 1.0f : 0.0f;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            out[w * 6 + a] = __fmul_rn(c[a], k);
            out[w * 6 + 3 + a] = __fmul_rn(e[a], k);
        }
    }
}

static inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace gspn

using namespace gspn;

extern "C" int gspn_gather_point(int b, int n, int m, int c, const float *inp, const int *idx, float *out, gspn_stream_t stream) {
    GSPN_REQUIRE(b >= 0 && n > 0 && m >= 0 && c > 0);  // tf_sampling.cpp:131-137
    if (b == 0 || m == 0) return GSPN_OK;
    GSPN_REQUIRE_PTR(inp); GSPN_REQUIRE_PTR(idx); GSPN_REQUIRE_PTR(out);
    if (c % 4 == 0 && aligned16(inp) && aligned16(out)) {
        long total = (long)b * m * (c / 4);
        gather_rows_kernel<float4><<<blocks_for(total), kThreads, 0, as_stream(stream)>>>(total, n, m, c / 4, (const float4 *)inp, idx, (float4 *)out);
    } else {
        long total = (long)b * m * c;
        gather_rows_kernel<float><<<blocks_for(total), kThreads, 0, as_stream(stream)>>>(total, n, m, c, inp, idx, out);
    }
    return check_launch();
}

extern "C" int gspn_gather_point_grad(int b, int n, int m, int c, const float *out_g, const int *idx, float *inp_g, gspn_stream_t stream) {
    GSPN_REQUIRE(b >= 0 && n > 0 && m >= 0 && c > 0);
    if (b == 0) return GSPN_OK;
    GSPN_REQUIRE_PTR(inp_g);
    GSPN_CUDA_OK(cudaMemsetAsync(inp_g, 0, sizeof(float) * (size_t)b * n * c, as_stream(stream)));  // tf_sampling.cpp:174
    if (m == 0) return GSPN_OK;
    GSPN_REQUIRE_PTR(out_g); GSPN_REQUIRE_PTR(idx);
    long total = (long)b * m * c;
    scatter_rows_kernel<<<blocks_for(total), kThreads, 0, as_stream(stream)>>>(total, n, m, c, out_g, idx, inp_g);
    return check_launch();
}

extern "C" int gspn_group_point(int b, int n, int c, int m, int nsample, const float *points, const int *idx, float *out, gspn_stream_t stream) {
    GSPN_REQUIRE(b >= 0 && n > 0 && c > 0 && m >= 0 && nsample > 0);  // tf_grouping.cpp:179-187
    if (b == 0 || m == 0) return GSPN_OK;
    GSPN_REQUIRE_PTR(points); GSPN_REQUIRE_PTR(idx); GSPN_REQUIRE_PTR(out);
    // a (b,m,nsample) index tensor is a (b, m*nsample) gather
    long mk = (long)m * nsample;
    GSPN_REQUIRE(mk < (1L << 31));
    if (c % 4 == 0 && aligned16(points) && aligned16(out)) {
        long total = (long)b * mk * (c / 4);
        gather_rows_kernel<float4><<<blocks_for(total), kThreads, 0, as_stream(stream)>>>(total, n, (int)mk, c / 4, (const float4 *)points, idx, (float4 *)out);
    } else {
        long total = (long)b * mk * c;
        gather_rows_kernel<float><<<blocks_for(total), kThreads, 0, as_stream(stream)>>>(total, n, (int)mk, c, points, idx, out);
    }
    return check_launch();
}

extern "C" int gspn_group_point_grad(int b, int n, int c, int m, int nsample, const float *grad_out, const int *idx, float *grad_points,
                                     gspn_stream_t stream) {
    GSPN_REQUIRE(b >= 0 && n > 0 && c > 0 && m >= 0 && nsample > 0);
    if (b == 0) return GSPN_OK;
    GSPN_REQUIRE_PTR(grad_points);
    GSPN_CUDA_OK(cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)b * n * c, as_stream(stream)));  // tf_grouping.cpp:234
    if (m == 0) return GSPN_OK;
    GSPN_REQUIRE_PTR(grad_out); GSPN_REQUIRE_PTR(idx);
    long mk = (long)m * nsample;
    GSPN_REQUIRE(mk < (1L << 31));
    if (c % 4 == 0 && aligned16(grad_out) && aligned16(grad_points)) {
        long total = (long)b * mk * (c / 4);
        scatter_rows_v4_kernel<<<blocks_for(total), kThreads, 0, as_stream(stream)>>>(total, n, (int)mk, c / 4, (const float4 *)grad_out, idx, (float4 *)grad_points);
    } else {
        long total = (long)b * mk * c;
        scatter_rows_kernel<<<blocks_for(total), kThreads, 0, as_stream(stream)>>>(total, n, (int)mk, c, grad_out, idx, grad_points);
    }
    return check_launch();
}

extern "C" int gspn_three_interpolate(int b, int m, int c, int n, const float *points, const int *idx, const float *weight, float *out,
                                      gspn_stream_t stream) {
    GSPN_REQUIRE(b >= 0 && m > 0 && c > 0 && n >= 0);  // tf_interpolate.cpp:197-206
    if (b == 0 || n == 0) return GSPN_OK;
    GSPN_REQUIRE_PTR(points); GSPN_REQUIRE_PTR(idx); GSPN_REQUIRE_PTR(weight); GSPN_REQUIRE_PTR(out);
    const long rows = (long)b * n;
    const bool a32 = ((reinterpret_cast<uintptr_t>(points) | reinterpret_cast<uintptr_t>(out)) & 31u) == 0;
    if (a32 && rows < (1L << 31) && (c == 64 || c == 128 || c == 256)) {
        // one group of c/8 lanes per row, two rows in flight per group; grid sized so that every SM holds 8 CTAs of work at most
        const int rpw = 2 * (256 / c);  // rows per warp per iteration
        long blk = ceil_div_l(ceil_div_l(rows, rpw) * 32, kThreads);
        const long cap = 148L * 8 * 4;
        if (blk > cap) blk = cap;
        const FastDiv dn((uint32_t)n);
        cudaStream_t st = as_stream(stream);
        if (c == 64) three_interpolate_rows_kernel<8><<<(unsigned)blk, kThreads, 0, st>>>(rows, m, n, points, idx, weight, out, dn);
        else if (c == 128) three_interpolate_rows_kernel<16><<<(unsigned)blk, kThreads, 0, st>>>(rows, m, n, points, idx, weight, out, dn);
        else three_interpolate_rows_kernel<32><<<(unsigned)blk, kThreads, 0, st>>>(rows, m, n, points, idx, weight, out, dn);
    } else if (c % 4 == 0 && aligned16(points) && aligned16(out)) {
        long total = (long)b * n * (c / 4);
        three_interpolate_v4_kernel<<<blocks_for(total), kThreads, 0, as_stream(stream)>>>(total, m, n, c / 4, (const float4 *)points, idx, weight, (float4 *)out);
    } else {
        long total = (long)b * n * c;
        three_interpolate_kernel<<<blocks_for(total), kThreads, 0, as_stream(stream)>>>(total, m, n, c, points, idx, weight, out);
    }
    return check_launch();
}

extern "C" int gspn_three_interpolate_grad(int b, int n, int c, int m, const float *grad_out, const int *idx, const float *weight,
                                           float *grad_points, gspn_stream_t stream) {
    GSPN_REQUIRE(b >= 0 && m > 0 && c > 0 && n >= 0);
    if (b == 0) return GSPN_OK;
    GSPN_REQUIRE_PTR(grad_points);
    GSPN_CUDA_OK(cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)b * m * c, as_stream(stream)));  // tf_interpolate.cpp:258
    if (n == 0) return GSPN_OK;
    GSPN_REQUIRE_PTR(grad_out); GSPN_REQUIRE_PTR(idx); GSPN_REQUIRE_PTR(weight);
    long total = (long)b * n * c;
    three_interpolate_grad_kernel<<<blocks_for(total), kThreads, 0, as_stream(stream)>>>(total, m, n, c, grad_out, idx, weight, grad_points);
    return check_launch();
}

extern "C" int gspn_nn_distance_grad(int b, int n, int m, const float *xyz1, const float *xyz2, const float *grad_dist1, const int *idx1,
                                     const float *grad_dist2, const int *idx2, float *grad_xyz1, float *grad_xyz2, gspn_stream_t stream) {
    GSPN_REQUIRE(b >= 0 && n > 0 && m > 0);
    if (b == 0) return GSPN_OK;
    GSPN_REQUIRE_PTR(xyz1); GSPN_REQUIRE_PTR(xyz2); GSPN_REQUIRE_PTR(grad_dist1); GSPN_REQUIRE_PTR(idx1);
    GSPN_REQUIRE_PTR(grad_dist2); GSPN_REQUIRE_PTR(idx2); GSPN_REQUIRE_PTR(grad_xyz1); GSPN_REQUIRE_PTR(grad_xyz2);
    cudaStream_t s = as_stream(stream);
    GSPN_CUDA_OK(cudaMemsetAsync(grad_xyz1, 0, sizeof(float) * (size_t)b * n * 3, s));  // tf_nndistance_g.cu:153-154
    GSPN_CUDA_OK(cudaMemsetAsync(grad_xyz2, 0, sizeof(float) * (size_t)b * m * 3, s));
    long t1 = (long)b * n, t2 = (long)b * m;
    nn_distance_grad_kernel<<<blocks_for(t1), kThreads, 0, s>>>(t1, n, m, xyz1, xyz2, grad_dist1, idx1, grad_xyz1, grad_xyz2);
    nn_distance_grad_kernel<<<blocks_for(t2), kThreads, 0, s>>>(t2, m, n, xyz2, xyz1, grad_dist2, idx2, grad_xyz2, grad_xyz1);
    return check_launch();
}

extern "C" int gspn_box_shrink(int b, int nbox, int n, const float *box, const float *pc, float *out, gspn_stream_t stream) {
    GSPN_REQUIRE(b >= 0 && nbox >= 0 && n > 0);
    if (b == 0 || nbox == 0) return GSPN_OK;
    GSPN_REQUIRE_PTR(box); GSPN_REQUIRE_PTR(pc); GSPN_REQUIRE_PTR(out);
    const long total = (long)b * nbox;
    box_shrink_kernel<<<(unsigned)ceil_div_l(total * 32, kThreads), kThreads, 0, as_stream(stream)>>>(nbox, n, total, box, pc, out);
    return check_launch();
}


// ---- deterministic forms of the backward ops (see det_scatter_kernel) ---------------------------------------------------------
extern "C" size_t gspn_scatter_det_workspace_bytes(int b, int n_dst, int c) {
    if (b <= 0 || n_dst <= 0 || c <= 0) return 0;
    return 256 + sizeof(long long) * (size_t)b * n_dst * c;
}

static int scatter_det(int b, int n_dst, long E, int c, int rep, const float *src, const int *idx, const float *weight, float *dst,
                       void *workspace, size_t workspace_bytes, gspn_stream_t stream) {
    GSPN_REQUIRE(b >= 0 && n_dst > 0 && c > 0 && E >= 0);
    if (b == 0) return GSPN_OK;
    GSPN_REQUIRE_PTR(dst);
    cudaStream_t st = as_stream(stream);
    const long ndst = (long)b * n_dst * c;
    if (E == 0) { GSPN_CUDA_OK(cudaMemsetAsync(dst, 0, sizeof(float) * ndst, st)); return GSPN_OK; }
    GSPN_REQUIRE_PTR(src); GSPN_REQUIRE_PTR(idx);
    if (workspace == nullptr || workspace_bytes < gspn_scatter_det_workspace_bytes(b, n_dst, c)) return GSPN_E_WORKSPACE;
    unsigned *amax = reinterpret_cast<unsigned *>(workspace);
    long long *acc = reinterpret_cast<long long *>(reinterpret_cast<unsigned char *>(workspace) + 256);
    GSPN_CUDA_OK(cudaMemsetAsync(workspace, 0, 256 + sizeof(long long) * (size_t)ndst, st));
    const long total = (long)b * E * c;
    det_absmax_kernel<<<blocks_for(total), kThreads, 0, st>>>(total, c, rep, src, weight, amax);
    det_scatter_kernel<<<blocks_for(total), kThreads, 0, st>>>(total, n_dst, E, c, rep, src, idx, weight, amax, acc);
    det_finish_kernel<<<blocks_for(ndst), kThreads, 0, st>>>(ndst, E, amax, acc, dst);
    return check_launch();
}

extern "C" int gspn_gather_point_grad_det(int b, int n, int m, int c, const float *out_g, const int *idx, float *inp_g, void *workspace,
                                          size_t workspace_bytes, gspn_stream_t stream) {
    return scatter_det(b, n, m, c, 1, out_g, idx, nullptr, inp_g, workspace, workspace_bytes, stream);
}
extern "C" int gspn_group_point_grad_det(int b, int n, int c, int m, int nsample, const float *grad_out, const int *idx, float *grad_points,
                                         void *workspace, size_t workspace_bytes, gspn_stream_t stream) {
    GSPN_REQUIRE(m >= 0 && nsample > 0);
    return scatter_det(b, n, (long)m * nsample, c, 1, grad_out, idx, nullptr, grad_points, workspace, workspace_bytes, stream);
}
extern "C" int gspn_three_interpolate_grad_det(int b, int n, int c, int m, const float *grad_out, const int *idx, const float *weight,
                                               float *grad_points, void *workspace, size_t workspace_bytes, gspn_stream_t stream) {
    GSPN_REQUIRE(n >= 0);
    if (b > 0 && n > 0) GSPN_REQUIRE_PTR(weight);
    return scatter_det(b, m, (long)n * 3, c, 3, grad_out, idx, weight, grad_points, workspace, workspace_bytes, stream);
}
