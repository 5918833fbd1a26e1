// group_rows.cuh -- the row writer shared by the brute-force and the grid ball-query kernels: once a query's
// nsample indices are known, the owning warp writes the neighbourhood rows [features | xyz - centre | 0]
// either as fp32 rows or straight into the bf16 tile image (common.cuh).
#pragma once
#include "common.cuh"

namespace gspn {

struct GroupArgs {
    const float *shift;  // optional (b,m,3) extra shift (multi_encoding_net, model_rpointnet.py:56-57)
    const void *points;          // (b,n,c) f32 or bf16, may be null
    int c;
    int points_bf16;
    void *grouped;  // null -> indices only
    int grouped_bf16;  // 0: f32 rows, 1: bf16 tile image, 2: split (hi | lo) tile image (GSPN_DT_BF16X2)
    int ld;
};

// one element of a neighbourhood row: [ features(c) | (xyz - centre) - shift | 0 ]
struct RowSrc {
    const float *pts_f;
    const __nv_bfloat16 *pts_h;
    const float *xyz;  // cloud base
    int c;
    float qx, qy, qz, sx, sy, sz;
    bool has_shift;
    __device__ __forceinline__ float at(int ii, int col) const {
        if (col < c) {
            size_t o = (size_t)ii * c + col;
            return pts_h ? __bfloat162float(pts_h[o]) : __ldg(pts_f + o);
        }
        if (col < c + 3) {
            int a = col - c;
            float q = a == 0 ? qx : (a == 1 ? qy : qz);
            float v = __fsub_rn(__ldg(xyz + (size_t)ii * 3 + a), q);  // grouped_xyz -= new_xyz (pointnet_util.py:42)
            if (has_shift) v = __fsub_rn(v, a == 0 ? sx : (a == 1 ? sy : sz));  // -= shift_pred (model_rpointnet.py:56-57)
            return v;
        }
        return 0.f;
    }
};

// write one neighbourhood (nsample rows) of query (cloud,j); sidx = the row's indices in smem.  The (row, column) space is swept by
// `stride` threads of which this is number `lane`: one warp (lane, 32), or a whole CTA (threadIdx.x, blockDim.x).
__device__ __forceinline__ void write_group(const GroupArgs &g, int n, int m, int nsample, int cloud, int j, const int *sidx,
                                            const float *__restrict__ xyz, float qx, float qy, float qz, int lane, int stride = 32) {
    const long row0 = ((long)cloud * m + j) * nsample;
    RowSrc src;
    src.c = g.c;
    src.pts_f = g.points_bf16 ? nullptr : (const float *)g.points + (size_t)cloud * n * g.c;
    src.pts_h = g.points_bf16 ? (const __nv_bfloat16 *)g.points + (size_t)cloud * n * g.c : nullptr;
    src.xyz = xyz + (size_t)cloud * n * 3;
    src.qx = qx; src.qy = qy; src.qz = qz;
    src.has_shift = g.shift != nullptr;
    src.sx = src.sy = src.sz = 0.f;
    if (src.has_shift) {
        const float *sp = g.shift + ((size_t)cloud * m + j) * 3;
        src.sx = __ldg(sp); src.sy = __ldg(sp + 1); src.sz = __ldg(sp + 2);
    }
    if (!g.grouped_bf16) {
        // lanes sweep the (row, column) space of the block; columns are contiguous in memory
        float *out = (float *)g.grouped;
        const int ld = g.ld;
        for (int e = lane; e < nsample * ld; e += stride) {
            int s = e / ld, col = e - s * ld;
            out[(row0 + s) * ld + col] = src.at(sidx[s], col);
        }
        return;
    }
    // bf16 tile image: one 16-byte chunk (8 columns) per lane-step; the split image also gets the chunk of remainders
    unsigned char *img = (unsigned char *)g.grouped;
    const int split = g.grouped_bf16 == 2;
    const int chunks = g.ld >> 3;
    const bool vec_f = src.pts_f && (g.c % 4 == 0) && ((reinterpret_cast<uintptr_t>(src.pts_f) & 15u) == 0);
    const bool vec_h = src.pts_h && (g.c % 8 == 0) && ((reinterpret_cast<uintptr_t>(src.pts_h) & 15u) == 0);
    for (int e = lane; e < nsample * chunks; e += stride) {
        int s = e / chunks, ch = e - s * chunks;
        int ii = sidx[s];
        uint4 pk, pl = make_uint4(0, 0, 0, 0);
        if (ch * 8 + 8 <= g.c && vec_h) {
            pk = __ldg(reinterpret_cast<const uint4 *>(src.pts_h + (size_t)ii * g.c + ch * 8));  // bf16 features: remainders are zero
        } else {
            float v[8];
            if (ch * 8 + 8 <= g.c && vec_f) {
                const float4 *fp = reinterpret_cast<const float4 *>(src.pts_f + (size_t)ii * g.c + ch * 8);
                float4 a = __ldg(fp), b = __ldg(fp + 1);
                v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
            } else {
#pragma unroll
                for (int t = 0; t < 8; ++t) v[t] = src.at(ii, ch * 8 + t);
            }
            if (split) {
                split_bf16x2(v[0], v[1], pk.x, pl.x); split_bf16x2(v[2], v[3], pk.y, pl.y);
                split_bf16x2(v[4], v[5], pk.z, pl.z); split_bf16x2(v[6], v[7], pk.w, pl.w);
            } else {
                pk.x = pack_bf16x2(v[0], v[1]); pk.y = pack_bf16x2(v[2], v[3]); pk.z = pack_bf16x2(v[4], v[5]); pk.w = pack_bf16x2(v[6], v[7]);
            }
        }
        const size_t off = tile_chunk_offset(row0 + s, ch, g.ld, split);
        *reinterpret_cast<uint4 *>(img + off) = pk;
        if (split) *reinterpret_cast<uint4 *>(img + off + kTileBytes) = pl;
    }
}

}  // namespace gspn
