/*
 * gspn_b200.h -- C ABI of libgspn_b200.so: the PointNet++ set-abstraction /
 * feature-propagation hot path of ericyi/GSPN as hand-written sm_100a CUDA.
 *
 * This is the boundary a maintainer of the reference would bind instead of the
 * four TensorFlow op libraries (tf_sampling_so.so, tf_grouping_so.so,
 * tf_interpolate_so.so, tf_nndistance_so.so).  Every entry point cites the
 * reference interface it replaces (file:line relative to the reference root).
 *
 * Conventions:
 *   - plain pointers and sizes only; all tensor pointers are DEVICE pointers,
 *     C-contiguous, channel-last, float32 / int32 exactly like the reference
 *     tensors ((b,n,3), (b,n,c), (b,m,nsample), (b,m,nsample,c)).
 *   - the caller owns every buffer (outputs and workspace); the library never
 *     allocates, frees or keeps state between calls, and is re-entrant.
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*) and the
 *     call returns without synchronising; it is CUDA-graph capturable.
 *   - return value: GSPN_OK (0) or a negative GSPN_E_* code; nothing throws.
 *     The checks mirror the reference's OP_REQUIRES shape/attr checks.
 *   - backward entry points zero their own outputs on `stream` first, as the
 *     reference's Compute() does with cudaMemset.
 */
#ifndef GSPN_B200_H_
#define GSPN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *gspn_stream_t; /* cudaStream_t */

enum {
    GSPN_OK = 0,
    GSPN_E_BAD_SHAPE = -1,   /* OP_REQUIRES shape / attr check failed            */
    GSPN_E_NULL_PTR = -2,    /* a required pointer is NULL                       */
    GSPN_E_BAD_DTYPE = -3,   /* unknown GSPN_DT_* value                          */
    GSPN_E_WORKSPACE = -4,   /* workspace missing or smaller than *_workspace_bytes */
    GSPN_E_CUDA = -5,        /* cudaGetLastError() after the launch was not success */
    GSPN_E_UNSUPPORTED = -6  /* valid request outside what this build implements */
};

/* element / image types: F32 and BF16 rows or tile images; BF16X2 = the split tile image of the bf16x3 arithmetic (every block a
 * [hi | lo] pair with hi = bf16(x), lo = bf16(x - hi)); F16 = IEEE half (16-bit output copy only). */
enum { GSPN_DT_F32 = 0, GSPN_DT_BF16 = 1, GSPN_DT_BF16X2 = 2, GSPN_DT_F16 = 3 };
/* arithmetic of the tensor-core MLP chain.  BF16: one bf16 x bf16 product per term (unit round-off 2^-8; what north_star names).
 * BF16X3: split-bf16 -- operands carried as hi + lo bf16 pairs, a*b accumulated as hi*hi + lo*hi + hi*lo in the fp32 accumulator
 * (three tcgen05.mma per k-slice, error ~2^-16): the mode that keeps the MLP inside the reference's fp32 results to 1e-3. */
enum { GSPN_MLP_BF16 = 0, GSPN_MLP_BF16X3 = 1 };

/* Human-readable text for a GSPN_E_* code (static storage). */
const char *gspn_error_string(int code);
/* ABI version: major*1000+minor. */
int gspn_version(void);
/* cudaGetErrorString of the last GSPN_E_CUDA seen by the calling thread. */
const char *gspn_last_cuda_error(void);

/* Measurement aid (bench.py): launches `blocks` x 256 threads of `iters` x 16 independent FFMA each on `stream` and writes the
 * flop count of the launch to *flops_out (host); timing it gives the FP32 FMA rate the search kernels' "pair evaluations / s"
 * roofline is quoted against.  scratch: one device float. */
int gspn_fp32_peak_probe(int blocks, int iters, float *scratch, double *flops_out, gspn_stream_t stream);

/* ------------------------------------------------------------------ sampling
 * farthest_point_sample(npoint, inp)  tf_ops/sampling/tf_sampling.py:48-56
 *   FarthestPointSampleGpuOp::Compute  tf_ops/sampling/tf_sampling.cpp:95-122
 *   farthestpointsamplingLauncher(b,n,m,inp,temp,out)  tf_sampling_g.cu:203
 * inp (b,n,3) f32 -> out (b,m) i32, bit-identical to the reference kernel
 * (out[:,0]=0; ties -> lowest (k mod 512, k)).  The reference's temp (32,n) scratch is not needed.  Kernels by cloud size:
 *   up to 131072 points (gspn_fps_max_resident_points()): register-resident cluster kernels (csrc/fps.cu), no workspace; up to
 *   524288 a 16-CTA cluster keeps the distances in registers and streams coordinates from L2; above that the single-CTA fallback
 *   needs workspace.  Opt-in for 8193 .. 32768 points with workspace, two exact bucket-PRUNED kernels (the cloud sorted along a
 *   Hilbert curve into 1024 buckets of 32 points; a round only updates the buckets whose bounding box the new sample can reach --
 *   same indices, 80x fewer distance evaluations): gspn_fps_tune(1) = csrc/fps_bucket.cu, resident in shared memory + tensor memory
 *   + registers of ONE SM; gspn_fps_tune(2) = csrc/fps_pruned.cu, register-resident on the 8-CTA cluster.  Both measured slower
 *   per cloud than the full scan (the round's dependent chain, not the arithmetic, is what binds), so neither is the default.
 * Pass workspace of gspn_farthest_point_sample_workspace_bytes(b,n,m) bytes (0 when none is used). */
size_t gspn_farthest_point_sample_workspace_bytes(int b, int n, int m);
int gspn_fps_max_resident_points(void);
int gspn_farthest_point_sample(int b, int n, int m, const float *inp, int *out,
                               void *workspace, size_t workspace_bytes, gspn_stream_t stream);
/* Tuning door used by bench/tests: force (threads per CTA, points per thread,
 * CTAs per cluster); 0 = choose.  Same results for every legal choice. */
int gspn_farthest_point_sample_cfg(int b, int n, int m, const float *inp, int *out,
                                   int threads, int ppt, int cluster, gspn_stream_t stream);

/* Tuning doors (process-wide, not thread-safe).  gspn_fps_tune(mode): 0 = full-scan kernels for every size (default), 1 = the
 * single-CTA bucket kernel where it applies, 2 = the pruned cluster kernel where it applies.  gspn_fps_bucket_profile runs the bucket kernel and writes for cloud 0: prof3[0] = SM cycles of the round loop,
 * prof3[1] = rounds, prof3[2] = bucket updates, prof3[3..7] = warp 0's cycles in: box tests, bucket updates, warp argmax, barrier
 * wait, table reduce, prof3[8] = bucket updates that needed the full argmax (prof3: 12 x int64, zeroed by the caller). */
void gspn_fps_tune(int mode);
/* (threads per CTA, points per thread, CTAs per cluster) of the register-resident kernel for clouds above 16384 points; 0,0,0 = the
 * built-in table.  Same results for every legal choice. */
void gspn_fps_tune_mapping(int threads, int ppt, int cluster);
/* clouds per CTA of the 8-CTA cluster kernel for 16385 .. 32768 points: 2 = every 256-thread CTA is two independent 128-thread
 * halves, one cloud each (same results and latency; a batch's FPS state fills half as many SMs instead of a third of twice as many). */
void gspn_fps_tune_pack(int clouds_per_cta);
int gspn_fps_bucket_profile(int b, int n, int m, const float *inp, int *out, void *workspace, size_t workspace_bytes,
                            long long *prof3, gspn_stream_t stream);

/* Runs the pruned cluster kernel (curve sort + rounds) and writes for cloud 0: prof5[0] = thread 0's cycles in box test + bucket
 * updates + warp candidate, prof5[2] = candidate exchange, prof5[3] = table reduce, prof5[4] = bucket updates over all warps
 * (prof5: 8 x int64, zeroed by the caller). */
int gspn_fps_pruned_profile(int b, int n, int m, const float *inp, int *out, void *workspace, size_t workspace_bytes,
                            long long *prof5, gspn_stream_t stream);

/* Tuning door: per-phase SM-cycle counts of thread 0 of cloud 0, summed over the m-1 rounds:
 * prof4[0..3] = distance update + argmax, warp reduce, candidate exchange, table reduce. */
int gspn_fps_profile(int b, int n, int m, const float *inp, int *out, int threads, int ppt, int cluster,
                     long long *prof4, gspn_stream_t stream);

/* gather_point(inp, idx)  tf_sampling.py:29-37; gatherpointLauncher tf_sampling_g.cu:206
 * inp (b,n,c) , idx (b,m) -> out (b,m,c).  The reference is hard-wired to c=3;
 * c is explicit here because the model also gathers colour (model_rpointnet.py:151). */
int gspn_gather_point(int b, int n, int m, int c, const float *inp, const int *idx, float *out, gspn_stream_t stream);
/* GatherPointGrad  tf_sampling.py:43-47; scatteraddpointLauncher tf_sampling_g.cu:209
 * out_g (b,m,c), idx (b,m) -> inp_g (b,n,c) (zeroed here, then scatter-added). */
int gspn_gather_point_grad(int b, int n, int m, int c, const float *out_g, const int *idx, float *inp_g, gspn_stream_t stream);

/* ------------------------------------------------------------------ grouping
 * query_ball_point(radius, nsample, xyz1, xyz2)  tf_ops/grouping/tf_grouping.py:8-20
 *   queryBallPointLauncher(b,n,m,radius,nsample,xyz1,xyz2,idx,pts_cnt)  tf_grouping_g.cu:186
 * xyz1 (b,n,3) dataset, xyz2 (b,m,3) queries -> idx (b,m,nsample) i32, pts_cnt (b,m) i32.
 * First nsample hits in index order, first hit back-fills the row; a row with
 * no hit (unwritten by the reference) is zero.
 * workspace (optional): gspn_grid_workspace_bytes(b, n) bytes.  With it, clouds of >= 4096 points are searched
 * through a per-cloud uniform grid instead of the O(n*m) scan -- identical results. */
size_t gspn_grid_workspace_bytes(int b, int npoints_scanned);
/* Workspace for the point-query searches (gspn_three_nn, gspn_nn_distance, gspn_nearest_point) that additionally lets them
 * bucket a large QUERY set (>= 65536 points per cloud) and visit the queries in cell order (neighbouring threads then walk the same cells);
 * results are identical, only the visiting order changes.  Passing just gspn_grid_workspace_bytes(b, scanned) is still valid. */
size_t gspn_grid_query_workspace_bytes(int b, int n_queries, int npoints_scanned);
int gspn_query_ball_point(int b, int n, int m, float radius, int nsample, const float *xyz1, const float *xyz2,
                          int *idx, int *pts_cnt, void *workspace, size_t workspace_bytes, gspn_stream_t stream);
/* Several nested balls around the same queries in ONE ordered scan (multi_encoding_net, models/model_rpointnet.py:49-61: three
 * query_ball_point calls with radii 0.5 / 1.0 / 1.5 around the same seeds).  radii / nsamples / idx / pts_cnt are HOST arrays of
 * nrad (<= 4) entries; idx[r] (b,m,nsamples[r]) and pts_cnt[r] (b,m) are device buffers.  Bit-identical to nrad calls of
 * gspn_query_ball_point. */
int gspn_query_ball_point_multi(int b, int n, int m, int nrad, const float *radii, const int *nsamples, const float *xyz1,
                                const float *xyz2, int *const *idx, int *const *pts_cnt, gspn_stream_t stream);
/* Tuning door (process-wide): queries per warp (1, 2, 4) / per CTA of the ordered-scan ball query; 0 = choose. */
void gspn_ballquery_tune(int queries_per_warp, int queries_per_cta);
/* group_point(points, idx)  tf_grouping.py:54-62; groupPointLauncher tf_grouping_g.cu:194
 * points (b,n,c), idx (b,m,nsample) -> out (b,m,nsample,c). */
int gspn_group_point(int b, int n, int c, int m, int nsample, const float *points, const int *idx, float *out, gspn_stream_t stream);
/* GroupPointGrad  tf_grouping.py:63-67; groupPointGradLauncher tf_grouping_g.cu:198
 * grad_out (b,m,nsample,c), idx -> grad_points (b,n,c) (zeroed here). */
int gspn_group_point_grad(int b, int n, int c, int m, int nsample, const float *grad_out, const int *idx, float *grad_points, gspn_stream_t stream);

/* Fused ball query + group (sample_and_group, utils/pointnet_util.py:40-48, and
 * multi_encoding_net, models/model_rpointnet.py:53-61) in ONE kernel.
 * Besides idx / pts_cnt it writes the neighbourhood rows
 *     row(b,j,s) = [ points[b,idx,:c] | xyz[b,idx]-new_xyz[b,j]-shift[b,j] | 0.. ]   (features first)
 * `points` may be NULL (c=0).  shift (b,m,3) may be NULL.  grouped_dtype selects
 *   GSPN_DT_F32 : plain row-major (b*m*nsample, ld) float rows, ld >= c+3;
 *   GSPN_DT_BF16: the tensor-core-ready tile image consumed by gspn_mlp_chain
 *                 (128-row x 64-col bf16 blocks, 128B-swizzled, ld = 64*ceil((c+3)/64));
 *   GSPN_DT_BF16X2: the same with every block stored as a [hi | lo] pair (GSPN_MLP_BF16X3 arithmetic).
 * points_dtype is the dtype of `points` (f32 as in the reference, or bf16 as
 * produced by gspn_mlp_chain). */
size_t gspn_grouped_bytes(long rows, int c_plus_xyz, int grouped_dtype);
int gspn_ballquery_group(int b, int n, int m, int c, float radius, int nsample,
                         const float *xyz, const float *new_xyz, const float *shift,
                         const void *points, int points_dtype,
                         int *idx, int *pts_cnt, void *grouped, int grouped_dtype, int ld,
                         void *workspace, size_t workspace_bytes, gspn_stream_t stream);

/* ------------------------------------------------------------- interpolation
 * three_nn(xyz1, xyz2)  tf_ops/3d_interpolation/tf_interpolate.py:8-17; threenn_cpu tf_interpolate.cpp:60
 * xyz1 (b,n,3) unknown, xyz2 (b,m,3) known -> dist (b,n,3) squared f32 ascending, idx (b,n,3) i32.
 * If weight != NULL also writes the inverse-distance weights of
 * pointnet_fp_module (utils/pointnet_util.py:157-160) so no elementwise pass is needed.
 * workspace (optional): gspn_grid_workspace_bytes(b, m) bytes; with it and m >= 1024 the known points are
 * bucketed in a uniform grid and each query visits a growing block of cells -- identical results. */
int gspn_three_nn(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx, float *weight,
                  void *workspace, size_t workspace_bytes, gspn_stream_t stream);
/* three_interpolate(points, idx, weight)  tf_interpolate.py:19-28; threeinterpolate_cpu tf_interpolate.cpp:107
 * points (b,m,c), idx (b,n,3), weight (b,n,3) -> out (b,n,c); (p1*w1+p2*w2)+p3*w3 without FMA. */
int gspn_three_interpolate(int b, int m, int c, int n, const float *points, const int *idx, const float *weight, float *out, gspn_stream_t stream);
/* ThreeInterpolateGrad  tf_interpolate.py:29-34; threeinterpolate_grad_cpu tf_interpolate.cpp:131
 * grad_out (b,n,c) -> grad_points (b,m,c) (zeroed here). */
int gspn_three_interpolate_grad(int b, int n, int c, int m, const float *grad_out, const int *idx, const float *weight, float *grad_points, gspn_stream_t stream);

/* Deterministic forms of the scatter-add backward ops (GatherPointGrad, GroupPointGrad, ThreeInterpolateGrad): the reference's GPU
 * kernels add with float atomics in hardware order (tf_sampling_g.cu:183-192, tf_grouping_g.cu:66-83), so its gradients differ from
 * run to run in the last bits.  Here every contribution is scaled by a power of two, rounded to a 64-bit integer and added with integer
 * atomics (integer addition is associative): bit-reproducible, order-independent, quantisation far below fp32 round-off.
 * workspace: gspn_scatter_det_workspace_bytes(b, rows of the gradient being written, c). */
size_t gspn_scatter_det_workspace_bytes(int b, int n_dst, int c);
int gspn_gather_point_grad_det(int b, int n, int m, int c, const float *out_g, const int *idx, float *inp_g, void *workspace,
                               size_t workspace_bytes, gspn_stream_t stream);
int gspn_group_point_grad_det(int b, int n, int c, int m, int nsample, const float *grad_out, const int *idx, float *grad_points,
                              void *workspace, size_t workspace_bytes, gspn_stream_t stream);
int gspn_three_interpolate_grad_det(int b, int n, int c, int m, const float *grad_out, const int *idx, const float *weight,
                                    float *grad_points, void *workspace, size_t workspace_bytes, gspn_stream_t stream);

/* --------------------------------------------------------------- nn_distance
 * nn_distance(xyz1, xyz2)  tf_ops/nn_distance/tf_nndistance.py:14-24
 *   nnsearch tf_nndistance.cpp:21-43 (CPU op) / NmDistanceKernelLauncher tf_nndistance_g.cu:128 (GPU op)
 * xyz1 (b,n,3), xyz2 (b,m,3) -> dist1 (b,n), idx1 (b,n), dist2 (b,m), idx2 (b,m).
 * rounding: 0 = as the CPU op rounds (mul,mul,add,mul,add), 1 = as the compiled
 * GPU kernel rounds (mul,fma,fma).  Ties -> lowest index in both.
 * workspace (optional): gspn_grid_workspace_bytes(b, max(n,m)); with it and n,m >= 2048 both directions search a
 * uniform grid over the scanned set instead of the O(n*m) scan -- identical results. */
int gspn_nn_distance(int b, int n, int m, const float *xyz1, const float *xyz2,
                     float *dist1, int *idx1, float *dist2, int *idx2, int rounding,
                     void *workspace, size_t workspace_bytes, gspn_stream_t stream);
/* NnDistanceGrad  tf_nndistance.py:31-37; tf_nndistance.cpp:126-163 / tf_nndistance_g.cu:152 */
int gspn_nn_distance_grad(int b, int n, int m, const float *xyz1, const float *xyz2,
                          const float *grad_dist1, const int *idx1, const float *grad_dist2, const int *idx2,
                          float *grad_xyz1, float *grad_xyz2, gspn_stream_t stream);

/* ---------------------------------------------------------------- shared MLP
 * The per-point MLP of pointnet_sa_module / pointnet_fp_module
 * (utils/pointnet_util.py:109-113,124 and :167-172): layers of
 * tf_util.conv2d 1x1 + bias + batch_norm(inference) + ReLU (utils/tf_util.py:170-184,530-534),
 * optionally followed by tf.reduce_max over groups of `pool` consecutive rows.
 *
 * fp32 reference-precision path (CUDA cores): one layer per call.
 *   x (rows,cin) f32, w (cin,cout) f32 (the [1,1,Cin,Cout] TF kernel), scale/shift (cout)
 *   y = act(x@w * scale + shift)  [bias and BN folded by the caller: scale=g/sqrt(v+eps),
 *   shift=(bias-mean)*scale+beta];  pool>1 (a divisor of 64): y (rows/pool, cout) = max over each group. */
int gspn_mlp_layer_f32(long rows, int cin, int cout, const float *x, int ldx, const float *w,
                       const float *scale, const float *shift, int relu, int pool, float *y, gspn_stream_t stream);

/* tf.reduce_max over groups of k consecutive rows: x (groups*k, c) -> y (groups, c)  (pointnet_util.py:124). */
int gspn_max_pool_rows(long groups, int k, int c, const float *x, float *y, gspn_stream_t stream);

/* Tensor-core path (tcgen05 / TMEM), whole chain in one kernel (csrc/mlp_tc.cu); `arith` = GSPN_MLP_BF16 or GSPN_MLP_BF16X3.
 *   a       : tile image from gspn_ballquery_group / gspn_fp_assemble (rows padded to 128); GSPN_DT_BF16 image for GSPN_MLP_BF16,
 *             GSPN_DT_BF16X2 image for GSPN_MLP_BF16X3
 *   nlayers : 1..4;  dims[0]=K0 (multiple of 64, as `ld` above), dims[l+1]=cout of layer l (multiples of 32, <= 512; hidden
 *             widths <= 256); k0_used = columns of the image that hold data (0: all K0) -- only those k-slices are multiplied
 *   wimg[l] : weight image from gspn_mlp_pack_weights (same arith); scale/shift as above (f32, cout_l)
 *   pool    : 1 (no pooling) or a multiple of 32 dividing rows (nsample)
 *   out_f32 (rows/pool, cout_last) and/or out_h (same shape, row-major, out_h_dtype = GSPN_DT_BF16 or GSPN_DT_F16 with
 *   saturation to the finite half range) may be NULL. */
size_t gspn_mlp_weight_image_bytes(int cin_padded, int cout, int arith);
int gspn_mlp_pack_weights(int cin, int cin_padded, int cout, const float *w_f32, const int *row_perm,
                          void *wimg, int arith, gspn_stream_t stream);
int gspn_mlp_chain(long rows, int nlayers, const int *dims, int k0_used, const void *a,
                   const void *const *wimg, const float *const *scale, const float *const *shift,
                   const int *relu, int pool, float *out_f32, void *out_h, int out_h_dtype, int arith, gspn_stream_t stream);

/* Same chain, but layer 0's operand rows [points[b,idx,:c] | xyz[b,idx]-new_xyz[b,j]-shift[b,j] | 0] (c+3 <= 8, dims[0]=64) are
 * gathered by the kernel's producer warps straight from the ball-query indices idx (b,m,nsample): the grouped tensor of
 * sample_and_group (utils/pointnet_util.py:40-48) is never written to HBM.  Same values as gspn_ballquery_group + gspn_mlp_chain. */
int gspn_mlp_chain_gather(int b, int n, int m, int nsample, int c, const float *xyz, const float *new_xyz, const float *shift_pred,
                          const float *points, const int *idx, int nlayers, const int *dims, const void *const *wimg,
                          const float *const *scale, const float *const *shift, const int *relu, int pool,
                          float *out_f32, void *out_h, int out_h_dtype, int arith, gspn_stream_t stream);

/* pointnet_fp_module's interpolate + concat + MLP (utils/pointnet_util.py:156-172) with the interpolated map never written:
 * three_interpolate is linear, so  concat(interp3(points2), points1) @ W0 = interp3(points2 @ W0[:c2]) + points1 @ W0[c2:].
 * The caller multiplies the m known points once (y2 = points2 @ W0[:c2], (b,m,n0) f32, e.g. a one-layer gspn_mlp_chain with scale 1 /
 * shift 0 / no ReLU); the kernel's producer warps gather the three rows of y2 per point, finish layer 0 on the CUDA cores
 * (act(scale[0] * (w1*y2[i1] + w2*y2[i2] + w3*y2[i3] + points1 @ w0b) + shift[0])) and feed layers 1.. to the tensor cores.
 *   idx / weight (b,n,3) from gspn_three_nn; points1 (b,n,c1) with c1 <= 4, or NULL; w0b = W0[c2:] (c1,n0) f32
 *   nlayers >= 2 counts layer 0; dims[0] ignored, dims[1] = n0 = 128, dims[l+1] = cout of layer l; wimg[0] unused. */
int gspn_mlp_chain_fp(int b, int n, int m, int c1, const float *y2, const int *idx, const float *weight, const float *points1,
                      const float *w0b, int nlayers, const int *dims, const void *const *wimg, const float *const *scale,
                      const float *const *shift, const int *relu, float *out_f32, void *out_h, int out_h_dtype, int arith,
                      gspn_stream_t stream);

/* ---- training form of the shared MLP (fp32): conv 1x1 + bias -> batch norm over the BATCH moments
 * (tf.contrib.layers.batch_norm, is_training=True, utils/tf_util.py:515-534) -> ReLU -> reduce_max, and its backward.
 * The GEMMs are gspn_mlp_layer_f32 (forward: scale=1, shift=bias, relu=0; dX: the same with W^T). */
/* sum[c] = sum_r z[r,c], sumsq[c] = sum_r z[r,c]^2 in double (zeroed here). */
int gspn_col_moments_f32(long rows, int c, const float *z, double *sum, double *sumsq, gspn_stream_t stream);
/* y = act((z-mean)*invstd*gamma+beta) */
int gspn_bn_act_f32(long rows, int c, const float *z, const float *mean, const float *invstd, const float *gamma, const float *beta,
                    int relu, float *y, gspn_stream_t stream);
/* out[g,c] = max_s y[g*k+s,c], argmax[g,c] = the winning s (first maximum). */
int gspn_maxpool_argmax_f32(long groups, int k, int c, const float *y, float *out, int *argmax, gspn_stream_t stream);
/* backward of pool(act(bn(z))): dy is (rows/pool, c) when pool>1 (argmax from the forward) else (rows, c).
 * Writes s1[c] = sum dy' (= dbeta), s2[c] = sum dy'*xhat (= dgamma) in double and dz (rows,c). dgamma/dbeta unused (NULL).
 * bn=0: layer without batch norm (pass mean=0, invstd=gamma=1, beta=0): dz = dy'. */
int gspn_bn_act_pool_bwd_f32(long rows, int c, int pool, int relu, int bn, const float *z, const float *dy, const int *argmax,
                             const float *mean, const float *invstd, const float *gamma, const float *beta,
                             double *s1, double *s2, float *dz, float *dgamma, float *dbeta, gspn_stream_t stream);
/* The same backward split in two, for batch norm over the WHOLE batch of a data-parallel job (utils/tf_util.py:530-534 normalises
 * over every row of the batch; sharded over ranks that needs the sums all-reduced): _sums writes this rank's s1 / s2, the caller
 * all-reduces them (and the row count), _apply computes dz with the global sums and total_rows. */
int gspn_bn_bwd_sums_f32(long rows, int c, int pool, int relu, const float *z, const float *dy, const int *argmax, const float *mean,
                         const float *invstd, const float *gamma, const float *beta, double *s1, double *s2, gspn_stream_t stream);
int gspn_bn_bwd_apply_f32(long rows, long total_rows, int c, int pool, int relu, const float *z, const float *dy, const int *argmax,
                          const float *mean, const float *invstd, const float *gamma, const float *beta, const double *s1,
                          const double *s2, float *dz, gspn_stream_t stream);
/* dW (cin,cout) = x^T dz, dbias (cout, may be NULL) = colsum(dz); both zeroed here; x has row stride ldx. */
int gspn_mlp_wgrad_f32(long rows, int cin, int cout, const float *x, int ldx, const float *dz, float *dW, float *dbias,
                       gspn_stream_t stream);
/* backward of the fused grouping w.r.t. the features: grad_points (b,n,c) (zeroed here) += grad_rows[(b,j,s), :c]
 * (rows of stride ld, features first) scattered through idx (b,m,nsample); GroupPointGrad tf_grouping_g.cu:66-83. */
int gspn_group_rows_grad(int b, int n, int c, int m, int nsample, int ld, const float *grad_rows, const int *idx,
                         float *grad_points, gspn_stream_t stream);

/* Tuning doors (benchmark A/B runs; process-wide, NOT thread-safe, nothing is read from the environment).
 * gspn_mlp_chain_set_profile: when prof (device, 16 x int64, zeroed by the caller) is non-NULL, later chain launches add CTA 0's
 * cycle counts: epilogue thread 0: [1] wait for the MMAs, [2] epilogue, [3] fences+hand-off, [4] number of steps;
 * MMA issuer: [5] issue, [6] drain until tcgen05.commit lands, [7] hand-off waits.  NULL switches it off.
 * gspn_mlp_chain_tune: occ_cap 1|2 = most chain CTAs per SM, bufs_cap 1|2 = most TMEM accumulator buffers,
 * tma_out 0 = row-per-lane 256-bit output stores instead of TMA tensor stores.  Defaults (2, 2, 1) are the measured best. */
/* The plan the launcher makes for a chain of this shape -- no device needed, nothing launched.  mode 0 = gspn_mlp_chain, 1 =
 * gspn_mlp_chain_gather, 2 = gspn_mlp_chain_fp (dims start at n0).  plan12 = {CTAs per SM, epilogue warps, threads per CTA, weight rows
 * per ring stage, accumulator columns per buffer, accumulator buffers, TMEM columns allocated, TMEM column of the activation operand,
 * operand ring stages, weight ring stages, dynamic shared memory bytes, passes of the last layer}. */
int gspn_mlp_chain_plan(int mode, long rows, int nlayers, const int *dims, int k0_used, int pool, int want_f32, int want_h, int arith,
                        int *plan12);
void gspn_mlp_chain_set_profile(long long *prof);
void gspn_mlp_chain_tune(int occ_cap, int bufs_cap, int tma_out);
/* gather warps of gspn_mlp_chain_fp: 8 (default; 448 threads, fastest alone) or 4 (320 threads: leaves registers for an FPS CTA on the SM) */
void gspn_mlp_chain_tune_fp(int gather_warps);
/* tile scheduling of the chain kernels: 0 (default) = a grid sized to the SMs walking the 128-row tiles with a static stride;
 * 1 = one CTA per tile in the grid, and a CTA that finishes a tile takes over a CTA the hardware has not launched yet
 * (clusterlaunchcontrol.try_cancel): whatever SMs are free share the tiles as they come.  Same results either way; measured a wash
 * in the pipelined step, so it is a door. */
void gspn_mlp_chain_tune_sched(int dynamic_tiles);

/* Feature-propagation front end (utils/pointnet_util.py:156-165) fused: three_interpolate of
 * points2 (b,m,c2) with idx/weight (b,n,3), concatenated with points1 (b,n,c1) (may be NULL, c1=0),
 * written straight into the tile image (ld = 64*ceil((c1+c2)/64); image_dtype GSPN_DT_BF16 or GSPN_DT_BF16X2) for gspn_mlp_chain.
 * c2 = 0 (points2, idx, weight NULL): plain rows points1 -> tile image. */
int gspn_fp_assemble(int b, int n, int m, int c1, int c2, const float *points1, const float *points2,
                     const int *idx, const float *weight, void *a_img, int ld, int image_dtype, gspn_stream_t stream);

/* ---- peer-memory all-reduce of small vectors over NVLink (csrc/p2p.cu): the whole-batch batch-norm statistics of the data-parallel
 * training form, 52 dependent collectives of a few hundred bytes per step.  Every rank owns a mailbox (created here -- the only entry
 * points of the library that allocate -- and mapped into the peers through its 64-byte CUDA IPC handle); one small kernel writes the
 * rank's vector into every peer's mailbox, publishes an epoch flag, waits for all contributions to its own mailbox and sums them in
 * rank order (bit-identical, deterministic sums on every rank).  No NCCL call, no host synchronisation, CUDA-graph capturable.
 *   mailboxes: HOST array of `world` device pointers, mailboxes[r] = rank r's mailbox as mapped in this process ([rank] = its own);
 *   data (device, `count` <= max_doubles doubles) is reduced in place; every rank must issue the same sequence of calls. */
size_t gspn_p2p_mailbox_bytes(int world, int max_doubles);
int gspn_p2p_mailbox_create(int world, int max_doubles, void **mailbox, unsigned char *ipc_handle64);
int gspn_p2p_mailbox_open(const unsigned char *ipc_handle64, void **peer_mailbox);
int gspn_p2p_mailbox_close(void *peer_mailbox);
int gspn_p2p_mailbox_destroy(void *mailbox);
int gspn_p2p_allreduce_f64(int rank, int world, int max_doubles, void *const *mailboxes, int count, double *data, gspn_stream_t stream);

/* ---- brute-force nearest-neighbour glue around the path ------------------------------
 * One-directional 1-NN: for every query (b,n,3) the nearest reference point (b,m,3): squared distance dist (b,n)
 * and index idx (b,n), lowest index on ties.  Replaces the model's dense (B,N,M) distance tensors + argmin:
 * nearest seed per point  tf.argmin(reduce_sum(square(pc - pc_seed), -1), 2)   models/model_rpointnet.py:1136,
 * nearest cropped ROI point in unmold_segmentation  :1032-1033, and test.py's sklearn ball-tree 1-NN
 * (test.py:165-166,184-185).  rounding as gspn_nn_distance (0: every product and sum rounded, the TF / CPU
 * arithmetic; 1: FMA chain).  workspace: gspn_grid_workspace_bytes(b, m) bytes or NULL (brute-force scan). */
int gspn_nearest_point(int b, int n, int m, const float *queries, const float *refs, float *dist, int *idx, int rounding,
                       void *workspace, size_t workspace_bytes, gspn_stream_t stream);
/* box_shrink (models/model_rpointnet.py:529-551): box (b,nbox,6) = (centre, extent), pc (b,n,3) -> out (b,nbox,6):
 * the tight box of the points inside each box (extent + 1e-3), all zeros for a box that holds no point or is
 * degenerate along an axis.  The reference's gamma = 1e4 shift of outside points is reproduced operation by operation. */
int gspn_box_shrink(int b, int nbox, int n, const float *box, const float *pc, float *out, gspn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GSPN_B200_H_ */
