/*
 * gspn_oracle.c -- CPU restatement of the GSPN PointNet++ SA/FP op kernels.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (gspn_b200/) may
 * import, link or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs do, and only as the checker
 * or as the timed CPU baseline.
 *
 * Each function restates the arithmetic, visiting order and tie-breaking of one
 * reference kernel (file:line relative to /root/reference).  Rounding rules:
 *
 *  - ops whose reference implementation is a CUDA kernel (FPS, ball query,
 *    NmDistance GPU variant) use the contraction nvcc 12.9 -O2 emits for
 *    `(x2-x1)*(x2-x1)+(y2-y1)*(y2-y1)+(z2-z1)*(z2-z1)`, read from the PTX of the
 *    unmodified reference sources:  t=dy*dy; t=fma(dx,dx,t); d=fma(dz,dz,t).
 *  - ops whose reference implementation is a g++ -O2 CPU loop (three_nn,
 *    three_interpolate, nnsearch) use separately rounded float mul/add, left
 *    to right (x86-64 baseline has no FMA, so g++ cannot contract).
 *
 * Build with -ffp-contract=off so the compiler adds no contraction of its own;
 * fmaf() below is the only fused operation.  Pinning: see oracle/README.md --
 * the CPU-side functions are checked bit-for-bit against the reference's own
 * loops compiled into oracle/_ref/, the GPU-side ones against the reference's
 * own .cu kernels run on a B200 (tests/golden/).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#if defined(__GNUC__) && defined(__x86_64__) && !defined(__FMA__)
#error "build the oracle with -mfma (hardware fmaf) and -ffp-contract=off"
#endif

/* squared distance exactly as the compiled reference CUDA kernels round it
 * (tf_sampling_g.cu:142, tf_grouping_g.cu:27, tf_nndistance_g.cu:26). */
static inline float sqdist_gpu(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = ax - bx, dy = ay - by, dz = az - bz;
    float t = dy * dy;
    t = fmaf(dx, dx, t);
    return fmaf(dz, dz, t);
}

/* squared distance as g++ -O2 rounds it in the CPU ops
 * (tf_interpolate.cpp:73, tf_nndistance.cpp:30-33). */
static inline float sqdist_cpu(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = ax - bx, dy = ay - by, dz = az - bz;
    float xx = dx * dx, yy = dy * dy, zz = dz * dz;
    float s = xx + yy;
    return s + zz;
}

int gspn_oracle_has_fma(void) { return __builtin_cpu_supports("fma") ? 1 : 0; }

/* ---- farthestpointsamplingKernel, tf_sampling_g.cu:105-170 ----------------
 * 512 "threads"; thread t visits k=t,t+512,... ascending with a strict
 * d2>best (init best=-1,besti=0, :125-126,:146-149); then the 9-level tree of
 * :153-164 where the left operand survives ties.  temp[] starts at 1e38 (:118).
 * out[i*m+0]=0 (:114-116). */
void gspn_oracle_fps(int b, int n, int m, const float *xyz, int *out) {
    enum { BS = 512 };
    float *temp = (float *)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
    float dists[BS];
    int dists_i[BS];
    if (m <= 0) { free(temp); return; }
    for (int i = 0; i < b; ++i) {
        const float *p = xyz + (size_t)i * n * 3;
        int old = 0;
        out[(size_t)i * m] = old;
        for (int k = 0; k < n; ++k) temp[k] = 1e38f;
        for (int j = 1; j < m; ++j) {
            float x1 = p[old * 3 + 0], y1 = p[old * 3 + 1], z1 = p[old * 3 + 2];
            for (int t = 0; t < BS; ++t) { dists[t] = -1.0f; dists_i[t] = 0; }
            /* ascending k == every thread sees its own k's ascending */
            for (int k = 0; k < n; ++k) {
                int t = k & (BS - 1);
                float td = temp[k];
                float d = sqdist_gpu(p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2], x1, y1, z1);
                float d2 = d < td ? d : td; /* min(d,td) */
                if (d2 != td) temp[k] = d2;
                if (d2 > dists[t]) { dists[t] = d2; dists_i[t] = k; }
            }
            for (int u = 0; (1 << u) < BS; ++u) {
                for (int t = 0; t < (BS >> (u + 1)); ++t) {
                    int i1 = (t * 2) << u, i2 = (t * 2 + 1) << u;
                    if (dists[i1] < dists[i2]) { dists[i1] = dists[i2]; dists_i[i1] = dists_i[i2]; }
                }
            }
            old = dists_i[0];
            out[(size_t)i * m + j] = old;
        }
    }
    free(temp);
}

/* ---- gatherpointKernel, tf_sampling_g.cu:172-181 (3 channels, as the
 * reference; the product op also accepts c channels, oracle restates that
 * as the obvious row copy). */
void gspn_oracle_gather_point(int b, int n, int m, int c, const float *inp, const int *idx, float *out) {
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j) {
            int a = idx[(size_t)i * m + j];
            for (int l = 0; l < c; ++l)
                out[((size_t)i * m + j) * c + l] = inp[((size_t)i * n + a) * c + l];
        }
}

/* ---- scatteraddpointKernel, tf_sampling_g.cu:183-192 (serial order). */
void gspn_oracle_gather_point_grad(int b, int n, int m, int c, const float *out_g, const int *idx, float *inp_g) {
    memset(inp_g, 0, sizeof(float) * (size_t)b * n * c);
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j) {
            int a = idx[(size_t)i * m + j];
            for (int l = 0; l < c; ++l)
                inp_g[((size_t)i * n + a) * c + l] += out_g[((size_t)i * m + j) * c + l];
        }
}

/* ---- query_ball_point_gpu, tf_grouping_g.cu:6-39 --------------------------
 * ascending k, stop at nsample hits (:18-20), predicate
 * max(sqrtf(s),1e-20f) < radius (:27-28), first hit back-fills the row
 * (:29-32).  Rows with zero hits are left unwritten by the reference; both the
 * oracle and the product define them as zeros. */
void gspn_oracle_query_ball_point(int b, int n, int m, float radius, int nsample,
                                  const float *xyz1, const float *xyz2, int *idx, int *pts_cnt) {
    for (int i = 0; i < b; ++i) {
        const float *p = xyz1 + (size_t)i * n * 3;
        const float *q = xyz2 + (size_t)i * m * 3;
        int *id = idx + (size_t)i * m * nsample;
        for (int j = 0; j < m; ++j) {
            int cnt = 0;
            for (int l = 0; l < nsample; ++l) id[(size_t)j * nsample + l] = 0;
            for (int k = 0; k < n; ++k) {
                if (cnt == nsample) break;
                float s = sqdist_gpu(q[j * 3 + 0], q[j * 3 + 1], q[j * 3 + 2], p[k * 3 + 0], p[k * 3 + 1], p[k * 3 + 2]);
                float d = sqrtf(s);
                d = d > 1e-20f ? d : 1e-20f; /* max(sqrtf(s),1e-20f); NaN -> 1e-20f like CUDA fmaxf */
                if (d < radius) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; ++l) id[(size_t)j * nsample + l] = k;
                    id[(size_t)j * nsample + cnt] = k;
                    cnt += 1;
                }
            }
            pts_cnt[(size_t)i * m + j] = cnt;
        }
    }
}

/* ---- group_point_gpu, tf_grouping_g.cu:43-60. */
void gspn_oracle_group_point(int b, int n, int c, int m, int nsample, const float *points, const int *idx, float *out) {
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j)
            for (int k = 0; k < nsample; ++k) {
                int ii = idx[((size_t)i * m + j) * nsample + k];
                memcpy(out + (((size_t)i * m + j) * nsample + k) * c, points + ((size_t)i * n + ii) * c, sizeof(float) * c);
            }
}

/* ---- group_point_grad_gpu, tf_grouping_g.cu:66-83 (serial order). */
void gspn_oracle_group_point_grad(int b, int n, int c, int m, int nsample, const float *grad_out, const int *idx, float *grad_points) {
    memset(grad_points, 0, sizeof(float) * (size_t)b * n * c);
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < m; ++j)
            for (int k = 0; k < nsample; ++k) {
                int ii = idx[((size_t)i * m + j) * nsample + k];
                for (int l = 0; l < c; ++l)
                    grad_points[((size_t)i * n + ii) * c + l] += grad_out[(((size_t)i * m + j) * nsample + k) * c + l];
            }
}

/* ---- threenn_cpu, tf_interpolate.cpp:60-103 -------------------------------
 * float distance widened to double (:73), strict < insertion (:74-89),
 * bests start at 1e40 (:66) which becomes +inf when stored as float (:91-95). */
void gspn_oracle_three_nn(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx) {
    for (int i = 0; i < b; ++i) {
        const float *u = xyz1 + (size_t)i * n * 3;
        const float *kn = xyz2 + (size_t)i * m * 3;
        for (int j = 0; j < n; ++j) {
            float x1 = u[j * 3 + 0], y1 = u[j * 3 + 1], z1 = u[j * 3 + 2];
            double best1 = 1e40, best2 = 1e40, best3 = 1e40;
            int besti1 = 0, besti2 = 0, besti3 = 0;
            for (int k = 0; k < m; ++k) {
                double d = sqdist_cpu(kn[k * 3 + 0], kn[k * 3 + 1], kn[k * 3 + 2], x1, y1, z1);
                if (d < best1) {
                    best3 = best2; besti3 = besti2; best2 = best1; besti2 = besti1; best1 = d; besti1 = k;
                } else if (d < best2) {
                    best3 = best2; besti3 = besti2; best2 = d; besti2 = k;
                } else if (d < best3) {
                    best3 = d; besti3 = k;
                }
            }
            size_t o = ((size_t)i * n + j) * 3;
            dist[o + 0] = (float)best1; idx[o + 0] = besti1;
            dist[o + 1] = (float)best2; idx[o + 1] = besti2;
            dist[o + 2] = (float)best3; idx[o + 2] = besti3;
        }
    }
}

/* ---- threeinterpolate_cpu, tf_interpolate.cpp:107-127: (p1*w1+p2*w2)+p3*w3. */
void gspn_oracle_three_interpolate(int b, int m, int c, int n, const float *points, const int *idx, const float *weight, float *out) {
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < n; ++j) {
            size_t o = ((size_t)i * n + j) * 3;
            float w1 = weight[o], w2 = weight[o + 1], w3 = weight[o + 2];
            const float *p1 = points + ((size_t)i * m + idx[o]) * c;
            const float *p2 = points + ((size_t)i * m + idx[o + 1]) * c;
            const float *p3 = points + ((size_t)i * m + idx[o + 2]) * c;
            float *dst = out + ((size_t)i * n + j) * c;
            for (int l = 0; l < c; ++l) {
                float a = p1[l] * w1, bb = p2[l] * w2, cc = p3[l] * w3;
                float s = a + bb;
                dst[l] = s + cc;
            }
        }
}

/* ---- threeinterpolate_grad_cpu, tf_interpolate.cpp:131-153 (serial order). */
void gspn_oracle_three_interpolate_grad(int b, int n, int c, int m, const float *grad_out, const int *idx, const float *weight, float *grad_points) {
    memset(grad_points, 0, sizeof(float) * (size_t)b * m * c);
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < n; ++j) {
            size_t o = ((size_t)i * n + j) * 3;
            float w1 = weight[o], w2 = weight[o + 1], w3 = weight[o + 2];
            float *g1 = grad_points + ((size_t)i * m + idx[o]) * c;
            float *g2 = grad_points + ((size_t)i * m + idx[o + 1]) * c;
            float *g3 = grad_points + ((size_t)i * m + idx[o + 2]) * c;
            const float *go = grad_out + ((size_t)i * n + j) * c;
            for (int l = 0; l < c; ++l) {
                float a = go[l] * w1; g1[l] += a;
                float bb = go[l] * w2; g2[l] += bb;
                float cc = go[l] * w3; g3[l] += cc;
            }
        }
}

/* ---- nnsearch, tf_nndistance.cpp:21-43: 1-NN, strict <, first wins ties.
 * variant 0 = CPU rounding (the op TF-CPU runs), variant 1 = the rounding of
 * the compiled NmDistanceKernel (tf_nndistance_g.cu:5-127; same strict < and
 * ascending visiting order, so lowest index wins ties there too). */
static void nn_one_way(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx, int variant) {
    for (int i = 0; i < b; ++i)
        for (int j = 0; j < n; ++j) {
            const float *a = xyz1 + ((size_t)i * n + j) * 3;
            float best = 0.0f;
            int besti = 0;
            for (int k = 0; k < m; ++k) {
                const float *q = xyz2 + ((size_t)i * m + k) * 3;
                float d = variant ? sqdist_gpu(q[0], q[1], q[2], a[0], a[1], a[2])
                                  : sqdist_cpu(q[0], q[1], q[2], a[0], a[1], a[2]);
                if (k == 0 || d < best) { best = d; besti = k; }
            }
            dist[(size_t)i * n + j] = best;
            idx[(size_t)i * n + j] = besti;
        }
}

void gspn_oracle_nn_distance(int b, int n, int m, const float *xyz1, const float *xyz2,
                             float *dist1, int *idx1, float *dist2, int *idx2, int gpu_variant) {
    nn_one_way(b, n, m, xyz1, xyz2, dist1, idx1, gpu_variant);
    nn_one_way(b, m, n, xyz2, xyz1, dist2, idx2, gpu_variant);
}

/* ---- NnDistanceGradOp (CPU), tf_nndistance.cpp:126-163 (serial order). */
void gspn_oracle_nn_distance_grad(int b, int n, int m, const float *xyz1, const float *xyz2,
                                  const float *grad_dist1, const int *idx1, const float *grad_dist2, const int *idx2,
                                  float *grad_xyz1, float *grad_xyz2) {
    memset(grad_xyz1, 0, sizeof(float) * (size_t)b * n * 3);
    memset(grad_xyz2, 0, sizeof(float) * (size_t)b * m * 3);
    for (int i = 0; i < b; ++i) {
        for (int j = 0; j < n; ++j) {
            const float *a = xyz1 + ((size_t)i * n + j) * 3;
            int j2 = idx1[(size_t)i * n + j];
            const float *q = xyz2 + ((size_t)i * m + j2) * 3;
            float g = grad_dist1[(size_t)i * n + j] * 2;
            for (int l = 0; l < 3; ++l) {
                float v = g * (a[l] - q[l]);
                grad_xyz1[((size_t)i * n + j) * 3 + l] += v;
                grad_xyz2[((size_t)i * m + j2) * 3 + l] -= v;
            }
        }
        for (int j = 0; j < m; ++j) {
            const float *a = xyz2 + ((size_t)i * m + j) * 3;
            int j2 = idx2[(size_t)i * m + j];
            const float *q = xyz1 + ((size_t)i * n + j2) * 3;
            float g = grad_dist2[(size_t)i * m + j] * 2;
            for (int l = 0; l < 3; ++l) {
                float v = g * (a[l] - q[l]);
                grad_xyz2[((size_t)i * m + j) * 3 + l] += v;
                grad_xyz1[((size_t)i * n + j2) * 3 + l] -= v;
            }
        }
    }
}

/* ---- shared MLP layer: conv2d 1x1 + bias (tf_util.py:170-176) -> batch norm
 * in inference form (tf_util.py:530-534; tf.nn.batch_normalization with the
 * tf.contrib.layers.batch_norm default eps 1e-3) -> ReLU (:183-184).
 * x: (rows,cin) row-major, w: (cin,cout) (the [1,1,Cin,Cout] TF kernel),
 * bn == NULL skips normalisation (bn=False call sites), relu flag as
 * activation_fn.  Textbook fp32, accumulation in ascending cin order. */
void gspn_oracle_mlp_layer(long rows, int cin, int cout, const float *x, const float *w, const float *bias,
                           const float *gamma, const float *beta, const float *mean, const float *var,
                           int relu, float *y) {
    float *scale = (float *)malloc(sizeof(float) * cout), *shift = (float *)malloc(sizeof(float) * cout);
    for (int o = 0; o < cout; ++o) {
        if (gamma) {
            float inv = gamma[o] / sqrtf(var[o] + 1e-3f);
            scale[o] = inv;
            shift[o] = beta[o] - mean[o] * inv;
        } else { scale[o] = 1.0f; shift[o] = 0.0f; }
    }
    for (long r = 0; r < rows; ++r) {
        const float *xr = x + (size_t)r * cin;
        float *yr = y + (size_t)r * cout;
        for (int o = 0; o < cout; ++o) yr[o] = 0.0f;
        for (int c = 0; c < cin; ++c) {
            float xv = xr[c];
            const float *wr = w + (size_t)c * cout;
            for (int o = 0; o < cout; ++o) yr[o] += xv * wr[o];
        }
        for (int o = 0; o < cout; ++o) {
            float v = (yr[o] + bias[o]) * scale[o] + shift[o];
            yr[o] = (relu && v < 0.0f) ? 0.0f : v;
        }
    }
    free(scale); free(shift);
}

/* ---- tf.reduce_max over the nsample axis (pointnet_util.py:124). */
void gspn_oracle_max_over_k(long groups, int k, int c, const float *x, float *y) {
    for (long g = 0; g < groups; ++g)
        for (int o = 0; o < c; ++o) {
            float v = x[((size_t)g * k) * c + o];
            for (int s = 1; s < k; ++s) {
                float t = x[((size_t)g * k + s) * c + o];
                if (t > v) v = t;
            }
            y[(size_t)g * c + o] = v;
        }
}
