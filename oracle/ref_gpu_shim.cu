// C-linkage doors onto the reference's own CUDA launchers, which are compiled
// unmodified from where they lie under /root/reference (oracle/Makefile):
//   tf_ops/sampling/tf_sampling_g.cu:203-211, tf_ops/grouping/tf_grouping_g.cu:186-199,
//   tf_ops/nn_distance/tf_nndistance_g.cu:128-157.
// All pointers are device pointers; the reference launches on the legacy default stream.
#include <cuda_runtime.h>
void farthestpointsamplingLauncher(int b, int n, int m, const float *inp, float *temp, int *out);
void gatherpointLauncher(int b, int n, int m, const float *inp, const int *idx, float *out);
void scatteraddpointLauncher(int b, int n, int m, const float *out_g, const int *idx, float *inp_g);
void queryBallPointLauncher(int b, int n, int m, float radius, int nsample, const float *xyz1, const float *xyz2, int *idx, int *pts_cnt);
void groupPointLauncher(int b, int n, int c, int m, int nsample, const float *points, const int *idx, float *out);
void groupPointGradLauncher(int b, int n, int c, int m, int nsample, const float *grad_out, const int *idx, float *grad_points);
void NmDistanceKernelLauncher(int b, int n, const float *xyz, int m, const float *xyz2, float *result, int *result_i, float *result2, int *result2_i);
void NmDistanceGradKernelLauncher(int b, int n, const float *xyz1, int m, const float *xyz2, const float *grad_dist1, const int *idx1, const float *grad_dist2, const int *idx2, float *grad_xyz1, float *grad_xyz2);

extern "C" {
int ref_gpu_fps(int b, int n, int m, const float *inp, float *temp, int *out) {
    farthestpointsamplingLauncher(b, n, m, inp, temp, out);
    return (int)cudaDeviceSynchronize();
}
int ref_gpu_gather_point(int b, int n, int m, const float *inp, const int *idx, float *out) {
    gatherpointLauncher(b, n, m, inp, idx, out);
    return (int)cudaDeviceSynchronize();
}
int ref_gpu_gather_point_grad(int b, int n, int m, const float *out_g, const int *idx, float *inp_g) {
    cudaMemset(inp_g, 0, sizeof(float) * (size_t)b * n * 3);  // tf_sampling.cpp:174
    scatteraddpointLauncher(b, n, m, out_g, idx, inp_g);
    return (int)cudaDeviceSynchronize();
}
int ref_gpu_query_ball_point(int b, int n, int m, float radius, int nsample, const float *xyz1, const float *xyz2, int *idx, int *pts_cnt) {
    queryBallPointLauncher(b, n, m, radius, nsample, xyz1, xyz2, idx, pts_cnt);
    return (int)cudaDeviceSynchronize();
}
int ref_gpu_group_point(int b, int n, int c, int m, int nsample, const float *points, const int *idx, float *out) {
    groupPointLauncher(b, n, c, m, nsample, points, idx, out);
    return (int)cudaDeviceSynchronize();
}
int ref_gpu_group_point_grad(int b, int n, int c, int m, int nsample, const float *grad_out, const int *idx, float *grad_points) {
    cudaMemset(grad_points, 0, sizeof(float) * (size_t)b * n * c);  // tf_grouping.cpp:234
    groupPointGradLauncher(b, n, c, m, nsample, grad_out, idx, grad_points);
    return (int)cudaDeviceSynchronize();
}
int ref_gpu_nn_distance(int b, int n, const float *xyz, int m, const float *xyz2, float *result, int *result_i, float *result2, int *result2_i) {
    NmDistanceKernelLauncher(b, n, xyz, m, xyz2, result, result_i, result2, result2_i);
    return (int)cudaDeviceSynchronize();
}
int ref_gpu_nn_distance_grad(int b, int n, const float *xyz1, int m, const float *xyz2, const float *grad_dist1, const int *idx1, const float *grad_dist2, const int *idx2, float *grad_xyz1, float *grad_xyz2) {
    NmDistanceGradKernelLauncher(b, n, xyz1, m, xyz2, grad_dist1, idx1, grad_dist2, idx2, grad_xyz1, grad_xyz2);
    return (int)cudaDeviceSynchronize();
}
}
