
// Suffix spliced after the reference loops: C-linkage doors for ctypes.
extern "C" {
void ref_threenn_cpu(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx) {
    threenn_cpu(b, n, m, xyz1, xyz2, dist, idx);
}
void ref_threeinterpolate_cpu(int b, int m, int c, int n, const float *points, const int *idx, const float *weight, float *out) {
    threeinterpolate_cpu(b, m, c, n, points, idx, weight, out);
}
void ref_threeinterpolate_grad_cpu(int b, int n, int c, int m, const float *grad_out, const int *idx, const float *weight, float *grad_points) {
    threeinterpolate_grad_cpu(b, n, c, m, grad_out, idx, weight, grad_points);
}
void ref_nnsearch(int b, int n, int m, const float *xyz1, const float *xyz2, float *dist, int *idx) {
    nnsearch(b, n, m, xyz1, xyz2, dist, idx);
}
}
