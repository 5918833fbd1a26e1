// Prefix spliced (by oracle/Makefile) in front of the reference's own CPU loops,
// which are streamed by line range straight from /root/reference into g++'s stdin:
//   threenn_cpu / threeinterpolate_cpu / threeinterpolate_grad_cpu  tf_ops/3d_interpolation/tf_interpolate.cpp:60-153
//   nnsearch                                                         tf_ops/nn_distance/tf_nndistance.cpp:21-43
// Nothing of the reference is written to disk; only oracle/_ref/libref_cpu.so is.
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cmath>
