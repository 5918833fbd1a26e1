"""The reference's own CUDA kernels (oracle/_ref/libref_gpu.so) on torch tensors: lives in oracle/ (test infrastructure), re-exported
here so the parity tests keep importing `refgpu`."""
from oracle.refgpu import *  # noqa: F401,F403
from oracle.refgpu import available, fps, gather_point, group_point, lib, nn_distance, query_ball_point  # noqa: F401
