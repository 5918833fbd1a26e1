"""The reference's OWN CUDA kernels (oracle/_ref/libref_gpu.so, compiled unmodified
from /root/reference by oracle/Makefile) called on torch CUDA tensors.  Test infrastructure only."""
import ctypes

import torch

from oracle import oracle as O

_lib = None


def available():
    return O.ref_gpu_path() is not None and torch.cuda.is_available()


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(O.ref_gpu_path())
    return _lib


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _ok(rc):
    assert rc == 0, "reference kernel failed: cudaError %d" % rc


def fps(npoint, xyz):
    b, n, _ = xyz.shape
    temp = torch.empty((32, n), dtype=torch.float32, device=xyz.device)  # tf_sampling.cpp:115
    out = torch.empty((b, npoint), dtype=torch.int32, device=xyz.device)
    torch.cuda.synchronize()
    _ok(lib().ref_gpu_fps(b, n, npoint, _p(xyz), _p(temp), _p(out)))
    return out


def query_ball_point(radius, nsample, xyz1, xyz2):
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    idx = torch.full((b, m, nsample), -7, dtype=torch.int32, device=xyz1.device)  # unwritten rows stay -7
    cnt = torch.empty((b, m), dtype=torch.int32, device=xyz1.device)
    torch.cuda.synchronize()
    _ok(lib().ref_gpu_query_ball_point(b, n, m, ctypes.c_float(radius), nsample, _p(xyz1), _p(xyz2), _p(idx), _p(cnt)))
    return idx, cnt


def group_point(points, idx):
    b, n, c = points.shape
    _, m, k = idx.shape
    out = torch.empty((b, m, k, c), dtype=torch.float32, device=points.device)
    torch.cuda.synchronize()
    _ok(lib().ref_gpu_group_point(b, n, c, m, k, _p(points), _p(idx), _p(out)))
    return out


def gather_point(inp, idx):
    b, n, _ = inp.shape
    m = idx.shape[1]
    out = torch.empty((b, m, 3), dtype=torch.float32, device=inp.device)
    torch.cuda.synchronize()
    _ok(lib().ref_gpu_gather_point(b, n, m, _p(inp), _p(idx), _p(out)))
    return out


def nn_distance(xyz1, xyz2):
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dev = xyz1.device
    d1 = torch.empty((b, n), dtype=torch.float32, device=dev)
    i1 = torch.empty((b, n), dtype=torch.int32, device=dev)
    d2 = torch.empty((b, m), dtype=torch.float32, device=dev)
    i2 = torch.empty((b, m), dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    _ok(lib().ref_gpu_nn_distance(b, n, _p(xyz1), m, _p(xyz2), _p(d1), _p(i1), _p(d2), _p(i2)))
    return d1, i1, d2, i2
