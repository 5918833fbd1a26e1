"""numpy/ctypes front of the CPU oracle (oracle/gspn_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of gspn_oracle.c.  Nothing under
gspn_b200/ imports this module.  Function names, argument order and return
tuples follow the reference's Python op wrappers so the parity tests read like
the reference's own call sites:

  farthest_point_sample(npoint, inp)      tf_ops/sampling/tf_sampling.py:48-56
  gather_point(inp, idx)                  tf_ops/sampling/tf_sampling.py:29-37
  query_ball_point(radius, nsample, xyz1, xyz2)  tf_ops/grouping/tf_grouping.py:8-20
  group_point(points, idx)                tf_ops/grouping/tf_grouping.py:54-62
  three_nn(xyz1, xyz2)                    tf_ops/3d_interpolation/tf_interpolate.py:8-17
  three_interpolate(points, idx, weight)  tf_ops/3d_interpolation/tf_interpolate.py:19-28
  nn_distance(xyz1, xyz2)                 tf_ops/nn_distance/tf_nndistance.py:14-24
  sample_and_group / pointnet_sa_module / pointnet_fp_module   utils/pointnet_util.py:17,85,142
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libgspn_oracle.so")
_REF_CPU_PATH = os.path.join(_HERE, "_ref", "libref_cpu.so")
_REF_GPU_PATH = os.path.join(_HERE, "_ref", "libref_gpu.so")

_f = ctypes.POINTER(ctypes.c_float)
_i = ctypes.POINTER(ctypes.c_int)


def build(quiet=True):
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    out = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        if not _lib.gspn_oracle_has_fma():
            raise RuntimeError("oracle needs a CPU with FMA (fmaf must be a single rounding)")
    return _lib


def _fp(a):
    return a.ctypes.data_as(_f)


def _ip(a):
    return a.ctypes.data_as(_i)


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


# ----------------------------------------------------------------------------- ops
def farthest_point_sample(npoint, inp):
    inp = _f32(inp)
    b, n, _ = inp.shape
    out = np.zeros((b, npoint), np.int32)
    lib().gspn_oracle_fps(b, n, npoint, _fp(inp), _ip(out))
    return out


def gather_point(inp, idx):
    inp, idx = _f32(inp), _i32(idx)
    b, n, c = inp.shape
    m = idx.shape[1]
    out = np.empty((b, m, c), np.float32)
    lib().gspn_oracle_gather_point(b, n, m, c, _fp(inp), _ip(idx), _fp(out))
    return out


def gather_point_grad(inp, idx, out_g):
    inp, idx, out_g = _f32(inp), _i32(idx), _f32(out_g)
    b, n, c = inp.shape
    m = idx.shape[1]
    g = np.empty((b, n, c), np.float32)
    lib().gspn_oracle_gather_point_grad(b, n, m, c, _fp(out_g), _ip(idx), _fp(g))
    return g


def query_ball_point(radius, nsample, xyz1, xyz2):
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    idx = np.zeros((b, m, nsample), np.int32)
    cnt = np.zeros((b, m), np.int32)
    lib().gspn_oracle_query_ball_point(b, n, m, ctypes.c_float(radius), nsample, _fp(xyz1), _fp(xyz2), _ip(idx), _ip(cnt))
    return idx, cnt


def group_point(points, idx):
    points, idx = _f32(points), _i32(idx)
    b, n, c = points.shape
    _, m, k = idx.shape
    out = np.empty((b, m, k, c), np.float32)
    lib().gspn_oracle_group_point(b, n, c, m, k, _fp(points), _ip(idx), _fp(out))
    return out


def group_point_grad(points, idx, grad_out):
    points, idx, grad_out = _f32(points), _i32(idx), _f32(grad_out)
    b, n, c = points.shape
    _, m, k = idx.shape
    g = np.empty((b, n, c), np.float32)
    lib().gspn_oracle_group_point_grad(b, n, c, m, k, _fp(grad_out), _ip(idx), _fp(g))
    return g


def three_nn(xyz1, xyz2):
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dist = np.empty((b, n, 3), np.float32)
    idx = np.empty((b, n, 3), np.int32)
    lib().gspn_oracle_three_nn(b, n, m, _fp(xyz1), _fp(xyz2), _fp(dist), _ip(idx))
    return dist, idx


def three_interpolate(points, idx, weight):
    points, idx, weight = _f32(points), _i32(idx), _f32(weight)
    b, m, c = points.shape
    n = idx.shape[1]
    out = np.empty((b, n, c), np.float32)
    lib().gspn_oracle_three_interpolate(b, m, c, n, _fp(points), _ip(idx), _fp(weight), _fp(out))
    return out


def three_interpolate_grad(points, idx, weight, grad_out):
    points, idx, weight, grad_out = _f32(points), _i32(idx), _f32(weight), _f32(grad_out)
    b, m, c = points.shape
    n = idx.shape[1]
    g = np.empty((b, m, c), np.float32)
    lib().gspn_oracle_three_interpolate_grad(b, n, c, m, _fp(grad_out), _ip(idx), _fp(weight), _fp(g))
    return g


def nn_distance(xyz1, xyz2, gpu_variant=False):
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    d1 = np.empty((b, n), np.float32)
    i1 = np.empty((b, n), np.int32)
    d2 = np.empty((b, m), np.float32)
    i2 = np.empty((b, m), np.int32)
    lib().gspn_oracle_nn_distance(b, n, m, _fp(xyz1), _fp(xyz2), _fp(d1), _ip(i1), _fp(d2), _ip(i2), int(bool(gpu_variant)))
    return d1, i1, d2, i2


def nn_distance_grad(xyz1, xyz2, grad_dist1, idx1, grad_dist2, idx2):
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    grad_dist1, grad_dist2, idx1, idx2 = _f32(grad_dist1), _f32(grad_dist2), _i32(idx1), _i32(idx2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    g1 = np.empty((b, n, 3), np.float32)
    g2 = np.empty((b, m, 3), np.float32)
    lib().gspn_oracle_nn_distance_grad(b, n, m, _fp(xyz1), _fp(xyz2), _fp(grad_dist1), _ip(idx1), _fp(grad_dist2), _ip(idx2), _fp(g1), _fp(g2))
    return g1, g2


# ----------------------------------------------------------------------------- nearest-neighbour glue
def nearest_point(queries, refs, gpu_variant=False):
    """argmin over the dense squared-distance tensor (models/model_rpointnet.py:1136, :1032-1033; test.py:165-166):
    one direction of nn_distance, same loop and rounding (tf_nndistance.cpp:21-43), first index on ties as tf.argmin."""
    d1, i1, _, _ = nn_distance(queries, refs, gpu_variant)
    return d1, i1


def box_shrink(box, pc):
    """models/model_rpointnet.py:529-551 restated operation by operation in float32 numpy."""
    box, pc = _f32(box), _f32(pc)
    pc_aug = pc[:, None, :, :]                                   # (B,1,N,3)
    box_aug = box[:, :, None, :]                                 # (B,S,1,6)
    half = box_aug[..., 3:] / np.float32(2)
    m = np.logical_and(pc_aug >= (box_aug[..., :3] - half), pc_aug <= (box_aug[..., :3] + half))
    m = np.logical_and(np.logical_and(m[..., 0], m[..., 1]), m[..., 2])       # (B,S,N)
    out_mask = (np.float32(1) - m[..., None].astype(np.float32))             # (B,S,N,1)
    gamma = np.float32(1e4)
    box_max = np.max(pc_aug - gamma * out_mask, axis=2)                      # (B,S,3)
    box_min = np.min(pc_aug + gamma * out_mask, axis=2)
    out = np.concatenate(((box_max + box_min) / np.float32(2), box_max - box_min + np.float32(1e-3)), axis=2)
    keep = (box_max - box_min) > 0
    keep = np.logical_and(np.logical_and(keep[..., 0], keep[..., 1]), keep[..., 2])[..., None].astype(np.float32)
    return (out * keep).astype(np.float32)


# ----------------------------------------------------------------------------- shared MLP
def mlp_layer(x, layer, relu=True):
    """One tf_util.conv2d(1x1)+bias(+BN inference)+ReLU layer (tf_util.py:155-185,515-534).

    layer: dict with 'weights' (cin,cout), 'biases' (cout,), and -- when the call
    site has bn=True -- 'gamma','beta','moving_mean','moving_variance'."""
    x = _f32(x)
    shp = x.shape
    cin = shp[-1]
    w = _f32(layer["weights"])
    cout = w.shape[1]
    assert w.shape[0] == cin, (w.shape, cin)
    rows = int(np.prod(shp[:-1]))
    y = np.empty((rows, cout), np.float32)
    bias = _f32(layer["biases"])
    if layer.get("gamma") is not None:
        g, be, mu, var = (_f32(layer[k]) for k in ("gamma", "beta", "moving_mean", "moving_variance"))
        args = (_fp(g), _fp(be), _fp(mu), _fp(var))
    else:
        args = (None, None, None, None)
    lib().gspn_oracle_mlp_layer(ctypes.c_long(rows), cin, cout, _fp(x.reshape(rows, cin)), _fp(w), _fp(bias), *args, int(relu), _fp(y))
    return y.reshape(shp[:-1] + (cout,))


def mlp_layer_blas(x, layer, relu=True):
    """The same layer with the matrix product done by a BLAS-class fp32 GEMM (torch CPU matmul -> MKL/oneDNN sgemm): the stand-in
    for what TensorFlow-CPU (Eigen contraction) would run, used by bench.py's reference arm / cpu_baseline (TensorFlow itself is not
    installable here).  mlp_layer above (naive loop, strict left-to-right fp32) stays the CHECKER."""
    import torch
    x = _f32(x)
    shp = x.shape
    w = torch.from_numpy(_f32(layer["weights"]))
    y = torch.from_numpy(x.reshape(-1, shp[-1])) @ w + torch.from_numpy(_f32(layer["biases"]))
    if layer.get("gamma") is not None:
        g, be, mu, var = (torch.from_numpy(_f32(layer[k])) for k in ("gamma", "beta", "moving_mean", "moving_variance"))
        y = (y - mu) * (g / torch.sqrt(var + 1e-3)) + be
    if relu:
        y = torch.relu_(y)
    return y.numpy().reshape(shp[:-1] + (w.shape[1],))


def max_over_k(x):
    """tf.reduce_max(new_points, axis=[2]) (pointnet_util.py:124)."""
    x = _f32(x)
    b, m, k, c = x.shape
    y = np.empty((b, m, c), np.float32)
    lib().gspn_oracle_max_over_k(ctypes.c_long(b * m), k, c, _fp(x), _fp(y))
    return y


# ----------------------------------------------------------------------------- modules
def sample_and_group(npoint, radius, nsample, xyz, points, use_xyz=True, fps_idx=None):
    """utils/pointnet_util.py:17-54 (knn=False, tnet_spec=None)."""
    xyz = _f32(xyz)
    if fps_idx is None:
        fps_idx = farthest_point_sample(npoint, xyz)
    new_xyz = gather_point(xyz, fps_idx)
    idx, pts_cnt = query_ball_point(radius, nsample, xyz, new_xyz)
    grouped_xyz = group_point(xyz, idx)
    grouped_xyz = grouped_xyz - new_xyz[:, :, None, :]
    if points is not None:
        grouped_points = group_point(points, idx)
        new_points = np.concatenate([grouped_xyz, grouped_points], axis=-1) if use_xyz else grouped_points
    else:
        new_points = grouped_xyz
    return new_xyz, new_points, idx, grouped_xyz


def pointnet_sa_module(xyz, points, npoint, radius, nsample, mlp, params, use_xyz=True):
    """utils/pointnet_util.py:85-139, pooling='max', mlp2=None, group_all=False,
    is_training=False.  params: list of layer dicts ('conv%d' scopes in order)."""
    new_xyz, new_points, idx, _ = sample_and_group(npoint, radius, nsample, xyz, points, use_xyz=use_xyz)
    assert len(params) == len(mlp)
    for layer in params:
        new_points = mlp_layer(new_points, layer)
    return new_xyz, max_over_k(new_points), idx


def fp_weights(dist):
    """pointnet_util.py:157-160."""
    dist = np.maximum(_f32(dist), np.float32(1e-10))
    inv = (np.float32(1.0) / dist).astype(np.float32)
    norm = np.sum(inv, axis=2, keepdims=True, dtype=np.float32)
    return (inv / norm).astype(np.float32)


def pointnet_fp_module(xyz1, xyz2, points1, points2, mlp, params):
    """utils/pointnet_util.py:142-174 with is_training=False."""
    dist, idx = three_nn(xyz1, xyz2)
    weight = fp_weights(dist)
    interpolated = three_interpolate(points2, idx, weight)
    new_points1 = np.concatenate([interpolated, _f32(points1)], axis=2) if points1 is not None else interpolated
    assert len(params) == len(mlp)
    for layer in params:
        new_points1 = mlp_layer(new_points1, layer)
    return new_points1


def multi_encoding_net(xyz, points, fps_idx, radius_list, nsample_list, params_list, shift_pred=None, use_xyz=True):
    """models/model_rpointnet.py:28-77 with mlp_list2=[], output_shift=False, is_training=False.
    params_list[i]: layer dicts of radius i ('conv_prev_<i>_<j>' scopes)."""
    xyz = _f32(xyz)
    new_xyz = gather_point(xyz, fps_idx)
    outs = []
    for radius, nsample, params in zip(radius_list, nsample_list, params_list):
        idx, _ = query_ball_point(radius, nsample, xyz, new_xyz)
        grouped_xyz = group_point(xyz, idx) - new_xyz[:, :, None, :]
        if shift_pred is not None:
            grouped_xyz = grouped_xyz - _f32(shift_pred)[:, :, None, :]
        if points is not None:
            g = group_point(points, idx)
            if use_xyz:
                g = np.concatenate([g, grouped_xyz], axis=-1)  # points FIRST (:61)
        else:
            g = grouped_xyz
        for layer in params:
            g = mlp_layer(g, layer)
        outs.append(g.max(axis=2))
    return new_xyz, np.concatenate(outs, axis=-1)


# ----------------------------------------------------------------------------- oracle/_ref (the reference's own code)
_ref_cpu = None


def ref_cpu():
    """The reference's own CPU loops (oracle/_ref/libref_cpu.so) or None."""
    global _ref_cpu
    if _ref_cpu is None and os.path.exists(_REF_CPU_PATH):
        _ref_cpu = ctypes.CDLL(_REF_CPU_PATH)
    return _ref_cpu


def ref_three_nn(xyz1, xyz2):
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    dist = np.empty((b, n, 3), np.float32)
    idx = np.empty((b, n, 3), np.int32)
    ref_cpu().ref_threenn_cpu(b, n, m, _fp(xyz1), _fp(xyz2), _fp(dist), _ip(idx))
    return dist, idx


def ref_three_interpolate(points, idx, weight):
    points, idx, weight = _f32(points), _i32(idx), _f32(weight)
    b, m, c = points.shape
    n = idx.shape[1]
    out = np.empty((b, n, c), np.float32)
    ref_cpu().ref_threeinterpolate_cpu(b, m, c, n, _fp(points), _ip(idx), _fp(weight), _fp(out))
    return out


def ref_three_interpolate_grad(points, idx, weight, grad_out):
    points, idx, weight, grad_out = _f32(points), _i32(idx), _f32(weight), _f32(grad_out)
    b, m, c = points.shape
    n = idx.shape[1]
    g = np.zeros((b, m, c), np.float32)  # the op memsets first (tf_interpolate.cpp:258)
    ref_cpu().ref_threeinterpolate_grad_cpu(b, n, c, m, _fp(grad_out), _ip(idx), _fp(weight), _fp(g))
    return g


def ref_nn_distance(xyz1, xyz2):
    xyz1, xyz2 = _f32(xyz1), _f32(xyz2)
    b, n, _ = xyz1.shape
    m = xyz2.shape[1]
    d1 = np.empty((b, n), np.float32)
    i1 = np.empty((b, n), np.int32)
    d2 = np.empty((b, m), np.float32)
    i2 = np.empty((b, m), np.int32)
    ref_cpu().ref_nnsearch(b, n, m, _fp(xyz1), _fp(xyz2), _fp(d1), _ip(i1))
    ref_cpu().ref_nnsearch(b, m, n, _fp(xyz2), _fp(xyz1), _fp(d2), _ip(i2))
    return d1, i1, d2, i2


def ref_gpu_path():
    return _REF_GPU_PATH if os.path.exists(_REF_GPU_PATH) else None
