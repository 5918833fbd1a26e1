"""Host side of the tcgen05 shared-MLP chain (csrc/mlp_tc.cu): weight-image cache and the two
fused module bodies used by pointnet_util when precision='bf16'.

  sa_group_mlp_max : fused ball-query+group (bf16 tile image) -> MLP chain -> max over nsample
                     (utils/pointnet_util.py:40-48,109-113,124)
  fp_interp_mlp    : three_interpolate + concat (bf16 tile image) -> MLP chain
                     (utils/pointnet_util.py:161-172)
"""
import ctypes

import torch

from . import _lib, ops
from ._lib import check

MAX_LAYERS = 4


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _pad64(c):
    return ((c + 63) // 64) * 64


def _packed(layer, cin_padded, row_perm=None, tag="w"):
    """bf16 K-major 128B-swizzled weight image of one layer, cached on the layer dict."""
    key = "_img_%s_%d" % (tag, cin_padded)
    w = layer["weights"]
    cached = layer.get(key)
    if cached is not None and cached[0] == (w.data_ptr(), w._version):
        return cached[1]
    cin, cout = w.shape
    L = _lib.lib()
    nbytes = L.gspn_mlp_weight_image_bytes(cin_padded, cout)
    if nbytes == 0:
        raise _lib.GspnError("mlp weight image: cout=%d must be a multiple of 8" % cout)
    img = torch.empty((nbytes,), dtype=torch.uint8, device=w.device)
    perm = None
    if row_perm is not None:
        perm = torch.tensor(row_perm, dtype=torch.int32, device=w.device)
    check(L.gspn_mlp_pack_weights(cin, cin_padded, cout, w.contiguous().data_ptr(), None if perm is None else perm.data_ptr(),
                                  img.data_ptr(), _stream()), "mlp_pack_weights")
    layer[key] = ((w.data_ptr(), w._version), img)
    return img


def mlp_chain(a_img, rows, k0, layers, first_perm, pool, want_bf16=False):
    """Run the whole layer chain on the bf16 tile image `a_img` ((rows padded to 128) x k0).
    Returns (out_f32 (rows/pool, cout), out_bf16 or None)."""
    from .pointnet_util import fold_layer
    L = _lib.lib()
    n = len(layers)
    assert 1 <= n <= MAX_LAYERS
    dims = [k0] + [l["weights"].shape[1] for l in layers]
    imgs, scales, shifts = [], [], []
    for i, layer in enumerate(layers):
        kp = k0 if i == 0 else _pad64(dims[i])
        imgs.append(_packed(layer, kp, first_perm if i == 0 else None, tag="first" if i == 0 else "w"))
        sc, sh = fold_layer(layer)
        scales.append(sc)
        shifts.append(sh)
    dev = a_img.device
    cout = dims[-1]
    out = torch.empty((rows // pool, cout), dtype=torch.float32, device=dev)
    out_h = torch.empty((rows // pool, cout), dtype=torch.bfloat16, device=dev) if want_bf16 else None
    arr_i = (ctypes.c_int * (n + 1))(*dims)
    arr_w = (ctypes.c_void_p * n)(*[t.data_ptr() for t in imgs])
    arr_s = (ctypes.c_void_p * n)(*[t.data_ptr() for t in scales])
    arr_b = (ctypes.c_void_p * n)(*[t.data_ptr() for t in shifts])
    arr_r = (ctypes.c_int * n)(*([1] * n))
    check(L.gspn_mlp_chain(rows, n, ctypes.cast(arr_i, ctypes.c_void_p), a_img.data_ptr(), ctypes.cast(arr_w, ctypes.c_void_p),
                           ctypes.cast(arr_s, ctypes.c_void_p), ctypes.cast(arr_b, ctypes.c_void_p), ctypes.cast(arr_r, ctypes.c_void_p),
                           pool, out.data_ptr(), None if out_h is None else out_h.data_ptr(), _stream()), "mlp_chain")
    return out, out_h


def mlp_chain_gather(xyz, new_xyz, shift, points, idx, layers, first_perm, pool):
    """Layer chain whose first operand is gathered in-kernel from the ball-query indices (c+3 <= 8): no grouped tensor
    in HBM.  Returns out_f32 (b*m*nsample/pool, cout)."""
    from .pointnet_util import fold_layer
    L = _lib.lib()
    n_layers = len(layers)
    b, n, _ = xyz.shape
    _, m, k = idx.shape
    c = 0 if points is None else points.shape[2]
    dims = [64] + [l["weights"].shape[1] for l in layers]
    imgs, scales, shifts = [], [], []
    for i, layer in enumerate(layers):
        kp = 64 if i == 0 else _pad64(dims[i])
        imgs.append(_packed(layer, kp, first_perm if i == 0 else None, tag="first" if i == 0 else "w"))
        sc, sh = fold_layer(layer)
        scales.append(sc)
        shifts.append(sh)
    rows = b * m * k
    out = torch.empty((rows // pool, dims[-1]), dtype=torch.float32, device=xyz.device)
    arr_i = (ctypes.c_int * (n_layers + 1))(*dims)
    arr_w = (ctypes.c_void_p * n_layers)(*[t.data_ptr() for t in imgs])
    arr_s = (ctypes.c_void_p * n_layers)(*[t.data_ptr() for t in scales])
    arr_b = (ctypes.c_void_p * n_layers)(*[t.data_ptr() for t in shifts])
    arr_r = (ctypes.c_int * n_layers)(*([1] * n_layers))
    check(L.gspn_mlp_chain_gather(b, n, m, k, c, xyz.data_ptr(), new_xyz.data_ptr(), None if shift is None else shift.data_ptr(),
                                  None if points is None else points.data_ptr(), idx.data_ptr(), n_layers,
                                  ctypes.cast(arr_i, ctypes.c_void_p), ctypes.cast(arr_w, ctypes.c_void_p), ctypes.cast(arr_s, ctypes.c_void_p),
                                  ctypes.cast(arr_b, ctypes.c_void_p), ctypes.cast(arr_r, ctypes.c_void_p), pool, out.data_ptr(), None,
                                  _stream()), "mlp_chain_gather")
    return out


GATHER_IN_CHAIN = True  # narrow rows (c+3 <= 8): build the first operand inside the chain kernel instead of a tile image


def gather_ok(points):
    c = 0 if points is None else points.shape[2]
    return GATHER_IN_CHAIN and c + 3 <= 8 and (points is None or points.dtype == torch.float32)


def tc_supported(layers, pool):
    if not (1 <= len(layers) <= MAX_LAYERS):
        return False
    if any(l["weights"].shape[1] % 32 or l["weights"].shape[1] > 512 for l in layers):
        return False
    return pool == 1 or pool % 32 == 0


def sa_group_mlp_max(xyz, new_xyz, points, radius, nsample, layers, use_xyz, store, scope, timers):
    """-> (idx (b,m,nsample) int32, pooled features (b*m, cout) f32)."""
    from .pointnet_util import _stage, _run_mlp_f32, _features_first
    b, m, _ = new_xyz.shape
    c = 0 if points is None else points.shape[2]
    rows = b * m * nsample
    if not tc_supported(layers, nsample):
        # widths the tensor-core chain does not take: same kernels as precision='fp32'
        with _stage(timers, scope + ":ballquery_group"):
            idx, _, grouped, _ = ops.ballquery_group(radius, nsample, xyz, new_xyz, points, torch.float32)
        with _stage(timers, scope + ":mlp"):
            first = _features_first(layers[0], c, use_xyz, points is not None)
            return idx, _run_mlp_f32(grouped, [first] + list(layers[1:]), pool_last=nsample)
    ld = 64 if gather_ok(points) else None
    if ld is None:
        with _stage(timers, scope + ":ballquery_group"):
            idx, _, img, ld = ops.ballquery_group(radius, nsample, xyz, new_xyz, points, torch.bfloat16)
    # tile-image columns are [features(c) | xyz(3) | 0]; the reference's kernel rows are [xyz | features]
    if points is None:
        perm = [0, 1, 2] + [-1] * (ld - 3)
    elif use_xyz:
        perm = [3 + k for k in range(c)] + [0, 1, 2] + [-1] * (ld - c - 3)
    else:
        perm = list(range(c)) + [-1] * (ld - c)
    if gather_ok(points):
        with _stage(timers, scope + ":ballquery_group"):
            idx, _ = ops.query_ball_point(radius, nsample, xyz, new_xyz)
        with _stage(timers, scope + ":mlp"):
            pts = None if points is None else points.contiguous()
            return idx, mlp_chain_gather(xyz.contiguous(), new_xyz.contiguous(), None, pts, idx, layers, perm, nsample)
    with _stage(timers, scope + ":mlp"):
        out, _ = mlp_chain(img, rows, ld, layers, perm, nsample)
    return idx, out


def fp_interp_mlp(points1, points2, idx, weight, layers, store, scope, timers, want_bf16=False):
    """-> (b,n,cout) f32  (and the same map in bf16 when want_bf16)."""
    from .pointnet_util import _stage, _run_mlp_f32
    b, n, _ = idx.shape
    m, c2 = points2.shape[1], points2.shape[2]
    c1 = 0 if points1 is None else points1.shape[2]
    if not tc_supported(layers, 1):
        with _stage(timers, scope + ":interpolate"):
            interp = ops.three_interpolate(points2, idx, weight)
            x = torch.cat([interp, points1], dim=2) if points1 is not None else interp
        with _stage(timers, scope + ":mlp"):
            y = _run_mlp_f32(x.reshape(b * n, c1 + c2), layers)
        return y.reshape(b, n, y.shape[-1])
    L = _lib.lib()
    ld = _pad64(c1 + c2)
    rows = b * n
    with _stage(timers, scope + ":interpolate"):
        nbytes = L.gspn_grouped_bytes(rows, c1 + c2, _lib.GSPN_DT_BF16)
        img = torch.empty((nbytes,), dtype=torch.uint8, device=points2.device)
        if rows % 128:
            img[-(ld // 64) * 16384:].zero_()
        p1 = None if points1 is None else points1.contiguous()
        check(L.gspn_fp_assemble(b, n, m, c1, c2, None if p1 is None else p1.data_ptr(), points2.contiguous().data_ptr(), idx.data_ptr(),
                                 weight.data_ptr(), img.data_ptr(), ld, _stream()), "fp_assemble")
    with _stage(timers, scope + ":mlp"):
        out, out_h = mlp_chain(img, rows, ld, layers, None, 1, want_bf16=want_bf16)
    if want_bf16:
        return out.reshape(b, n, out.shape[-1]), out_h.reshape(b, n, out_h.shape[-1])
    return out.reshape(b, n, out.shape[-1])
