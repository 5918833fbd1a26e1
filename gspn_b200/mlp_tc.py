"""Host side of the tcgen05 shared-MLP chain (csrc/mlp_tc.cu): weight-image cache and the fused module bodies used by
pointnet_util when precision is 'bf16x3' (default) or 'bf16'.

  sa_group_mlp_max : fused ball-query+group (tile image, or in-kernel gather for narrow rows) -> MLP chain -> max over nsample
                     (utils/pointnet_util.py:40-48,109-113,124)
  fp_interp_mlp    : three_interpolate + concat + MLP chain (utils/pointnet_util.py:161-172); with <= 4 skip-link channels the
                     interpolation is commuted with the first layer (gspn_mlp_chain_fp) and no interpolated map is written

Arithmetic (include/gspn_b200.h): 'bf16x3' = split-bf16 (hi*hi + lo*hi + hi*lo into the fp32 accumulator, error ~2^-16; inside the
reference's fp32 results to 1e-3), 'bf16' = one bf16 product per term (unit round-off 2^-8).
"""
import ctypes

import torch

from . import _lib, ops
from ._lib import check

MAX_LAYERS = 4
ARITH = {"bf16": _lib.GSPN_MLP_BF16, "bf16x3": _lib.GSPN_MLP_BF16X3}
IMAGE_DT = {"bf16": _lib.GSPN_DT_BF16, "bf16x3": _lib.GSPN_DT_BF16X2}
HALF_DT = {torch.bfloat16: _lib.GSPN_DT_BF16, torch.float16: _lib.GSPN_DT_F16}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _pad64(c):
    return ((c + 63) // 64) * 64


def _packed(layer, cin_padded, precision, row_perm=None, tag="w"):
    """K-major 128B-swizzled weight image of one layer (bf16, or the [hi | lo] pair for bf16x3), cached on the layer dict."""
    key = "_img_%s_%s_%d" % (precision, tag, cin_padded)
    w = layer["weights"]
    cached = layer.get(key)
    if cached is not None and cached[0] == (w.data_ptr(), w._version):
        return cached[1]
    w = ops._cuda_f32(w, "weights")
    cin, cout = w.shape
    L = _lib.lib()
    nbytes = L.gspn_mlp_weight_image_bytes(cin_padded, cout, ARITH[precision])
    if nbytes == 0:
        raise _lib.GspnError("mlp weight image: cout=%d must be a multiple of 8" % cout)
    img = torch.empty((nbytes,), dtype=torch.uint8, device=w.device)
    perm = None
    if row_perm is not None:
        perm = torch.tensor(row_perm, dtype=torch.int32, device=w.device)
    check(L.gspn_mlp_pack_weights(cin, cin_padded, cout, w.data_ptr(), None if perm is None else perm.data_ptr(), img.data_ptr(),
                                  ARITH[precision], _stream()), "mlp_pack_weights")
    layer[key] = ((layer["weights"].data_ptr(), layer["weights"]._version), img)
    return img


class _Args:
    """ctypes argument arrays of one chain launch (kept alive until the call returns)."""

    def __init__(self, dims, imgs, scales, shifts, relus):
        n = len(scales)
        self.dims = (ctypes.c_int * len(dims))(*dims)
        self.w = (ctypes.c_void_p * n)(*[None if t is None else t.data_ptr() for t in imgs])
        self.s = (ctypes.c_void_p * n)(*[t.data_ptr() for t in scales])
        self.b = (ctypes.c_void_p * n)(*[t.data_ptr() for t in shifts])
        self.r = (ctypes.c_int * n)(*relus)
        self.keep = (imgs, scales, shifts)

    def ptrs(self):
        c = ctypes.cast
        return c(self.dims, ctypes.c_void_p), c(self.w, ctypes.c_void_p), c(self.s, ctypes.c_void_p), c(self.b, ctypes.c_void_p), c(self.r, ctypes.c_void_p)


def _layer_args(layers, k0, precision, first_perm):
    from .pointnet_util import fold_layer
    dims = [k0] + [l["weights"].shape[1] for l in layers]
    imgs, scales, shifts = [], [], []
    for i, layer in enumerate(layers):
        kp = k0 if i == 0 else _pad64(dims[i])
        imgs.append(_packed(layer, kp, precision, first_perm if i == 0 else None, tag="first" if i == 0 else "w"))
        sc, sh = fold_layer(layer)
        scales.append(sc)
        shifts.append(sh)
    return dims, imgs, scales, shifts


def _outputs(rows, cout, dev, want_f32, want_half):
    out = torch.empty((rows, cout), dtype=torch.float32, device=dev) if want_f32 else None
    out_h = torch.empty((rows, cout), dtype=want_half, device=dev) if want_half is not None else None
    return out, out_h, (HALF_DT[want_half] if want_half is not None else _lib.GSPN_DT_BF16)


def mlp_chain(a_img, rows, k0, layers, first_perm, pool, precision="bf16x3", k0_used=0, want_half=None, want_f32=True, relus=None):
    """Run the whole layer chain on the tile image `a_img` ((rows padded to 128) x k0; GSPN_DT_BF16 image for 'bf16',
    GSPN_DT_BF16X2 image for 'bf16x3').  want_half: None, torch.bfloat16 or torch.float16 -> a 16-bit copy of the output.
    Returns (out_f32 (rows/pool, cout) or None, out_half or None)."""
    L = _lib.lib()
    n = len(layers)
    assert 1 <= n <= MAX_LAYERS
    dims, imgs, scales, shifts = _layer_args(layers, k0, precision, first_perm)
    out, out_h, hdt = _outputs(rows // pool, dims[-1], a_img.device, want_f32, want_half)
    args = _Args(dims, imgs, scales, shifts, relus or [1] * n)
    d, w, s, b, r = args.ptrs()
    check(L.gspn_mlp_chain(rows, n, d, k0_used, a_img.data_ptr(), w, s, b, r, pool, ops._p(out), ops._p(out_h), hdt, ARITH[precision],
                           _stream()), "mlp_chain")
    return out, out_h


def mlp_chain_gather(xyz, new_xyz, shift, points, idx, layers, first_perm, pool, precision="bf16x3"):
    """Layer chain whose first operand is gathered in-kernel from the ball-query indices (c+3 <= 8): no grouped tensor
    in HBM.  Returns out_f32 (b*m*nsample/pool, cout)."""
    L = _lib.lib()
    xyz, new_xyz = ops._cuda_f32(xyz.detach(), "xyz"), ops._cuda_f32(new_xyz.detach(), "new_xyz")
    shift = None if shift is None else ops._cuda_f32(shift.detach(), "shift_pred")
    points = None if points is None else ops._cuda_f32(points.detach(), "points")
    idx = ops._cuda_i32(idx, "idx")
    n_layers = len(layers)
    b, n, _ = xyz.shape
    _, m, k = idx.shape
    if shift is not None and tuple(shift.shape) != (b, m, 3):
        raise ValueError("shift_pred must be (b, npoint, 3)")
    c = 0 if points is None else points.shape[2]
    dims, imgs, scales, shifts = _layer_args(layers, 64, precision, first_perm)
    rows = b * m * k
    out = torch.empty((rows // pool, dims[-1]), dtype=torch.float32, device=xyz.device)
    args = _Args(dims, imgs, scales, shifts, [1] * n_layers)
    d, w, s, bb, r = args.ptrs()
    check(L.gspn_mlp_chain_gather(b, n, m, k, c, xyz.data_ptr(), new_xyz.data_ptr(), ops._p(shift), ops._p(points), idx.data_ptr(), n_layers,
                                  d, w, s, bb, r, pool, out.data_ptr(), None, _lib.GSPN_DT_BF16, ARITH[precision], _stream()),
          "mlp_chain_gather")
    return out


GATHER_IN_CHAIN = True  # narrow rows (c+3 <= 8): build the first operand inside the chain kernel instead of a tile image
FP_COMMUTE = True       # <= 4 skip-link channels: interpolate the pre-multiplied coarse features inside the chain kernel


def gather_ok(points):
    c = 0 if points is None else points.shape[2]
    return GATHER_IN_CHAIN and c + 3 <= 8 and (points is None or points.dtype == torch.float32)


def tc_supported(layers, pool):
    if not (1 <= len(layers) <= MAX_LAYERS):
        return False
    widths = [l["weights"].shape[1] for l in layers]
    if any(w % 32 or w > 512 for w in widths) or any(w > 256 for w in widths[:-1]):
        return False
    return pool == 1 or pool % 32 == 0


def sa_group_mlp_max(xyz, new_xyz, points, radius, nsample, layers, use_xyz, store, scope, timers, precision="bf16x3"):
    """-> (idx (b,m,nsample) int32, pooled features (b*m, cout) f32)."""
    from .pointnet_util import _stage, _run_mlp_f32, _features_first, _pool_only
    b, m, _ = new_xyz.shape
    c = 0 if points is None else points.shape[2]
    rows = b * m * nsample
    if not tc_supported(layers, nsample):
        # widths the tensor-core chain does not take (or no layers at all): same kernels as precision='fp32'
        with _stage(timers, scope + ":ballquery_group"):
            idx, _, grouped, _ = ops.ballquery_group(radius, nsample, xyz, new_xyz, points, torch.float32)
        with _stage(timers, scope + ":mlp"):
            if not layers:
                return idx, _pool_only(grouped, nsample, c, use_xyz, points is not None)
            first = _features_first(layers[0], c, use_xyz, points is not None)
            return idx, _run_mlp_f32(grouped, [first] + list(layers[1:]), pool_last=nsample)
    ld = 64 if gather_ok(points) else None
    if ld is None:
        with _stage(timers, scope + ":ballquery_group"):
            idx, _, img, ld = ops.ballquery_group(radius, nsample, xyz, new_xyz, points, "image:" + precision)
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
            return idx, mlp_chain_gather(xyz, new_xyz, None, points, idx, layers, perm, nsample, precision)
    with _stage(timers, scope + ":mlp"):
        out, _ = mlp_chain(img, rows, ld, layers, perm, nsample, precision, k0_used=c + 3)
    return idx, out


def _identity_affine(cout, dev):
    return torch.ones(cout, dtype=torch.float32, device=dev), torch.zeros(cout, dtype=torch.float32, device=dev)


def _fp_commuted(points1, points2, idx, weight, layers, precision, timers, scope, want_half, want_f32):
    """interp3(points2) @ W0[:c2] = interp3(points2 @ W0[:c2]): multiply the m known points once, gather inside the chain kernel."""
    from .pointnet_util import _stage, fold_layer
    L = _lib.lib()
    b, n, _ = idx.shape
    m, c2 = points2.shape[1], points2.shape[2]
    c1 = 0 if points1 is None else points1.shape[2]
    dev = points2.device
    first = layers[0]
    w = first["weights"]
    n0 = w.shape[1]
    sig = (w.data_ptr(), w._version)
    cached = first.get("_fp_split")
    if cached is None or cached[0] != sig:
        ones, zeros = _identity_affine(n0, dev)
        head = {"weights": w[:c2].contiguous(), "_folded_identity": (ones, zeros)}
        first["_fp_split"] = cached = (sig, head, w[c2:].contiguous() if c1 else None)
    head, w0b = cached[1], cached[2]
    with _stage(timers, scope + ":interpolate"):
        # y2 = points2 @ W0[:c2]  (b*m rows): rows -> tile image -> one-layer chain without affine / ReLU
        ld2 = _pad64(c2)
        rows2 = b * m
        img2 = torch.empty((L.gspn_grouped_bytes(rows2, c2, IMAGE_DT[precision]),), dtype=torch.uint8, device=dev)
        check(L.gspn_fp_assemble(b, m, 0, c2, 0, points2.data_ptr(), None, None, None, img2.data_ptr(), ld2, IMAGE_DT[precision], _stream()),
              "fp_assemble(rows)")
        ones, zeros = head["_folded_identity"]
        y2 = torch.empty((rows2, n0), dtype=torch.float32, device=dev)
        args = _Args([ld2, n0], [_packed(head, ld2, precision, None, tag="fp_head")], [ones], [zeros], [0])
        d, wp, s, bb, r = args.ptrs()
        check(L.gspn_mlp_chain(rows2, 1, d, c2, img2.data_ptr(), wp, s, bb, r, 1, y2.data_ptr(), None, _lib.GSPN_DT_BF16, ARITH[precision],
                               _stream()), "mlp_chain(y2)")
    with _stage(timers, scope + ":mlp"):
        dims = [c1 + c2] + [l["weights"].shape[1] for l in layers]
        imgs, scales, shifts = [None], [], []
        for i, layer in enumerate(layers):
            if i > 0:
                imgs.append(_packed(layer, _pad64(dims[i]), precision, None, tag="w"))
            sc, sh = fold_layer(layer)
            scales.append(sc)
            shifts.append(sh)
        out, out_h, hdt = _outputs(b * n, dims[-1], dev, want_f32, want_half)
        args = _Args(dims, imgs, scales, shifts, [1] * len(layers))
        d, wp, s, bb, r = args.ptrs()
        check(L.gspn_mlp_chain_fp(b, n, m, c1, y2.data_ptr(), idx.data_ptr(), weight.data_ptr(), ops._p(points1), ops._p(w0b), len(layers),
                                  d, wp, s, bb, r, ops._p(out), ops._p(out_h), hdt, ARITH[precision], _stream()), "mlp_chain_fp")
    return out, out_h


def fp_interp_mlp(points1, points2, idx, weight, layers, store, scope, timers, precision="bf16x3", want_half=None, want_f32=True):
    """-> (b,n,cout) f32, or (f32 map or None, 16-bit map) when want_half (torch.bfloat16 / torch.float16) is given."""
    from .pointnet_util import _stage, _run_mlp_f32
    points2 = ops._cuda_f32(points2.detach(), "points2")
    points1 = None if points1 is None else ops._cuda_f32(points1.detach(), "points1")
    idx, weight = ops._cuda_i32(idx, "idx"), ops._cuda_f32(weight.detach(), "weight")
    b, n, _ = idx.shape
    m, c2 = points2.shape[1], points2.shape[2]
    c1 = 0 if points1 is None else points1.shape[2]
    if not tc_supported(layers, 1):
        with _stage(timers, scope + ":interpolate"):
            interp = ops.three_interpolate(points2, idx, weight)
            x = torch.cat([interp, points1], dim=2) if points1 is not None else interp
        with _stage(timers, scope + ":mlp"):
            y = _run_mlp_f32(x.reshape(b * n, c1 + c2), layers)
        y = y.reshape(b, n, y.shape[-1])
        return (y, y.to(want_half)) if want_half is not None else y
    L = _lib.lib()
    rows = b * n
    widths = [l["weights"].shape[1] for l in layers]
    if FP_COMMUTE and len(layers) >= 2 and c1 <= 4 and widths[0] == 128 and n >= 2 * m and (b * n) < 2 ** 31:
        out, out_h = _fp_commuted(points1, points2, idx, weight, layers, precision, timers, scope, want_half, want_f32)
    else:
        ld = _pad64(c1 + c2)
        with _stage(timers, scope + ":interpolate"):
            nbytes = L.gspn_grouped_bytes(rows, c1 + c2, IMAGE_DT[precision])
            img = torch.empty((nbytes,), dtype=torch.uint8, device=points2.device)
            check(L.gspn_fp_assemble(b, n, m, c1, c2, ops._p(points1), points2.data_ptr(), idx.data_ptr(), weight.data_ptr(), img.data_ptr(), ld,
                                     IMAGE_DT[precision], _stream()), "fp_assemble")
        with _stage(timers, scope + ":mlp"):
            out, out_h = mlp_chain(img, rows, ld, layers, None, 1, precision, k0_used=c1 + c2, want_half=want_half, want_f32=want_f32)
    shape = (b, n, widths[-1])
    if want_half is not None:
        return (None if out is None else out.reshape(shape)), out_h.reshape(shape)
    return out.reshape(shape)
