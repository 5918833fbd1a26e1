"""PointNet++ set-abstraction / feature-propagation modules with the reference's signatures.

  sample_and_group(npoint, radius, nsample, xyz, points, tnet_spec=None, knn=False, use_xyz=True)
      utils/pointnet_util.py:17-54
  pointnet_sa_module(xyz, points, npoint, radius, nsample, mlp, mlp2, group_all, is_training,
                     bn_decay, scope, bn=True, pooling='max', tnet_spec=None, knn=False, use_xyz=True)
      utils/pointnet_util.py:85-139
  pointnet_fp_module(xyz1, xyz2, points1, points2, mlp, is_training, bn_decay, scope, bn=True, reuse=False)
      utils/pointnet_util.py:142-174

The reference builds a TF graph whose variables live in TF's variable store under `scope`.
Here `scope` keys a `VariableStore` (a dict of per-layer tensors named like the TF variables:
<scope>/conv<i>/{weights,biases,bn/gamma,bn/beta,bn/moving_mean,bn/moving_variance}); the first call
under a scope creates the variables (Xavier-uniform weights, zero biases, identity BN) exactly as
tf.get_variable would, later calls reuse them.  Three execution precisions share every index op:

  precision='bf16x3' (default): tcgen05 tensor-core MLP chain in split-bf16 arithmetic (hi*hi + lo*hi + hi*lo into the fp32
                     accumulator, ~2^-16): within 1e-3 of the reference's fp32 results (tests assert 1e-3, measured ~1e-5),
  precision='bf16' : the same chain with one bf16 product per term (what BASELINE.json's north_star names; unit round-off 2^-8,
                     so 5e-3 .. 3e-2 away from the fp32 reference -- OUTSIDE its 1e-3 bound; opt-in),
  precision='fp32' : reference-precision MLP on CUDA cores (parity tolerance 1e-5).

is_training=True (any mlp / mlp2 / group_all) runs the fp32 training form (gspn_b200/train.py): batch-statistics batch norm with in-place moving
average updates, autograd through MLP, max-pool, grouping and interpolation.
Not built: knn=True, tnet_spec (undefined `tnet` in the reference itself, pointnet_util.py:44), pooling other than
'max' (no call site in the model).  Those raise NotImplementedError instead of approximating.
"""
import math

import torch

from . import ops

BN_EPS = 1e-3  # tf.contrib.layers.batch_norm default epsilon (utils/tf_util.py:530-534)

DEFAULT_PRECISION = "bf16x3"  # tcgen05 chain, split-bf16 arithmetic: meets the 1e-3 float bound.  "bf16" and "fp32" are opt-in
TC_PRECISIONS = ("bf16x3", "bf16")


class VariableStore(dict):
    """scope -> list of layer dicts; the analogue of TF's global variable store."""

    def __init__(self, device="cuda", seed=0):
        super().__init__()
        self.device = device
        self.gen = torch.Generator(device="cpu")
        self.gen.manual_seed(seed)

    def layers(self, scope, prefix, cin, widths, bn):
        key = "%s/%s" % (scope, prefix)
        if key not in self:
            made = []
            c = cin
            for cout in widths:
                # tf.contrib.layers.xavier_initializer (uniform) on a [1,1,cin,cout] kernel (tf_util.py:29-33,162-167)
                lim = math.sqrt(6.0 / (c + cout))
                w = (torch.rand((c, cout), generator=self.gen, dtype=torch.float32) * 2 - 1) * lim
                layer = {"weights": w.to(self.device), "biases": torch.zeros(cout, device=self.device)}
                if bn:
                    layer.update(gamma=torch.ones(cout, device=self.device), beta=torch.zeros(cout, device=self.device),
                                 moving_mean=torch.zeros(cout, device=self.device), moving_variance=torch.ones(cout, device=self.device))
                made.append(layer)
                c = cout
            self[key] = made
        got = self[key]
        assert len(got) == len(widths) and (not got or got[0]["weights"].shape[0] == cin), \
            "variables under scope %r were created with a different shape" % key
        return got


VARIABLES = VariableStore()


def fold_layer(layer):
    """conv bias + inference batch norm as one affine: y = (x@W)*scale + shift.  Cached on the layer
    dict (re-folded when any of the tensors is replaced or modified in place)."""
    names = ("biases", "gamma", "beta", "moving_mean", "moving_variance")
    sig = tuple((layer[k].data_ptr(), layer[k]._version) for k in names if layer.get(k) is not None)
    cached = layer.get("_folded")
    if cached is not None and cached[0] == sig:
        return cached[1], cached[2]
    if layer.get("gamma") is not None:
        scale = layer["gamma"] / torch.sqrt(layer["moving_variance"] + BN_EPS)
        shift = (layer["biases"] - layer["moving_mean"]) * scale + layer["beta"]
    else:
        scale = torch.ones_like(layer["biases"])
        shift = layer["biases"]
    scale, shift = scale.contiguous(), shift.contiguous()
    layer["_folded"] = (sig, scale, shift)
    return scale, shift


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NULL = _Null()


def _stage(timers, name):
    """bench.py passes timers(name) -> context manager bracketing a stage with CUDA events."""
    return timers(name) if timers is not None else _NULL


def _check_unbuilt(is_training, knn=False, tnet_spec=None, pooling="max"):
    if knn:
        raise NotImplementedError("knn=True has no call site in the reference model (SURVEY.md 2.1 row 2)")
    if tnet_spec is not None:
        raise NotImplementedError("tnet_spec: `tnet` is undefined in the reference (utils/pointnet_util.py:44)")
    if pooling != "max":
        raise NotImplementedError("pooling=%r has no call site in the reference model; only 'max' is built" % pooling)


def sample_and_group(npoint, radius, nsample, xyz, points, tnet_spec=None, knn=False, use_xyz=True):
    """Reference-shaped outputs (new_xyz, new_points (b,m,K,3+c) with xyz FIRST, idx, grouped_xyz),
    composed from the 1:1 ops like utils/pointnet_util.py:36-48."""
    _check_unbuilt(False, knn, tnet_spec)
    new_xyz = ops.gather_point(xyz, ops.farthest_point_sample(npoint, xyz))
    idx, _ = ops.query_ball_point(radius, nsample, xyz, new_xyz)
    grouped_xyz = ops.group_point(xyz, idx) - new_xyz.unsqueeze(2)
    if points is not None:
        grouped_points = ops.group_point(points, idx)
        new_points = torch.cat([grouped_xyz, grouped_points], dim=-1) if use_xyz else grouped_points
    else:
        new_points = grouped_xyz
    return new_xyz, new_points, idx, grouped_xyz


def sample_and_group_all(xyz, points, use_xyz=True):
    """utils/pointnet_util.py:57-82."""
    b, n, _ = xyz.shape
    new_xyz = torch.zeros((b, 1, 3), dtype=torch.float32, device=xyz.device)
    idx = torch.arange(n, dtype=torch.int32, device=xyz.device).reshape(1, 1, n).repeat(b, 1, 1)
    grouped_xyz = xyz.reshape(b, 1, n, 3)
    if points is not None:
        new_points = (torch.cat([xyz, points], dim=2) if use_xyz else points).unsqueeze(1)
    else:
        new_points = grouped_xyz
    return new_xyz, new_points, idx, grouped_xyz


def _run_mlp_f32(x2d, layers, pool_last=1):
    for i, layer in enumerate(layers):
        scale, shift = fold_layer(layer)
        pool = pool_last if i == len(layers) - 1 else 1
        x2d = ops.mlp_layer_f32(x2d, layer["weights"], scale, shift, relu=True, pool=pool)
    return x2d


def _pool_only(grouped, nsample, c, use_xyz, has_points):
    """mlp == []: tf.reduce_max of the grouped rows themselves, columns back in the reference's order [xyz | features]
    (pointnet_util.py:48,124); the fused grouping kernel writes [features | xyz]."""
    x = ops.mlp_pool(grouped, nsample)
    if not has_points:
        return x
    return torch.cat([x[:, c:c + 3], x[:, :c]], dim=1).contiguous() if use_xyz else x[:, :c].contiguous()


def _features_first(layer, c, use_xyz, has_points):
    """First-layer dict whose kernel rows are re-ordered for the fused grouping layout [features | xyz]
    (reference order is [xyz | features], pointnet_util.py:48).  Cached on the layer dict."""
    w = layer["weights"]
    if not has_points:
        return layer  # rows are xyz only
    sig = (w.data_ptr(), w._version, use_xyz)
    cached = layer.get("_ff")
    if cached is None or cached[0] != sig:
        if use_xyz:
            wp = torch.cat([w[3:], w[:3]], dim=0).contiguous()
        else:
            wp = torch.cat([w, torch.zeros((3, w.shape[1]), dtype=w.dtype, device=w.device)], dim=0)  # xyz columns ignored
        layer["_ff"] = cached = (sig, wp)
    first = dict(layer)
    first["weights"] = cached[1]
    return first


def pointnet_sa_module(xyz, points, npoint, radius, nsample, mlp, mlp2, group_all, is_training, bn_decay, scope, bn=True,
                       pooling="max", tnet_spec=None, knn=False, use_xyz=True, variables=None, precision=None, timers=None):
    """-> (new_xyz (b,m,3), new_points (b,m,mlp[-1] or mlp2[-1]), idx (b,m,nsample) int32)."""
    _check_unbuilt(is_training, knn, tnet_spec, pooling)
    store = VARIABLES if variables is None else variables
    precision = precision or DEFAULT_PRECISION
    b, n, _ = xyz.shape
    c = 0 if points is None else points.shape[2]
    cin = (3 + c if (use_xyz or points is None) else c)
    layers = store.layers(scope, "conv", cin, list(mlp), bn)
    layers2 = store.layers(scope, "conv_post_", mlp[-1] if mlp else cin, list(mlp2 or []), bn)

    if is_training:
        from . import train
        return train.sa_module_train(xyz, points, npoint, radius, nsample, layers, bn_decay, use_xyz, layers2, group_all)
    if group_all:
        new_xyz, new_points, idx, _ = sample_and_group_all(xyz, points, use_xyz)
        x = _run_mlp_f32(new_points.reshape(b * n, cin).contiguous(), layers)
        x = ops.mlp_pool(x, n)
        m = 1
    else:
        with _stage(timers, scope + ":fps"):
            fps_idx = ops.farthest_point_sample(npoint, xyz)
        with _stage(timers, scope + ":gather"):
            new_xyz = ops.gather_point(xyz, fps_idx)
        m = npoint
        if precision in TC_PRECISIONS:
            from . import mlp_tc
            idx, x = mlp_tc.sa_group_mlp_max(xyz, new_xyz, points, radius, nsample, layers, use_xyz, store, scope, timers, precision)
        elif precision != "fp32":
            raise ValueError("precision must be one of 'bf16x3', 'bf16', 'fp32'")
        else:
            with _stage(timers, scope + ":ballquery_group"):
                idx, _, grouped, _ = ops.ballquery_group(radius, nsample, xyz, new_xyz, points, torch.float32)
            with _stage(timers, scope + ":mlp"):
                if layers:
                    first = _features_first(layers[0], c, use_xyz, points is not None)
                    x = _run_mlp_f32(grouped, [first] + list(layers[1:]), pool_last=nsample)
                else:
                    x = _pool_only(grouped, nsample, c, use_xyz, points is not None)
    x = _run_mlp_f32(x, layers2)
    return new_xyz, x.reshape(b, m, x.shape[-1]), idx


def pointnet_fp_module(xyz1, xyz2, points1, points2, mlp, is_training, bn_decay, scope, bn=True, reuse=False, variables=None,
                       precision=None, timers=None, half_output=None, f32_output=True):
    """-> new_points1 (b,n,mlp[-1])  (or the concatenated (b,n,c2+c1) map when mlp == []).
    half_output (extension, tensor-core paths): torch.float16 / torch.bfloat16 -> return (f32 map, 16-bit map), the chain's epilogue
    emits the copy; f32_output=False then skips the fp32 map (returns (None, 16-bit map))."""
    _check_unbuilt(is_training)
    store = VARIABLES if variables is None else variables
    precision = precision or DEFAULT_PRECISION
    b, n, _ = xyz1.shape
    c2 = points2.shape[2]
    c1 = 0 if points1 is None else points1.shape[2]
    layers = store.layers(scope, "conv_", c1 + c2, list(mlp), bn)
    if is_training:
        from . import train
        return train.fp_module_train(xyz1, xyz2, points1, points2, layers, bn_decay)
    with _stage(timers, scope + ":three_nn"):
        _, idx, weight = ops.three_nn(xyz1, xyz2, return_weight=True)
    if precision in TC_PRECISIONS and layers:
        from . import mlp_tc
        return mlp_tc.fp_interp_mlp(points1, points2, idx, weight, layers, store, scope, timers, precision, want_half=half_output,
                                    want_f32=f32_output or half_output is None)
    if precision not in TC_PRECISIONS and precision != "fp32":
        raise ValueError("precision must be one of 'bf16x3', 'bf16', 'fp32'")
    with _stage(timers, scope + ":interpolate"):
        interpolated = ops.three_interpolate(points2, idx, weight)
        new_points1 = torch.cat([interpolated, points1], dim=2) if points1 is not None else interpolated
    if not layers:
        return new_points1
    with _stage(timers, scope + ":mlp"):
        x = _run_mlp_f32(new_points1.reshape(b * n, c1 + c2), layers)
    x = x.reshape(b, n, x.shape[-1])
    return (x, x.to(half_output)) if half_output is not None else x
