"""The SA x4 + FP x4 PointNet++ backbone of the reference's sem_net / shift_pred_net
(models/model_rpointnet.py:168-184 and :103-112) expressed with this package's drop-in modules.
It is the workload of BASELINE.json config 2 (bench.py) and of the end-to-end parity test; the
reference's heads (conv1d fc1/fc2, dropout) are outside the SA/FP hot path and not built.
"""
import numpy as np
import torch

from . import pointnet_util as pu

# (npoint, radius, nsample, mlp)  model_rpointnet.py:168-171
SA_SPECS = [
    (2048, 0.2, 32, [32, 32, 64]),
    (512, 0.4, 32, [64, 64, 128]),
    (128, 0.8, 32, [128, 128, 256]),
    (32, 1.6, 32, [256, 256, 512]),
]
# mlp of fa_layer1..4  model_rpointnet.py:181-184
FP_SPECS = [[256, 256], [256, 256], [256, 128], [128, 128, 128]]


def scaled_sa_specs(npoints):
    """SA specs for clouds smaller than 32768 points (tests): npoint scaled, never below 8."""
    f = npoints / 32768.0
    return [(max(8, int(round(m * f))), r, k, mlp) for (m, r, k, mlp) in SA_SPECS]


def random_variables(device, colour_channels=3, seed=7, bn_seed=8, sa_specs=SA_SPECS, fp_specs=FP_SPECS):
    """Xavier-uniform weights (seed 7) and random inference BN statistics (seed 8) for every layer
    (SURVEY.md 8d); returns (VariableStore, numpy copy for the oracle)."""
    store = pu.VariableStore(device=device, seed=seed)
    rng = np.random.RandomState(bn_seed)
    c = colour_channels
    chans = [c]
    for i, (_, _, _, mlp) in enumerate(sa_specs):
        store.layers("layer%d" % (i + 1), "conv", 3 + chans[-1], mlp, True)
        store.layers("layer%d" % (i + 1), "conv_post_", mlp[-1], [], True)
        chans.append(mlp[-1])
    # FP i interpolates level (4-i) features onto level (3-i): cin = c(points2) + c(points1)
    up = chans[4]
    for i, mlp in enumerate(fp_specs):
        c1 = chans[3 - i]
        store.layers("fa_layer%d" % (i + 1), "conv_", up + c1, mlp, True)
        up = mlp[-1]
    for key, layers in store.items():
        for layer in layers:
            co = layer["biases"].shape[0]
            layer["biases"] = torch.from_numpy((rng.randn(co) * 0.05).astype(np.float32)).to(device)
            layer["gamma"] = torch.from_numpy((0.75 + 0.5 * rng.rand(co)).astype(np.float32)).to(device)
            layer["beta"] = torch.from_numpy((rng.randn(co) * 0.05).astype(np.float32)).to(device)
            layer["moving_mean"] = torch.from_numpy((rng.randn(co) * 0.05).astype(np.float32)).to(device)
            layer["moving_variance"] = torch.from_numpy((0.75 + 0.5 * rng.rand(co)).astype(np.float32)).to(device)
    as_numpy = {k: [{n: v.cpu().numpy() for n, v in l.items()} for l in layers] for k, layers in store.items()}
    return store, as_numpy


def forward(xyz, colour, store, sa_specs=SA_SPECS, fp_specs=FP_SPECS, precision=None, timers=None, l0_half=None, l0_f32=True,
            is_training=False, bn_decay=None):
    """xyz (b,n,3), colour (b,n,c) CUDA f32 -> dict(l0_points (b,n,128), l1..l4 xyz/points, indices).
    l0_half: torch.float16 / torch.bfloat16 -> also "l0_points_half", a 16-bit copy of the per-point map written by the last chain's
    epilogue (l0_f32=False: only that copy).  timers: optional callable(name) -> context manager (bench.py stage brackets)."""
    def stage(name):
        return timers(name) if timers is not None else _null
    xs, ps, idxs = [xyz], [colour], []
    for i, (m, r, k, mlp) in enumerate(sa_specs):
        with stage("sa%d" % (i + 1)):
            nx, npts, idx = pu.pointnet_sa_module(xs[-1], ps[-1], m, r, k, mlp, None, False, is_training, bn_decay, "layer%d" % (i + 1),
                                                  variables=store, precision=precision, timers=timers)
        xs.append(nx)
        ps.append(npts)
        idxs.append(idx)
    up = ps[4]
    for i, mlp in enumerate(fp_specs):
        lvl = 3 - i
        want_h = l0_half if (i == len(fp_specs) - 1 and not is_training) else None
        with stage("fp%d" % (i + 1)):
            up = pu.pointnet_fp_module(xs[lvl], xs[lvl + 1], ps[lvl], up, mlp, is_training, bn_decay, "fa_layer%d" % (i + 1), variables=store,
                                       precision=precision, timers=timers, half_output=want_h, f32_output=l0_f32 or want_h is None)
    out = {"xyz": xs, "points": ps, "idx": idxs}
    if isinstance(up, tuple):
        out["l0_points"], out["l0_points_half"] = up
    else:
        out["l0_points"] = up
    return out


class _Null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_null = _Null()


def oracle_forward(O, xyz, colour, params, sa_specs=SA_SPECS, fp_specs=FP_SPECS):
    """The same backbone on the CPU oracle (tests / bench cpu_baseline only). O = oracle.oracle module."""
    xs, ps, idxs = [xyz], [colour], []
    for i, (m, r, k, mlp) in enumerate(sa_specs):
        nx, npts, idx = O.pointnet_sa_module(xs[-1], ps[-1], m, r, k, mlp, params["layer%d/conv" % (i + 1)])
        xs.append(nx)
        ps.append(npts)
        idxs.append(idx)
    up = ps[4]
    for i, mlp in enumerate(fp_specs):
        lvl = 3 - i
        up = O.pointnet_fp_module(xs[lvl], xs[lvl + 1], ps[lvl], up, mlp, params["fa_layer%d/conv_" % (i + 1)])
    return {"l0_points": up, "xyz": xs, "points": ps, "idx": idxs}
