"""Training form of the SA / FP modules (is_training=True): batch-statistics batch norm, autograd through the shared
MLP, max-pool, grouping and interpolation.  fp32 on CUDA cores (csrc/train_ops.cu + the fp32 GEMM of mlp_f32.cu);
torch supplies the autograd tape and a few c-length vector ops, the per-row work is all in this library's kernels.

Reference semantics: tf_util.conv2d -> tf.nn.bias_add -> tf.contrib.layers.batch_norm(center, scale, is_training,
decay=bn_decay or 0.9, updates_collections=None, epsilon default 1e-3) -> relu (utils/tf_util.py:170-184,515-534);
moments over every axis but channels; moving averages updated in place every step.
Gradients flow to the layer variables and to `points` (features), as in the reference's registered gradients
(GroupPointGrad, ThreeInterpolateGrad, GatherPointGrad); xyz, FPS, ball-query and three_nn indices carry none.

Data parallel: the reference normalises over the WHOLE batch on its single GPU.  When torch.distributed is initialised
with more than one rank and SYNC_BN is on (default), every batch-norm layer all-reduces [sum z, sum z^2, rows] in the forward and
[sum dy', sum dy' xhat] in the backward (2C+1 / 2C doubles), so a batch sharded over W ranks gives the same activations, moving
averages and -- after allreduce_gradients -- the same parameter gradients as the whole batch on one GPU
(tests/test_syncbn.py).  SYNC_BN = False is per-replica batch norm.
"""
import torch

from . import _lib, ops
from ._lib import check

BN_EPS = 1e-3
SYNC_BN = True      # whole-batch statistics across ranks (only acts when torch.distributed has > 1 rank)
SYNC_GROUP = None   # process group of the data-parallel replicas (None = the default group)
PEER_GROUP = None   # a gspn_b200.p2p.PeerGroup: the moment all-reduces go through NVLink peer memory instead of NCCL (use_peer_moments)
EQUAL_SHARDS = True # every rank feeds the same number of rows per layer (batch sharding of equal scenes): the global row count is
                    # rows * world and needs no communication -- and no host synchronisation inside the step


def _world():
    import torch.distributed as dist
    return dist.get_world_size(SYNC_GROUP) if (SYNC_BN and dist.is_available() and dist.is_initialized()) else 1


def use_peer_moments(group):
    """Route the batch-norm moment all-reduces through a p2p.PeerGroup (one small NVLink kernel per collective, no NCCL, no host
    synchronisation); None switches back to torch.distributed.  Needs EQUAL_SHARDS (the row count is not communicated)."""
    global PEER_GROUP
    PEER_GROUP = group


def allreduce_moments(s1, s2, rows, buf=None):
    """[sum, sum of squares, rows] of this rank's shard -> the same over every rank's shard: ONE all-reduce of 2C (+1) doubles.
    buf: optional (2C,) tensor whose halves ARE s1 and s2 (then nothing is concatenated and the result aliases it).
    Works on CUDA (NCCL / gloo) and CPU (gloo) tensors; returns (s1, s2, total_rows)."""
    import torch.distributed as dist
    world = _world()
    if world == 1:
        return s1, s2, rows
    c = s1.numel()
    if EQUAL_SHARDS:
        if buf is None:
            buf = torch.cat([s1.double(), s2.double()])
        if PEER_GROUP is not None and buf.is_cuda:
            PEER_GROUP.allreduce_(buf)
        else:
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=SYNC_GROUP)
        return buf[:c], buf[c:2 * c], rows * world
    buf = torch.cat([s1.double(), s2.double(), torch.tensor([float(rows)], dtype=torch.float64, device=s1.device)])
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=SYNC_GROUP)
    return buf[:c], buf[c:2 * c], int(round(float(buf[2 * c])))  # reads the count back: one host synchronisation per layer


def _s():
    return torch.cuda.current_stream().cuda_stream


def _linear(x, w, shift, y=None):
    """y = x @ w + shift  (rows,cin)x(cin,cout); fp32 CUDA-core GEMM (gspn_mlp_layer_f32 with scale=1, no activation)."""
    rows, cin = x.shape
    cout = w.shape[1]
    if y is None:
        y = torch.empty((rows, cout), dtype=torch.float32, device=x.device)
    ones = torch.ones(cout, dtype=torch.float32, device=x.device)
    check(_lib.lib().gspn_mlp_layer_f32(rows, cin, cout, x.data_ptr(), x.stride(0), w.data_ptr(), ones.data_ptr(), shift.data_ptr(), 0, 1,
                                        y.data_ptr(), _s()), "mlp_layer_f32")
    return y


class MlpLayerTrain(torch.autograd.Function):
    """act(bn_batch(x@W+b)) [+ max over groups of `pool` rows]; updates moving_mean / moving_variance in place."""

    @staticmethod
    def forward(ctx, x, w, bias, gamma, beta, moving_mean, moving_var, decay, pool, relu):
        L = _lib.lib()
        x = x.contiguous() if x.stride(1) != 1 else x
        w = w.contiguous()
        rows, cin = x.shape
        cout = w.shape[1]
        dev = x.device
        bn = gamma is not None
        z = _linear(x, w, bias.contiguous())
        if bn:
            mom = torch.empty(2 * cout, dtype=torch.float64, device=dev)
            s1, s2 = mom[:cout], mom[cout:]
            check(L.gspn_col_moments_f32(rows, cout, z.data_ptr(), s1.data_ptr(), s2.data_ptr(), _s()), "col_moments")
            s1, s2, total_rows = allreduce_moments(s1, s2, rows, mom)  # whole-batch moments across ranks (SyncBN); no-op on one rank
            mean64 = s1 / total_rows
            var64 = (s2 / total_rows - mean64 * mean64).clamp_(min=0.0)  # biased variance normalises (tf.nn.moments)
            mean, invstd = mean64.float(), torch.rsqrt(var64.float() + BN_EPS)
            with torch.no_grad():  # moving averages, updates_collections=None: updated as part of the forward
                moving_mean.mul_(decay).add_(mean * (1.0 - decay))
                moving_var.mul_(decay).add_(var64.float() * (1.0 - decay))
            g, be = gamma.contiguous(), beta.contiguous()
        else:
            total_rows = rows
            mean = torch.zeros(cout, device=dev)
            invstd = torch.ones(cout, device=dev)
            g, be = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
        y = torch.empty_like(z)
        check(L.gspn_bn_act_f32(rows, cout, z.data_ptr(), mean.data_ptr(), invstd.data_ptr(), g.data_ptr(), be.data_ptr(), int(relu), y.data_ptr(),
                                _s()), "bn_act")
        argmax = None
        out = y
        if pool > 1:
            groups = rows // pool
            out = torch.empty((groups, cout), dtype=torch.float32, device=dev)
            argmax = torch.empty((groups, cout), dtype=torch.int32, device=dev)
            check(L.gspn_maxpool_argmax_f32(groups, pool, cout, y.data_ptr(), out.data_ptr(), argmax.data_ptr(), _s()), "maxpool_argmax")
        ctx.save_for_backward(x, w, z, mean, invstd, g, be, argmax)
        ctx.meta = (pool, int(relu), bn, total_rows)
        return out

    @staticmethod
    def backward(ctx, dout):
        L = _lib.lib()
        x, w, z, mean, invstd, g, be, argmax = ctx.saved_tensors
        pool, relu, bn, total_rows = ctx.meta
        rows, cin = x.shape
        cout = w.shape[1]
        dev = x.device
        dout = dout.contiguous()
        mom = torch.empty(2 * cout, dtype=torch.float64, device=dev)
        s1, s2 = mom[:cout], mom[cout:]
        dz = torch.empty((rows, cout), dtype=torch.float32, device=dev)
        am = None if argmax is None else argmax.data_ptr()
        if bn and total_rows != rows:
            # SyncBN: dz needs the sums over the whole batch; dgamma / dbeta below stay this rank's share (the gradient all-reduce adds them up)
            check(L.gspn_bn_bwd_sums_f32(rows, cout, pool, relu, z.data_ptr(), dout.data_ptr(), am, mean.data_ptr(), invstd.data_ptr(), g.data_ptr(),
                                         be.data_ptr(), s1.data_ptr(), s2.data_ptr(), _s()), "bn_bwd_sums")
            dbeta_local, dgamma_local = s1.float(), s2.float()  # this rank's share, before the sums become global
            g1, g2, _ = allreduce_moments(s1, s2, rows, mom)
            check(L.gspn_bn_bwd_apply_f32(rows, total_rows, cout, pool, relu, z.data_ptr(), dout.data_ptr(), am, mean.data_ptr(), invstd.data_ptr(),
                                          g.data_ptr(), be.data_ptr(), g1.data_ptr(), g2.data_ptr(), dz.data_ptr(), _s()), "bn_bwd_apply")
        else:
            check(L.gspn_bn_act_pool_bwd_f32(rows, cout, pool, relu, int(bn), z.data_ptr(), dout.data_ptr(), am, mean.data_ptr(), invstd.data_ptr(),
                                             g.data_ptr(), be.data_ptr(), s1.data_ptr(), s2.data_ptr(), dz.data_ptr(), None, None, _s()),
                  "bn_act_pool_bwd")
        dW = torch.empty((cin, cout), dtype=torch.float32, device=dev)
        db = torch.empty(cout, dtype=torch.float32, device=dev)
        check(L.gspn_mlp_wgrad_f32(rows, cin, cout, x.data_ptr(), x.stride(0), dz.data_ptr(), dW.data_ptr(), db.data_ptr(), _s()), "mlp_wgrad")
        dx = None
        if ctx.needs_input_grad[0]:
            dx = _linear(dz, w.t().contiguous(), torch.zeros(cin, device=dev))
        if bn and total_rows != rows:
            dgamma, dbeta = dgamma_local, dbeta_local
        else:
            dgamma = s2.float() if bn else None
            dbeta = s1.float() if bn else None
        return dx, dW, db, dgamma, dbeta, None, None, None, None, None


class GroupRowsTrain(torch.autograd.Function):
    """Fused ball query + group -> fp32 rows [features | xyz - centre (- shift)]; gradient to the features only (the reference stops
    the gradient at shift_pred, models/model_rpointnet.py:377, and xyz / new_xyz are data)."""

    @staticmethod
    def forward(ctx, xyz, new_xyz, points, radius, nsample, shift=None):
        idx, _, grouped, ld = ops.ballquery_group(radius, nsample, xyz, new_xyz, points, torch.float32, shift=shift)
        ctx.save_for_backward(idx)
        ctx.shape = None if points is None else tuple(points.shape)
        ctx.ld = ld
        ctx.mark_non_differentiable(idx)
        return grouped, idx

    @staticmethod
    def backward(ctx, dgrouped, _didx):
        (idx,) = ctx.saved_tensors
        if ctx.shape is None:
            return None, None, None, None, None, None
        b, n, c = ctx.shape
        _, m, k = idx.shape
        dgrouped = dgrouped.contiguous()
        dp = torch.empty((b, n, c), dtype=torch.float32, device=dgrouped.device)
        check(_lib.lib().gspn_group_rows_grad(b, n, c, m, k, ctx.ld, dgrouped.data_ptr(), idx.data_ptr(), dp.data_ptr(), _s()), "group_rows_grad")
        return None, None, dp, None, None, None


def run_mlp_train(x2d, layers, bn_decay, pool_last=1, first_weight=None):
    decay = 0.9 if bn_decay is None else float(bn_decay)  # utils/tf_util.py:528
    for i, layer in enumerate(layers):
        w = first_weight if (i == 0 and first_weight is not None) else layer["weights"]
        pool = pool_last if i == len(layers) - 1 else 1
        x2d = MlpLayerTrain.apply(x2d, w, layer["biases"], layer.get("gamma"), layer.get("beta"), layer.get("moving_mean"),
                                  layer.get("moving_variance"), decay, pool, True)
    return x2d


def _max_rows(x2d, pool):
    """tf.reduce_max over the neighbourhood axis (pointnet_util.py:124) when no MLP layer is there to carry the pooling."""
    return x2d.reshape(-1, pool, x2d.shape[-1]).max(dim=1).values


def sa_module_train(xyz, points, npoint, radius, nsample, layers, bn_decay, use_xyz=True, layers2=(), group_all=False):
    """pointnet_sa_module(is_training=True), pooling='max': sample_and_group / sample_and_group_all -> mlp -> max -> mlp2
    (utils/pointnet_util.py:85-139)."""
    b, n, _ = xyz.shape
    if group_all:  # sample_and_group_all (:57-82): one group of all n points, rows [xyz | features], new_xyz = 0
        new_xyz = torch.zeros((b, 1, 3), dtype=torch.float32, device=xyz.device)
        idx = torch.arange(n, dtype=torch.int32, device=xyz.device).reshape(1, 1, n).repeat(b, 1, 1)
        rows = xyz if points is None else (torch.cat([xyz, points], dim=2) if use_xyz else points)
        x, npoint, nsample = rows.reshape(b * n, rows.shape[2]), 1, n
        w = None
    else:
        fps_idx = ops.farthest_point_sample(npoint, xyz)
        new_xyz = ops.gather_point(xyz, fps_idx)
        x, idx = GroupRowsTrain.apply(xyz, new_xyz.detach(), points, radius, nsample)
        w = layers[0]["weights"] if layers else None
        if points is not None and w is not None:  # grouped columns are [features | xyz]; the reference kernel rows are [xyz | features]
            w = torch.cat([w[3:], w[:3]], dim=0) if use_xyz else torch.cat([w, torch.zeros((3, w.shape[1]), device=w.device)], dim=0)
        if not layers and points is not None:     # no MLP: the pooled rows themselves, back in the reference's column order
            c = points.shape[2]
            x = torch.cat([x[:, c:c + 3], x[:, :c]], dim=1) if use_xyz else x[:, :c]
    x = run_mlp_train(x, layers, bn_decay, pool_last=nsample, first_weight=w) if layers else _max_rows(x, nsample)
    if layers2:
        x = run_mlp_train(x, list(layers2), bn_decay)
    return new_xyz, x.reshape(b, npoint, x.shape[-1]), idx


def encoding_net_train(xyz, new_xyz, points, radius_list, nsample_list, layer_lists, bn_decay, use_xyz, shift_pred):
    """multi_encoding_net(is_training=True) (models/model_rpointnet.py:28-77): per radius ball query + group (- new_xyz - shift_pred,
    points FIRST: this library's grouped column order) -> mlp -> max; concatenated over the radii."""
    b, m, _ = new_xyz.shape
    outs = []
    for radius, nsample, layers in zip(radius_list, nsample_list, layer_lists):
        x, _ = GroupRowsTrain.apply(xyz, new_xyz.detach(), points, radius, nsample, None if shift_pred is None else shift_pred.detach())
        w = layers[0]["weights"]
        if points is not None and not use_xyz:  # the xyz columns are there but the reference's rows do not have them
            w = torch.cat([w, torch.zeros((3, w.shape[1]), device=w.device)], dim=0)
        x = run_mlp_train(x, layers, bn_decay, pool_last=nsample, first_weight=w)
        outs.append(x.reshape(b, m, x.shape[-1]))
    return torch.cat(outs, dim=-1)


def fp_module_train(xyz1, xyz2, points1, points2, layers, bn_decay):
    """pointnet_fp_module(is_training=True)."""
    b, n, _ = xyz1.shape
    _, idx, weight = ops.three_nn(xyz1, xyz2, return_weight=True)
    interpolated = ops.three_interpolate(points2, idx, weight)
    x = torch.cat([interpolated, points1], dim=2) if points1 is not None else interpolated
    if not layers:
        return x
    y = run_mlp_train(x.reshape(b * n, x.shape[2]), layers, bn_decay)
    return y.reshape(b, n, y.shape[-1])


def trainable(store):
    """Mark every weight / bias / gamma / beta of a VariableStore as requiring grad; returns the list of leaves."""
    leaves = []
    for layers in store.values():
        for layer in layers:
            for k in ("weights", "biases", "gamma", "beta"):
                if layer.get(k) is not None:
                    layer[k].requires_grad_(True)
                    leaves.append(layer[k])
    return leaves


def allreduce_gradients(params, group=None):
    """Data-parallel step: ONE bucketed all-reduce (mean) of every parameter gradient per step.
    Works with any initialised torch.distributed backend (NCCL on the GPUs, gloo in the CPU test)."""
    import torch.distributed as dist
    if not params or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    # every rank contributes EVERY parameter (zeros where its shard produced no gradient), so the bucket has the same size everywhere
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    flat = torch.cat([p.grad.reshape(-1) for p in params])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= dist.get_world_size(group)
    o = 0
    for p in params:
        p.grad.copy_(flat[o:o + p.numel()].reshape(p.shape))
        o += p.numel()
