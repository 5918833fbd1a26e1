"""multi_encoding_net -- the GSPN multi-radius context encoder (models/model_rpointnet.py:28-77, called at :377 with
radius_list=[0.5,1.0,1.5], nsample_list=[256,256,512], mlp_list=[[64,128,256]]*3, mlp_list2=[], use_xyz=True,
fps_idx=ind_seed, shift_pred=stop_gradient(...)).  SURVEY.md 8f rank 2 / BASELINE config 3.

Per radius: query_ball_point -> group_point(xyz) - new_xyz - shift_pred -> group_point(points) ->
concat([grouped_points, grouped_xyz]) (points FIRST -- the opposite of sample_and_group, and exactly the column
order of this library's fused grouping kernel, so no weight permutation is needed) -> conv2d x len(mlp) ->
reduce_max; results concatenated over radii.  mlp_list2 (conv1d) and output_shift are model-head code outside
the SA/FP path and are not built (the reference call site passes mlp_list2=[] and output_shift=False).
"""
import torch

from . import mlp_tc, ops
from . import pointnet_util as pu


FUSED_MULTI_RADIUS = True  # one ordered scan for all radii (False: one query_ball_point per radius, as the reference)


def multi_encoding_net(xyz, points, npoint, radius_list, nsample_list, mlp_list, mlp_list2, is_training, bn_decay, scope, bn=True,
                       use_xyz=False, output_shift=False, shift_pred=None, fps_idx=None, variables=None, precision=None):
    """-> (new_xyz (b,npoint,3), new_points (b,npoint,sum mlp[-1]), shift_pred, fps_idx)."""
    pu._check_unbuilt(is_training)
    if mlp_list2 or output_shift:
        raise NotImplementedError("mlp_list2 / output_shift are conv1d head code outside the SA/FP path (reference call site uses neither)")
    store = pu.VARIABLES if variables is None else variables
    precision = precision or pu.DEFAULT_PRECISION
    b, n, _ = xyz.shape
    if fps_idx is None:
        fps_idx = ops.farthest_point_sample(npoint, xyz)
    new_xyz = ops.gather_point(xyz, fps_idx)
    m = new_xyz.shape[1]
    c = 0 if points is None else points.shape[2]
    cin = c + 3 if (use_xyz or points is None) else c
    outs = []
    all_layers = [store.layers(scope, "conv_prev_%d_" % i, cin, list(mlp), bn) for i, mlp in enumerate(mlp_list)]
    if is_training:  # fp32 training form: batch-statistics BN, autograd through grouping and the MLPs (gspn_b200/train.py)
        from . import train
        new_points = train.encoding_net_train(xyz, new_xyz, points, radius_list, nsample_list, all_layers, bn_decay,
                                              use_xyz or points is None, shift_pred)
        return new_xyz, new_points, shift_pred, fps_idx
    # the nested balls share their seeds: when every radius goes through the in-chain gather, ONE scan finds all index lists
    # (models/model_rpointnet.py:49-61 issues one query_ball_point per radius)
    fused_idx = None
    if (FUSED_MULTI_RADIUS and precision in pu.TC_PRECISIONS and mlp_tc.gather_ok(points) and 1 <= len(radius_list) <= 4
            and all(mlp_tc.tc_supported(l, k) for l, k in zip(all_layers, nsample_list))):
        fused_idx = [i_ for i_, _ in ops.query_ball_point_multi(radius_list, nsample_list, xyz, new_xyz)]
    for i, (radius, nsample, mlp) in enumerate(zip(radius_list, nsample_list, mlp_list)):
        layers = all_layers[i]
        rows = b * m * nsample
        tc = precision in pu.TC_PRECISIONS and mlp_tc.tc_supported(layers, nsample)
        if tc and mlp_tc.gather_ok(points):
            idx = fused_idx[i] if fused_idx is not None else ops.query_ball_point(radius, nsample, xyz, new_xyz)[0]
            perm = (list(range(c)) + ([c, c + 1, c + 2] if (use_xyz or points is None) else [-1, -1, -1]) + [-1] * (64 - c - 3))
            x = mlp_tc.mlp_chain_gather(xyz, new_xyz, shift_pred, points, idx, layers, perm, nsample, precision)  # validates its tensors
        elif tc:
            idx, _, img, ld = ops.ballquery_group(radius, nsample, xyz, new_xyz, points, "image:" + precision, shift=shift_pred)
            perm = (list(range(c)) + ([c, c + 1, c + 2] if (use_xyz or points is None) else [-1, -1, -1]) + [-1] * (ld - c - 3))
            x, _ = mlp_tc.mlp_chain(img, rows, ld, layers, perm, nsample, precision, k0_used=c + 3)
        else:
            idx, _, grouped, ld = ops.ballquery_group(radius, nsample, xyz, new_xyz, points, torch.float32, shift=shift_pred)
            first = layers[0]
            if points is not None and not use_xyz:
                first = dict(first)
                first["weights"] = torch.cat([first["weights"], torch.zeros((3, first["weights"].shape[1]), device=xyz.device)], dim=0)
            x = pu._run_mlp_f32(grouped, [first] + list(layers[1:]), pool_last=nsample)
        outs.append(x.reshape(b, m, x.shape[-1]))
    return new_xyz, torch.cat(outs, dim=-1), shift_pred, fps_idx
