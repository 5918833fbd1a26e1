"""Per-op timings on the config-2 shapes (CUDA events, median of N after warm-up). Run under gpurun.
usage: python tools/op_bench.py [ballquery] [three_nn] [mlp] [interp] [fps]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gspn_b200 import backbone, mlp_tc, ops, scenes  # noqa: E402
from gspn_b200 import pointnet_util as pu  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    what = set(sys.argv[1:]) or {"ballquery", "three_nn", "mlp", "interp", "fps"}
    B = 8
    xyz_np, col_np = scenes.scannet_like_batch(0, B, 32768)
    xyz, col = torch.from_numpy(xyz_np).to(dev), torch.from_numpy(col_np).to(dev)
    store, _ = backbone.random_variables(dev)
    out = backbone.forward(xyz, col, store, precision="bf16")
    xs, ps = out["xyz"], out["points"]
    res = {}
    if "fps" in what:
        for lvl, (m, r, k, mlp) in enumerate(backbone.SA_SPECS):
            res["fps_l%d" % (lvl + 1)] = timeit(lambda: ops.farthest_point_sample(m, xs[lvl]))
    if "ballquery" in what:
        for lvl, (m, r, k, mlp) in enumerate(backbone.SA_SPECS):
            for qpw in (0, 1, 2, 4):
                os.environ["GSPN_BQ_QPW"] = str(qpw)
                res["bq_only_l%d_qpw%d" % (lvl + 1, qpw)] = timeit(lambda: ops.query_ball_point(r, k, xs[lvl], xs[lvl + 1]))
                res["bq_group_bf16_l%d_qpw%d" % (lvl + 1, qpw)] = timeit(
                    lambda: ops.ballquery_group(r, k, xs[lvl], xs[lvl + 1], ps[lvl], torch.bfloat16))
            os.environ["GSPN_BQ_QPW"] = "0"
    if "three_nn" in what:
        for i in range(4):
            lvl = 3 - i
            res["three_nn_fp%d" % (i + 1)] = timeit(lambda: ops.three_nn(xs[lvl], xs[lvl + 1], return_weight=True))
    if "interp" in what:
        up = ps[4]
        for i, mlp in enumerate(backbone.FP_SPECS):
            lvl = 3 - i
            _, idx, w = ops.three_nn(xs[lvl], xs[lvl + 1], return_weight=True)
            res["interp_f32_fp%d" % (i + 1)] = timeit(lambda: ops.three_interpolate(up, idx, w))
            layers = store["fa_layer%d/conv_" % (i + 1)]
            res["fp_module_bf16_fp%d" % (i + 1)] = timeit(
                lambda: pu.pointnet_fp_module(xs[lvl], xs[lvl + 1], ps[lvl], up, mlp, False, None, "fa_layer%d" % (i + 1), variables=store,
                                              precision="bf16"))
            up = pu.pointnet_fp_module(xs[lvl], xs[lvl + 1], ps[lvl], up, mlp, False, None, "fa_layer%d" % (i + 1), variables=store,
                                       precision="bf16")
    if "mlp" in what:
        for lvl, (m, r, k, mlp) in enumerate(backbone.SA_SPECS):
            layers = store["layer%d/conv" % (lvl + 1)]
            idx, cnt, img, ld = ops.ballquery_group(r, k, xs[lvl], xs[lvl + 1], ps[lvl], torch.bfloat16)
            c = ps[lvl].shape[2]
            perm = [3 + t for t in range(c)] + [0, 1, 2] + [-1] * (ld - c - 3)
            rows = B * m * k
            t = timeit(lambda: mlp_tc.mlp_chain(img, rows, ld, layers, perm, k))
            dims = [c + 3] + mlp
            fl = 2 * rows * sum(a * b for a, b in zip(dims, dims[1:]))
            res["mlp_chain_sa%d" % (lvl + 1)] = t
            res["mlp_chain_sa%d_tflops" % (lvl + 1)] = fl / t / 1e9
    for k_, v in res.items():
        print("%-32s %.4f" % (k_, v), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/op_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
