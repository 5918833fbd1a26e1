"""BASELINE config 3: GSPN shape-proposal context encoder (multi_encoding_net, models/model_rpointnet.py:377): 128 seeds per
scene, radii 0.5/1.0/1.5, nsample 256/256/512, mlp [64,128,256] per radius, batch 16 scenes of 18000 points (config.py:14).
Reports ms, seeds/s, scene/s and the tensor-pipe rate of the grouped MLP. Run under gpurun."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gspn_b200
from gspn_b200 import context_encoder, scenes
from gspn_b200 import pointnet_util as pu

dev = torch.device("cuda:0")
B, N, SEEDS = 16, 18000, 128
radii, ks, mlps = [0.5, 1.0, 1.5], [256, 256, 512], [[64, 128, 256]] * 3
xyz, col = scenes.scannet_like_batch(0, B, N)
x, c = torch.from_numpy(xyz).to(dev), torch.from_numpy(col).to(dev)
store = pu.VariableStore(device=dev, seed=7)
fps = gspn_b200.farthest_point_sample(SEEDS, x)
shift = torch.zeros((B, SEEDS, 3), device=dev)


def run(prec):
    return context_encoder.multi_encoding_net(x, c, SEEDS, radii, ks, mlps, [], False, None, "ctx", use_xyz=True, shift_pred=shift, fps_idx=fps,
                                              variables=store, precision=prec)[1]


res = {}
for prec in ("bf16", "fp32"):
    for _ in range(3):
        out = run(prec)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = run(prec); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = float(np.median(ts))
    flops = 2.0 * B * SEEDS * sum(ks) * (6 * 64 + 64 * 128 + 128 * 256)
    res[prec] = dict(ms=ms, scenes_per_s=B / ms * 1e3, seeds_per_s=B * SEEDS / ms * 1e3, mlp_gflop=flops / 1e9, tflops_if_all_mlp=flops / ms / 1e9)
    print(prec, res[prec], tuple(out.shape), flush=True)
a = run("bf16"); b = run("fp32")
res["bf16_vs_fp32_relerr"] = float((a - b).abs().max() / b.abs().max())
print("bf16 vs fp32 normwise", res["bf16_vs_fp32_relerr"])
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/cfg3_context_encoder.json", "w"), indent=1)
