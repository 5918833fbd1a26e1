"""What would the feature-propagation stages gain if the level-0 points came in a spatially coherent order?  Same scenes, points in
the generator's random order vs sorted by a 0.2 m grid cell (raster order): eager stage times of backbone.forward.  Run under gpurun."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import bench
from gspn_b200 import backbone, scenes
dev = torch.device("cuda:0")
xyz, col = scenes.scannet_like_batch(0, 8, 32768)
store, _ = backbone.random_variables(dev)


def cell_sorted(xyz, col, h=0.2):
    xs, cs = [], []
    for x, c in zip(xyz, col):
        g = np.floor((x - x.min(0)) / h).astype(np.int64)
        key = (g[:, 2] * 4096 + g[:, 1]) * 4096 + g[:, 0]
        o = np.argsort(key, kind="stable")
        xs.append(x[o]); cs.append(c[o])
    return np.stack(xs), np.stack(cs)


for name, (x, c) in (("random order", (xyz, col)), ("cell order", cell_sorted(xyz, col))):
    xt, ct = torch.from_numpy(x).to(dev), torch.from_numpy(c).to(dev)
    timers = bench.StageTimers(torch)
    for _ in range(3):
        backbone.forward(xt, ct, store)
    timers.on = True
    for _ in range(6):
        backbone.forward(xt, ct, store, timers=timers)
    torch.cuda.synchronize()
    ms = timers.avg_ms()
    keys = ["layer1:ballquery_group", "layer1:mlp", "fa_layer4:three_nn", "fa_layer4:interpolate", "fa_layer4:mlp"]
    print("%-13s " % name + "  ".join("%s %.4f" % (k, ms[k]) for k in keys) + "  | total %.4f" % sum(ms.values()), flush=True)
