"""One eager forward of the config-2 backbone between cudaProfilerStart/Stop (for ncu --profile-from-start off)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gspn_b200 import backbone, scenes
dev = torch.device("cuda:0")
prec = sys.argv[1] if len(sys.argv) > 1 else None  # None = the default precision (bf16x3)
xyz, col = scenes.scannet_like_batch(0, 8, 32768)
x, c = torch.from_numpy(xyz).to(dev), torch.from_numpy(col).to(dev)
store, _ = backbone.random_variables(dev)
for _ in range(2):
    backbone.forward(x, c, store, precision=prec)
torch.cuda.synchronize()
torch.cuda.profiler.start()
backbone.forward(x, c, store, precision=prec)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
