"""BASELINE config 4 in miniature: data-parallel TRAIN step of the SA x4 + FP x4 backbone -- forward with batch-statistics
BN, a Chamfer (nn_distance) + feature loss, backward through every custom op, ONE bucketed NCCL all-reduce of the
gradients, Adam.  2 scenes of 18000 points per GPU (models/config.py:14,17).  Not a bench line: it shows that the train
path runs, that all ranks hold identical parameters after the step, and what a step costs.

  python tools/train_step.py                                   # 1 GPU
  torchrun --standalone --nproc-per-node 2 tools/train_step.py # data parallel
"""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gspn_b200
from gspn_b200 import backbone, scenes, train

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
B, N = 2, 18000
store, _ = backbone.random_variables(dev)  # same seed on every rank -> identical initial parameters
params = train.trainable(store)
opt = torch.optim.Adam(params, lr=1e-3)
steps = 4
for step in range(steps):
    xyz, col = scenes.scannet_like_batch((rank * steps + step) * B, B, N)
    x, c = torch.from_numpy(xyz).to(dev), torch.from_numpy(col).to(dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = backbone.forward(x, c, store, is_training=True, bn_decay=0.9)
    feats = out["l0_points"]
    # Chamfer between the two coarsest levels (model_rpointnet.py:1348-1352 uses nn_distance on 512-point sets) + feature term
    d1, _, d2, _ = gspn_b200.nn_distance(out["xyz"][2], out["xyz"][3])
    loss = feats.square().mean() + d1.mean() + d2.mean() + out["points"][4].mean()
    opt.zero_grad(set_to_none=True)
    loss.backward()
    train.allreduce_gradients(params)
    opt.step()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        print("step %d loss %.5f  %.1f ms  (%d x %d points per GPU, %d GPU)" % (step, float(loss), dt * 1e3, B, N, world), flush=True)
if world > 1:
    chk = torch.stack([p.detach().double().sum() for p in params]).sum()
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    if rank == 0:
        print("parameter checksums identical across ranks:", all(bool(a == allc[0]) for a in allc))
    dist.destroy_process_group()
