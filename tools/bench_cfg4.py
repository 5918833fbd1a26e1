"""BASELINE config 4: data-parallel TRAIN step in the shape of R-PointNet's (train.py:121-219,242-297): per GPU 2 scenes of 18000
points (models/config.py:14,17) through the SA x4 + FP x4 backbone in its training form (batch-statistics batch norm over the WHOLE
batch across ranks = SyncBN, autograd through every custom op), a Chamfer term (nn_distance) on 256 proposal point sets of 512 points per
scene ((B*256, 512, 3) pairs, models/model_rpointnet.py:1346-1353), ONE bucketed NCCL all-reduce of the gradients, Adam.

  python bench.py --workload cfg4 [--steps K --warmup W]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --workload cfg4 --gpus N

The model heads (proposal generation, NMS, losses) are outside the SA/FP hot path: the proposal sets are gathered scene
points and the "prediction" is those points plus a learnable offset, so that the Chamfer gradient (NnDistanceGrad) reaches a parameter.
The MLP arithmetic of the training form is fp32 on CUDA cores (gspn_b200/train.py).  Prints one JSON line (rank 0): scenes/s of the whole
job (weak scaling: per-GPU work fixed), CUDA-event timed, max over ranks; the parameter checksum proves every rank took the same step."""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gspn_b200
from gspn_b200 import backbone, scenes, train

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--warmup", type=int, default=3)
ap.add_argument("--precision", default=None)
ap.add_argument("--gpus", type=int, default=1)
ap.add_argument("--nccl-moments", action="store_true", help="A/B: batch-norm moment all-reduces through NCCL instead of NVLink peer memory")
ap.add_argument("--eager-launch", action="store_true", help="A/B: launch every kernel of the step from Python instead of replaying one CUDA graph")
args = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
peer = None
if world > 1 and not args.nccl_moments:
    from gspn_b200 import p2p
    peer = p2p.PeerGroup(dev)
    train.use_peer_moments(peer)
B, N, NSMP, NPTS = 2, 18000, 256, 512
store, _ = backbone.random_variables(dev)  # same seed on every rank -> identical initial parameters
offset = torch.zeros(3, device=dev, requires_grad=True)
params = train.trainable(store) + [offset]
opt = torch.optim.Adam(params, lr=1e-3, capturable=not args.eager_launch)  # capturable: the step count lives on the device
g = torch.Generator(device="cpu").manual_seed(1234 + rank)
total = args.warmup + args.steps
batches = []
for step in range(total):
    xyz, col = scenes.scannet_like_batch((rank * total + step) * B, B, N)
    sel = torch.stack([torch.randint(0, N, (NSMP * NPTS,), generator=g) for _ in range(B)]).to(torch.int32)
    batches.append((torch.from_numpy(xyz).to(dev), torch.from_numpy(col).to(dev), sel.to(dev)))


def step_fn(i):
    return step_on(*batches[i])


def step_on(x, c, sel):
    out = backbone.forward(x, c, store, is_training=True, bn_decay=0.9)
    gt = gspn_b200.gather_point(x, sel).reshape(B * NSMP, NPTS, 3)           # 256 proposal sets of 512 points per scene
    pred = gt.flip(1) * 0.98 + offset                                          # stand-in for the generated shapes
    d1, _, d2, _ = gspn_b200.nn_distance(pred, gt)                             # Chamfer, (B*256, 512, 3) pairs
    loss = out["l0_points"].square().mean() + d1.mean() + d2.mean() + out["points"][4].mean()
    opt.zero_grad(set_to_none=True)
    loss.backward()
    train.allreduce_gradients(params)
    opt.step()
    return loss


graph = None
if not args.eager_launch:
    # The whole step -- forward, backward, both kinds of all-reduce, Adam -- is a few thousand small launches: captured ONCE into a CUDA
    # graph and replayed on static input buffers (the step has no host synchronisation: equal shards, device-side epochs in the peer
    # all-reduce, capturable Adam).  Every rank captures and replays in lockstep.
    static = [torch.empty_like(t) for t in batches[0]]
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(min(3, args.warmup)):  # eager steps first: optimizer state, cached weight permutations, allocator pools
            loss = step_fn(i)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    graph = torch.cuda.CUDAGraph()
    try:
        with torch.cuda.graph(graph):
            static_loss = step_on(*static)
    except Exception as e:  # a capture that fails falls back to eager launches, and says so in the JSON line
        print("bench_cfg4: CUDA-graph capture failed (%s); launching eagerly" % (e,), file=sys.stderr)
        graph = None
        torch.cuda.synchronize()
    if graph is not None:
        eager_step = step_fn

        def step_fn(i):  # noqa: F811 -- from here on a step is a copy into the static buffers and one graph launch
            for dst, src in zip(static, batches[i]):
                dst.copy_(src, non_blocking=True)
            graph.replay()
            return static_loss

for i in range(args.warmup):
    loss = step_fn(i)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for i in range(args.warmup, total):
    loss = step_fn(i)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b) / args.steps
chk = torch.stack([p.detach().double().sum() for p in params]).sum().reshape(1)
same = True
if world > 1:
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    allc = [torch.zeros_like(chk) for _ in range(world)]
    dist.all_gather(allc, chk)
    same = all(bool(x == allc[0]) for x in allc)
if rank == 0:
    print(json.dumps({
        "metric": "R-PointNet-shaped train step scenes/sec (SA x4 + FP x4 fwd+bwd, SyncBN, Chamfer on 256 x 512-pt proposals/scene, Adam)",
        "value": world * B / (ms * 1e-3), "unit": "scenes/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (training form: CUDA-core GEMMs)", "data": "synthetic",
        "config": {"workload": "config4: data-parallel train step, 2 scenes x 18000 pts per GPU, 256 proposals/scene, NCCL all-reduce of "
                               "gradients + whole-batch batch-norm statistics", "sync_bn": bool(train.SYNC_BN), "launch": "eager" if graph is None else "one CUDA graph per step (forward + backward + all-reduces + Adam)", "moment_allreduce": "NVLink peer memory (csrc/p2p.cu)" if peer is not None else ("NCCL" if world > 1 else "none"), "params": int(sum(p.numel() for p in params))},
        "loss": float(loss.detach()), "parameters_identical_across_ranks": same}))
if peer is not None:
    peer.close()
if world > 1:
    dist.destroy_process_group()
