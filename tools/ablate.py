"""Throughput ablations of the pipelined executor: what does a stage cost the STEP (not its own kernel time)?
Replaces one op family by a cached result before the graphs are captured and re-measures ms/step; also sweeps the
number of graph lanes.  Run under gpurun.   usage: python tools/ablate.py [depth [precision]]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gspn_b200 import backbone, mlp_tc, ops, scenes
from gspn_b200.engine import BackboneEngine

dev = torch.device("cuda:0")
B, N, STEPS = 8, 32768, 32
DEPTH = int(sys.argv[1]) if len(sys.argv) > 1 else 8
PREC = sys.argv[2] if len(sys.argv) > 2 else None  # None = the default precision (bf16x3)
batches = []
for i in range(6):
    xyz, col = scenes.scannet_like_batch(i * B, B, N)
    batches.append((torch.from_numpy(xyz).to(dev), torch.from_numpy(col).to(dev)))
store, _ = backbone.random_variables(dev)


def measure(tag, depth=DEPTH):
    eng = BackboneEngine(store, B, N, precision=PREC, depth=depth, device=dev, warm_inputs=batches[0])
    for w in range(depth):
        eng.submit(*batches[w % 6])
    eng.synchronize()
    cur = torch.cuda.current_stream()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(cur)
    for s in range(STEPS):
        eng.submit(*batches[s % 6], after=a if s < depth else None)
    eng.join(cur)
    b.record(cur)
    torch.cuda.synchronize()
    print("%-40s depth %2d  %.4f ms/step" % (tag, depth, a.elapsed_time(b) / STEPS), flush=True)
    del eng


def cached(fn, keyfn):
    cache = {}

    def wrapper(*args, **kw):
        key = keyfn(*args, **kw)
        if key not in cache:
            cache[key] = fn(*args, **kw)
        return cache[key]
    return wrapper


def shp(t):
    return None if t is None else tuple(t.shape)


for d in (1, 2, 4, 8, 12, 16):
    measure("baseline", d)

real = {"fps": ops.farthest_point_sample, "chain": mlp_tc.mlp_chain, "chain_g": mlp_tc.mlp_chain_gather, "three_nn": ops.three_nn,
        "qbp": ops.query_ball_point, "bqg": ops.ballquery_group, "fpi": mlp_tc.fp_interp_mlp}

ops.farthest_point_sample = cached(real["fps"], lambda npoint, inp: (npoint, shp(inp)) if inp.shape[1] == N else (npoint, shp(inp), id(inp)))
measure("without FPS level 1")
ops.farthest_point_sample = cached(real["fps"], lambda npoint, inp: (npoint, shp(inp)))
measure("without any FPS")
mlp_tc.mlp_chain = cached(real["chain"], lambda a_img, rows, k0, layers, first_perm, pool, precision="bf16x3", k0_used=0, want_half=None, want_f32=True, relus=None: (rows, k0, pool, precision, want_half, want_f32))
mlp_tc.mlp_chain_gather = cached(real["chain_g"], lambda xyz, new_xyz, shift, points, idx, layers, first_perm, pool, precision="bf16x3": (shp(idx), pool, precision))
measure("without FPS and MLP chains")
ops.farthest_point_sample = real["fps"]
measure("without MLP chains")
mlp_tc.mlp_chain, mlp_tc.mlp_chain_gather = real["chain"], real["chain_g"]
ops.three_nn = cached(real["three_nn"], lambda xyz1, xyz2, **kw: (shp(xyz1), shp(xyz2), tuple(sorted(kw.items()))))
measure("without three_nn")
ops.three_nn = real["three_nn"]
ops.query_ball_point = cached(real["qbp"], lambda radius, nsample, xyz1, xyz2: (radius, nsample, shp(xyz1), shp(xyz2)))
ops.ballquery_group = cached(real["bqg"], lambda radius, nsample, xyz, new_xyz, points, dtype, *a, **k: (radius, nsample, shp(xyz), shp(new_xyz), shp(points)))
measure("without ball query / grouping")
ops.query_ball_point, ops.ballquery_group = real["qbp"], real["bqg"]
