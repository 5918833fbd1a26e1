"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck / initcheck / synccheck).
   compute-sanitizer --tool memcheck python tools/sanitize.py"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gspn_b200
from gspn_b200 import backbone, context_encoder, scenes
dev = torch.device("cuda:0")
if os.environ.get("SANITIZE_TMA_OUT") == "0":
    # initcheck does not see cp.async.bulk.tensor stores as writes: with the chains' TMA output path on, every consumer of a chain's
    # output is reported as reading uninitialised memory.  Row-per-lane stores instead for that tool.
    gspn_b200._lib.lib().gspn_mlp_chain_tune(2, 2, 0)
ONLY = os.environ.get("SANITIZE_ONLY")
if ONLY == "fps_bucket":
    gspn_b200._lib.lib().gspn_fps_tune(1)
    gspn_b200.farthest_point_sample(40, torch.rand(1, 9000, 3, device=dev))
    torch.cuda.synchronize()
    print("sanitize workload done")
    sys.exit(0)
xyz, col = scenes.scannet_like_batch(0, 2, 4608)
x, c = torch.from_numpy(xyz).to(dev), torch.from_numpy(col).to(dev)
specs = backbone.scaled_sa_specs(4608)
store, _ = backbone.random_variables(dev, sa_specs=specs)
for prec in ("bf16x3", "bf16", "fp32"):
    out = backbone.forward(x, c, store, sa_specs=specs, precision=prec, l0_half=torch.float16)
torch.cuda.synchronize()
fps = gspn_b200.farthest_point_sample(16, x)
context_encoder.multi_encoding_net(x, c, 16, [0.5, 1.0], [64, 128], [[64, 128, 256]] * 2, [], False, None, "ctx", use_xyz=True, fps_idx=fps,
                                   variables=store)
a = torch.randn(2, 3000, 3, device=dev, requires_grad=True)
b = torch.randn(2, 2500, 3, device=dev, requires_grad=True)
d1, i1, d2, i2 = gspn_b200.nn_distance(a, b)
(d1.sum() + d2.sum()).backward()
p = torch.randn(2, 300, 16, device=dev, requires_grad=True)
idx, _ = gspn_b200.query_ball_point(0.5, 8, x[:, :300].contiguous(), x[:, :40].contiguous())
gspn_b200.group_point(p, idx).sum().backward()
dist, i3 = gspn_b200.three_nn(x[:, :500].contiguous(), x[:, :300].contiguous())
w = torch.full((2, 500, 3), 1 / 3, device=dev)
gspn_b200.three_interpolate(p, i3, w).sum().backward()
gspn_b200.gather_point(p, fps % 300).sum().backward()
gspn_b200.nearest_point(x, x[:, :200].contiguous())
gspn_b200.box_shrink(torch.rand(2, 16, 6, device=dev), x)
big = torch.rand(1, 140000, 3, device=dev)
gspn_b200.farthest_point_sample(8, big)
# round-2 kernels: one scan for nested balls, order-independent backward, training form, bucket-pruned FPS (opt-in), single-rank mailbox
from gspn_b200 import ops, train, _lib
gspn_b200.ops.query_ball_point_multi([0.3, 0.6, 1.2], [8, 16, 32], x[:, :2000].contiguous(), x[:, :64].contiguous())
ops.DETERMINISTIC_BACKWARD = True
gspn_b200.group_point(p, idx).sum().backward()
gspn_b200.three_interpolate(p, i3, w).sum().backward()
gspn_b200.gather_point(p, fps % 300).sum().backward()
ops.DETERMINISTIC_BACKWARD = False
leaves = train.trainable(store)
out = backbone.forward(x, c, store, sa_specs=specs, is_training=True, bn_decay=0.9)
out["l0_points"].square().mean().backward()
# synccheck reports "Barrier error detected. Missing init ... shared address 0x0" at a PC outside fps_bucket_kernel, a kernel whose SASS
# holds no mbarrier instruction at all (BAR.SYNC, the tensor-memory allocator's UTCATOMSWS, LDTM/STTM only): SANITIZE_SKIP=fps_bucket
# leaves it out of that tool's pass; memcheck / racecheck / initcheck run it.  SANITIZE_ONLY=fps_bucket reproduces the report.
for mode in ((2,) if os.environ.get("SANITIZE_SKIP") == "fps_bucket" else (1, 2)):
    _lib.lib().gspn_fps_tune(mode)
    big2 = torch.rand(2, 9000, 3, device=dev)
    gspn_b200.farthest_point_sample(300, big2)
_lib.lib().gspn_fps_tune(0)
# doors: work-stealing tile scheduling of the chains, two clouds per FPS CTA
_lib.lib().gspn_mlp_chain_tune_sched(1)
backbone.forward(x, c, store, sa_specs=specs, l0_half=torch.float16)
_lib.lib().gspn_mlp_chain_tune_sched(0)
_lib.lib().gspn_fps_tune_pack(2)
gspn_b200.farthest_point_sample(64, torch.rand(3, 20000, 3, device=dev))
_lib.lib().gspn_fps_tune_pack(1)
torch.cuda.synchronize()
print("sanitize workload done")
