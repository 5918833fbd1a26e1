"""The tcgen05 MLP chain alone on the config-2 shapes (for ncu --set full --import-source on, and for A/B timing). Run under gpurun.
usage: python tools/chain_only.py [bf16x3|bf16] [sa1 sa2 sa3 sa4 fp1 fp2 fp3 fp4 fp4c]     (fp4c = commuted-interpolation form of fp4)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from gspn_b200 import mlp_tc
import test_gpu_parity as tp
dev = torch.device("cuda:0")
CFG = {"sa1": (8 * 2048 * 32, 6, [32, 32, 64], 32), "sa2": (8 * 512 * 32, 67, [64, 64, 128], 32), "sa3": (8 * 128 * 32, 131, [128, 128, 256], 32),
       "sa4": (8 * 32 * 32, 259, [256, 256, 512], 32), "fp1": (8 * 128, 768, [256, 256], 1), "fp2": (8 * 512, 384, [256, 256], 1),
       "fp3": (8 * 2048, 320, [256, 128], 1), "fp4": (8 * 32768, 131, [128, 128, 128], 1)}
args = sys.argv[1:]
prec = args.pop(0) if args and args[0] in ("bf16", "bf16x3") else "bf16x3"
which = args or ["sa1", "fp4"]
reps = int(os.environ.get("REPS", "5"))


def timed(fn):
    ts = []
    for r in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


for name in which:
    rng = np.random.RandomState(1)
    if name == "fp4c":
        b, n, m, c1, c2, widths = 8, 32768, 2048, 3, 128, [128, 128, 128]
        tl = [{k: tp.T(v, dev) for k, v in l.items()} for l in tp.rand_layers(rng, c1 + c2, widths)]
        p1 = torch.randn(b, n, c1, device=dev); p2 = torch.randn(b, m, c2, device=dev)
        idx = torch.randint(0, m, (b, n, 3), device=dev, dtype=torch.int32)
        w = torch.rand(b, n, 3, device=dev); w = w / w.sum(-1, keepdim=True)
        t = timed(lambda: mlp_tc.fp_interp_mlp(p1, p2, idx, w, tl, None, "x", None, prec))
        rows, fl = b * n, 2 * b * n * sum(x * y for x, y in zip([c1 + c2] + widths, widths))
    else:
        rows, cin, widths, pool = CFG[name]
        tl = [{k: tp.T(v, dev) for k, v in l.items()} for l in tp.rand_layers(rng, cin, widths)]
        ld = ((cin + 63) // 64) * 64
        tiles = (rows + 127) // 128
        mul = 2 if prec == "bf16x3" else 1
        img = (torch.randn(tiles * (ld // 64) * 8192 * mul, device=dev) * 0.5).to(torch.bfloat16).view(torch.uint8)
        t = timed(lambda: mlp_tc.mlp_chain(img, rows, ld, tl, None, pool, prec, k0_used=cin))
        fl = 2 * rows * sum(x * y for x, y in zip([cin] + widths, widths))
    print("%-6s %-4s rows %7d  %.4f ms (min of %d, incl. python launch)  %.1f TFLOP/s algorithmic" % (prec, name, rows, t, reps, fl / t / 1e9), flush=True)
