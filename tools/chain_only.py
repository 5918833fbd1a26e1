"""The tcgen05 MLP chain alone on the config-2 shapes (for ncu --set full --import-source on). Run under gpurun.
usage: python tools/chain_only.py [sa1 sa2 sa3 sa4 fp1 fp2 fp3 fp4]"""
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
which = sys.argv[1:] or ["sa1", "fp4"]
reps = int(os.environ.get("REPS", "3"))
for name in which:
    rows, cin, widths, pool = CFG[name]
    rng = np.random.RandomState(1)
    tl = [{k: tp.T(v, dev) for k, v in l.items()} for l in tp.rand_layers(rng, cin, widths)]
    ld = ((cin + 63) // 64) * 64
    tiles = (rows + 127) // 128
    img = (torch.randn(tiles * (ld // 64) * 8192, device=dev) * 0.5).to(torch.bfloat16).view(torch.uint8)
    ts = []
    for r in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); mlp_tc.mlp_chain(img, rows, ld, tl, None, pool, want_bf16=(pool == 1)); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    fl = 2 * rows * sum(x * y for x, y in zip([cin] + widths, widths))
    print("%-4s rows %7d  %.4f ms (min of %d)  %.1f TFLOP/s" % (name, rows, min(ts), reps, fl / min(ts) / 1e9), flush=True)
