"""Smallest possible exercise of the tcgen05 chain (run under `timeout`): one layer, K=64, N=32."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from gspn_b200 import mlp_tc
import test_gpu_parity as tp
from oracle import oracle as O
dev = torch.device("cuda:0")
for rows, cin, widths, pool in [(128, 64, [32], 1), (256, 6, [32], 1), (512, 6, [32, 32, 64], 32), (384, 259, [256, 256, 512], 32)]:
    rng = np.random.RandomState(1)
    x = rng.randn(rows, cin).astype(np.float32)
    layers = tp.rand_layers(rng, cin, widths)
    tl = [{k: tp.T(v, dev) for k, v in l.items()} for l in layers]
    ld = ((cin + 63) // 64) * 64
    for prec in ("bf16", "bf16x3"):
        img = tp.encode_tile_image(x, ld, dev, split=(prec == "bf16x3"))
        out, _ = mlp_tc.mlp_chain(img, rows, ld, tl, None, pool, prec, k0_used=cin)
        torch.cuda.synchronize()
        exp = tp.emulate_chain(x, layers, pool) if prec == "bf16" else tp.oracle_chain(O, x, layers, pool)
        print(prec, rows, cin, widths, pool, "relerr", tp.relerr(out.cpu().numpy(), exp), flush=True)
