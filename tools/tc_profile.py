"""Per-phase cycle breakdown of the tcgen05 MLP chain (CTA 0 / thread 0). Run under gpurun."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from gspn_b200 import _lib, mlp_tc
import test_gpu_parity as tp
dev = torch.device("cuda:0")
L = _lib.lib()
PREC = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
for name, rows, cin, widths, pool in [("sa1", 8 * 2048 * 32, 6, [32, 32, 64], 32), ("sa2", 8 * 512 * 32, 67, [64, 64, 128], 32),
                                      ("sa4", 8 * 32 * 32, 259, [256, 256, 512], 32), ("fp4", 8 * 32768, 131, [128, 128, 128], 1),
                                      ("fp1", 8 * 128, 768, [256, 256], 1)]:
    rng = np.random.RandomState(1)
    layers = tp.rand_layers(rng, cin, widths)
    tl = [{k: tp.T(v, dev) for k, v in l.items()} for l in layers]
    ld = ((cin + 63) // 64) * 64
    tiles = (rows + 127) // 128
    img = torch.zeros(tiles * (ld // 64) * 16384 * (2 if PREC == "bf16x3" else 1), dtype=torch.uint8, device=dev)
    mlp_tc.mlp_chain(img, rows, ld, tl, None, pool, PREC, k0_used=cin)
    prof = torch.zeros(16, dtype=torch.int64, device=dev)
    L.gspn_mlp_chain_set_profile(prof.data_ptr())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); mlp_tc.mlp_chain(img, rows, ld, tl, None, pool, PREC, k0_used=cin); b.record()
    torch.cuda.synchronize()
    L.gspn_mlp_chain_set_profile(None)
    p = prof.cpu().numpy().astype(float)
    n = max(p[4], 1)
    print(PREC + " %-4s rows %7d kernel+launch %.3f ms | per layer-step cycles: issue %.0f mma_wait %.0f epilogue %.0f sync %.0f (steps %d) | MMA thread: wait %.0f issue %.0f drain %.0f" %
          (name, rows, a.elapsed_time(b), p[0] / n, p[1] / n, p[2] / n, p[3] / n, int(p[4]), p[7] / n, p[5] / n, p[6] / n), flush=True)

# commuted feature-propagation form (gspn_mlp_chain_fp): config 2's fa_layer4
b, n, m, c1, c2, widths = 8, 32768, 2048, 3, 128, [128, 128, 128]
rng = np.random.RandomState(1)
tl = [{k: tp.T(v, dev) for k, v in l.items()} for l in tp.rand_layers(rng, c1 + c2, widths)]
p1 = torch.randn(b, n, c1, device=dev); p2 = torch.randn(b, m, c2, device=dev)
idx = torch.randint(0, m, (b, n, 3), device=dev, dtype=torch.int32)
w = torch.rand(b, n, 3, device=dev); w = w / w.sum(-1, keepdim=True)
for half, f32 in ((None, True), (torch.float16, False)):
    mlp_tc.fp_interp_mlp(p1, p2, idx, w, tl, None, "x", None, PREC, want_half=half, want_f32=f32)
    prof = torch.zeros(16, dtype=torch.int64, device=dev)
    L.gspn_mlp_chain_set_profile(prof.data_ptr())
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); mlp_tc.fp_interp_mlp(p1, p2, idx, w, tl, None, "x", None, PREC, want_half=half, want_f32=f32); e.record()
    torch.cuda.synchronize()
    L.gspn_mlp_chain_set_profile(None)
    p = prof.cpu().numpy().astype(float)
    nn = max(p[4], 1)
    print(PREC + " fp4c (%s out) rows %7d kernels+launch %.3f ms | per step cycles: mma_wait %.0f epilogue %.0f sync %.0f (steps %d) | MMA thread: wait %.0f issue %.0f drain %.0f" %
          ("f32" if f32 else "f16", b * n, a.elapsed_time(e), p[1] / nn, p[2] / nn, p[3] / nn, int(p[4]), p[7] / nn, p[5] / nn, p[6] / nn), flush=True)
