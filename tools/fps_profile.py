"""Per-phase cycle breakdown of the FPS round (gspn_fps_profile). Run under gpurun."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gspn_b200 import _lib, scenes
dev = torch.device("cuda:0")
L = _lib.lib()
for (b, n, m, thr, ppt, cl) in [(8, 32768, 2048, 128, 32, 8), (8, 32768, 2048, 256, 16, 8), (8, 2048, 512, 512, 4, 1), (8, 2048, 512, 64, 4, 8),
                                (8, 2048, 512, 32, 4, 16), (1, 4096, 1024, 128, 4, 8)]:
    x = torch.from_numpy(scenes.scannet_like_batch(0, b, n)[0]).to(dev)
    out = torch.empty((b, m), dtype=torch.int32, device=dev)
    prof = torch.zeros(4, dtype=torch.int64, device=dev)
    for _ in range(2):
        rc = L.gspn_fps_profile(b, n, m, x.data_ptr(), out.data_ptr(), thr, ppt, cl, prof.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    p = prof.cpu().numpy() / float(m - 1)
    print("b%d n%d m%d thr%d ppt%d cl%d rc=%d cycles/round: compute %.0f warp_reduce %.0f exchange %.0f table %.0f total %.0f" %
          (b, n, m, thr, ppt, cl, rc, p[0], p[1], p[2], p[3], p.sum()), flush=True)
