"""Per-phase cycle breakdown of the FPS round (gspn_fps_profile: full-scan cluster kernels), the two bucket-pruned kernels' cycles /
bucket updates per round (gspn_fps_pruned_profile: cluster form, gspn_fps_bucket_profile: single-CTA form), plus CUDA-event timings of
all three on config 2's SA1. Run under gpurun."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gspn_b200 import _lib, ops, scenes
dev = torch.device("cuda:0")
L = _lib.lib()
st = lambda: torch.cuda.current_stream().cuda_stream
for (b, n, m, thr, ppt, cl) in [(8, 32768, 2048, 128, 32, 8), (8, 2048, 512, 512, 4, 1)]:
    x = torch.from_numpy(scenes.scannet_like_batch(0, b, n)[0]).to(dev)
    out = torch.empty((b, m), dtype=torch.int32, device=dev)
    prof = torch.zeros(4, dtype=torch.int64, device=dev)
    for _ in range(2):
        rc = L.gspn_fps_profile(b, n, m, x.data_ptr(), out.data_ptr(), thr, ppt, cl, prof.data_ptr(), st())
    torch.cuda.synchronize()
    p = prof.cpu().numpy() / float(m - 1)
    print("full scan b%d n%d m%d thr%d ppt%d cl%d rc=%d cycles/round: compute %.0f warp_reduce %.0f exchange %.0f table %.0f total %.0f" %
          (b, n, m, thr, ppt, cl, rc, p[0], p[1], p[2], p[3], p.sum()), flush=True)
for (b, n, m) in [(8, 32768, 2048), (8, 16384, 1024), (1, 32768, 2048)]:
    x = torch.from_numpy(scenes.scannet_like_batch(0, b, n)[0]).to(dev)
    out = torch.empty((b, m), dtype=torch.int32, device=dev)
    L.gspn_fps_tune(2)
    wsb = L.gspn_farthest_point_sample_workspace_bytes(b, n, m)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    prof = torch.zeros(8, dtype=torch.int64, device=dev)
    for _ in range(2):
        prof.zero_()
        rc = L.gspn_fps_pruned_profile(b, n, m, x.data_ptr(), out.data_ptr(), ws.data_ptr(), wsb, prof.data_ptr(), st())
    torch.cuda.synchronize()
    p = prof.cpu().numpy().astype(float) / (m - 1)
    print("pruned cluster b%d n%d m%d rc=%d cycles/round (thread 0): test+update+candidate %.0f exchange %.0f table %.0f total %.0f | "
          "%.1f bucket updates/round (full scan: %d)" % (b, n, m, rc, p[0], p[2], p[3], p[0] + p[2] + p[3], p[4], (n + 31) // 32), flush=True)
    L.gspn_fps_tune(1)
    wsb = L.gspn_farthest_point_sample_workspace_bytes(b, n, m)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    prof = torch.zeros(12, dtype=torch.int64, device=dev)
    rc = L.gspn_fps_bucket_profile(b, n, m, x.data_ptr(), out.data_ptr(), ws.data_ptr(), wsb, prof.data_ptr(), st())
    torch.cuda.synchronize()
    p = prof.cpu().numpy().astype(float)
    print("bucket    b%d n%d m%d rc=%d: %.0f cycles/round, %.1f bucket updates/round (%.0f total; full scan would be %d)" %
          (b, n, m, rc, p[0] / max(p[1], 1), p[2] / max(p[1], 1), p[2], (m - 1) * ((n + 31) // 32)), flush=True)
    print("    warp 0 cycles/round: tests %.0f  updates %.0f  warp argmax %.0f  barrier wait %.0f  table reduce %.0f | full argmax on %.0f%% of the bucket updates" % (tuple(p[3:8] / max(p[1], 1)) + (100 * p[8] / max(p[2], 1),)), flush=True)
    for tune, name in ((2, "pruned cluster"), (1, "bucket single-CTA"), (0, "full-scan cluster")):
        L.gspn_fps_tune(tune)
        ts = []
        for rep in range(4):
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.farthest_point_sample(m, x); e.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(e))
        print("    %-18s %.4f ms (min of 4)" % (name, min(ts)), flush=True)
    L.gspn_fps_tune(0)
