"""Times every legal (threads, ppt, cluster) mapping of the FPS kernel on the config-2 clouds
(CUDA events, median of 5 after 2 warm-ups).  Run under gpurun; writes gpurun_out/fps_sweep.json."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gspn_b200 import _lib, scenes  # noqa: E402


def time_cfg(x, m, threads, ppt, cluster, reps=5):
    b, n, _ = x.shape
    out = torch.empty((b, m), dtype=torch.int32, device=x.device)
    L = _lib.lib()
    s = torch.cuda.current_stream().cuda_stream
    ts = []
    for i in range(reps + 2):
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rc = L.gspn_farthest_point_sample_cfg(b, n, m, x.data_ptr(), out.data_ptr(), threads, ppt, cluster, s)
        e.record()
        if rc != 0:
            return None, None
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(a.elapsed_time(e))
    return float(np.median(ts)), out


def main():
    dev = torch.device("cuda:0")
    res = []
    for (b, n, m) in [(8, 32768, 2048), (8, 2048, 512), (8, 512, 128), (8, 128, 32), (1, 4096, 1024), (16, 18000, 256)]:
        xyz = scenes.scannet_like_batch(0, b, n)[0]
        x = torch.from_numpy(xyz).to(dev)
        ref = None
        for cluster in (1, 2, 4, 8, 16):
            for ppt in (1, 2, 4, 8, 16, 32):
                for threads in (32, 64, 128, 256, 512, 1024):
                    cap = threads * ppt * cluster
                    if cap < n or cap >= 4 * n + 2048:
                        continue
                    t, out = time_cfg(x, m, threads, ppt, cluster)
                    if t is None:
                        continue
                    if ref is None:
                        ref = out.clone()
                    same = bool(torch.equal(ref, out))
                    res.append(dict(b=b, n=n, m=m, threads=threads, ppt=ppt, cluster=cluster, ms=t, us_per_round=t * 1e3 / (m - 1), same=same))
                    print(res[-1], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/fps_sweep.json", "w"), indent=1)
    for key in sorted(set((r["b"], r["n"], r["m"]) for r in res)):
        best = sorted((r for r in res if (r["b"], r["n"], r["m"]) == key), key=lambda r: r["ms"])[:4]
        print(key, [(r["threads"], r["ppt"], r["cluster"], round(r["ms"], 3)) for r in best])


if __name__ == "__main__":
    main()
