"""BASELINE config 5: nn_distance (chamfer) + three_interpolate sweep 1k-131k points, achieved GB/s vs the B200 HBM roofline
(and pair-evaluations/s for the O(n*m) scan, which is the binding resource for nn_distance).  Run under gpurun."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import gspn_b200  # noqa: E402

dev = torch.device("cuda:0")
PEAK = 6555.5
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass


def timeit(fn, reps=7, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


res = []
g = torch.Generator(device="cpu").manual_seed(0)
for n in [1024, 2048, 4096, 8192, 16384, 32768, 65536, 131072]:
    b = max(1, (1 << 20) // n)
    x1 = torch.randn((b, n, 3), generator=g).to(dev)
    x2 = torch.randn((b, n, 3), generator=g).to(dev)
    ms = timeit(lambda: gspn_b200.nn_distance(x1, x2))
    bytes_ = b * (12 * n + 12 * n + 8 * n + 8 * n)
    pairs = 2.0 * b * n * n
    res.append(dict(op="nn_distance", b=b, n=n, m=n, ms=ms, gbs=bytes_ / ms / 1e6, frac_hbm=bytes_ / ms / 1e6 / PEAK, gpairs_per_s=pairs / ms / 1e6))
    print(res[-1], flush=True)
    for c in (64, 128, 256):
        m = n // 16
        pts = torch.randn((b, m, c), generator=g).to(dev)
        idx = torch.randint(0, m, (b, n, 3), generator=g, dtype=torch.int32).to(dev)
        w = torch.rand((b, n, 3), generator=g).to(dev)
        ms = timeit(lambda: gspn_b200.three_interpolate(pts, idx, w))
        bytes_ = b * (24 * n + m * c * 4 + n * c * 4)
        res.append(dict(op="three_interpolate", b=b, n=n, m=m, c=c, ms=ms, gbs=bytes_ / ms / 1e6, frac_hbm=bytes_ / ms / 1e6 / PEAK))
        print(res[-1], flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(dict(peak_hbm_gbs=PEAK, rows=res), open("gpurun_out/cfg5_sweep.json", "w"), indent=1)
