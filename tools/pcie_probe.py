"""Raw pinned-host copy bandwidth of the box (what bounds bench.py's e2e leg). Run under gpurun."""
import torch
dev = torch.device("cuda:0")
out_d = torch.empty(8 * 32768 * 128, dtype=torch.bfloat16, device=dev)
in_h = torch.empty(8 * 32768 * 6, dtype=torch.float32).pin_memory()
in_d = torch.empty_like(in_h, device=dev)
outs = [torch.empty_like(out_d, device="cpu").pin_memory() for _ in range(4)]
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def t(fn, n=20):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n):
        fn(i)
    s1.synchronize(); s2.synchronize()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def d2h(i=0):
    with torch.cuda.stream(s1):
        outs[i % 4].copy_(out_d, non_blocking=True)


def both(i=0):
    d2h(i)
    with torch.cuda.stream(s2):
        in_d.copy_(in_h, non_blocking=True)


nb = out_d.numel() * 2
ms = t(d2h); print("D2H %.1f MB: %.3f ms  %.1f GB/s" % (nb / 1e6, ms, nb / ms / 1e6))
ms = t(both); print("D2H + concurrent H2D %.1f MB: %.3f ms/step" % (in_h.numel() * 4 / 1e6, ms))
