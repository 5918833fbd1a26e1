"""Does splitting one large pinned D2H copy over several streams (copy engines) beat a single cudaMemcpyAsync? Run under gpurun."""
import torch
dev = torch.device("cuda:0")
nbytes = 8 * 32768 * 128 * 2
src = torch.empty(nbytes, dtype=torch.uint8, device=dev).random_()
dst = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
h2d_src = torch.empty(8 * 32768 * 6 * 4, dtype=torch.uint8, pin_memory=True)
h2d_dst = torch.empty_like(h2d_src, device=dev)
up = torch.cuda.Stream()
for parts in (1, 2, 3, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(parts)]
    step = (nbytes + parts - 1) // parts
    for with_h2d in (False, True):
        best = 1e9
        for rep in range(6):
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i, st in enumerate(streams):
                st.wait_event(a)
                with torch.cuda.stream(st):
                    dst[i * step:(i + 1) * step].copy_(src[i * step:(i + 1) * step], non_blocking=True)
            if with_h2d:
                up.wait_event(a)
                with torch.cuda.stream(up):
                    h2d_dst.copy_(h2d_src, non_blocking=True)
                torch.cuda.current_stream().wait_stream(up)
            for st in streams:
                torch.cuda.current_stream().wait_stream(st)
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        print("D2H %.1f MB in %d part(s)%s: %.3f ms = %.1f GB/s" % (nbytes / 1e6, parts, " + 6.3 MB H2D alongside" if with_h2d else "", best, nbytes / best / 1e6), flush=True)
assert torch.equal(dst, src.cpu())
