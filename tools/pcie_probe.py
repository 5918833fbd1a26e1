"""Raw pinned-host copy bandwidth of the box: what bounds bench.py's e2e leg.  Single process, or one rank per GPU under torchrun
(all ranks copy CONCURRENTLY; the aggregate host ingest is what the 8-GPU e2e number runs into).  Run under gpurun.

  python tools/pcie_probe.py
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port P tools/pcie_probe.py
Variants: plain pinned buffers first-touched after binding to the GPU's NUMA-local cores, and write-combined pinned memory
(cudaHostAllocWriteCombined) for the D2H target."""
import ctypes
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
try:
    import bench
    bench.bind_to_gpu_numa_node(local)
except Exception:
    pass
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
MB = 67.1  # the half-precision per-point map of one step (8 x 32768 x 128 x 2 bytes)
nbytes = 8 * 32768 * 128 * 2
src = torch.empty(nbytes, dtype=torch.uint8, device=dev)
s1 = torch.cuda.Stream()


def wc_pinned(n):
    """write-combined pinned host memory through the runtime (torch has no flag for it)"""
    rt = ctypes.CDLL("libcudart.so.12")
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), ctypes.c_uint(0x04))  # cudaHostAllocWriteCombined
    if rc != 0:
        return None, None
    return p, rt


def timed(copy, n=20):
    copy(0); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(s1)
    for i in range(n):
        copy(i)
    b.record(s1)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


outs = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(4)]


def d2h_plain(i):
    with torch.cuda.stream(s1):
        outs[i % 4].copy_(src, non_blocking=True)


res = {"plain": timed(d2h_plain)}
p, rt = wc_pinned(nbytes)
if p is not None:
    def d2h_wc(i):
        rt.cudaMemcpyAsync(p, ctypes.c_void_p(src.data_ptr()), ctypes.c_size_t(nbytes), ctypes.c_int(2), ctypes.c_void_p(s1.cuda_stream))
    res["write_combined"] = timed(d2h_wc)
line = "rank %d (gpu %d, cpus %s): " % (rank, local, sorted(os.sched_getaffinity(0))[:1] + ["..."] + sorted(os.sched_getaffinity(0))[-1:])
line += "  ".join("%s D2H %.1f MB %.3f ms = %.1f GB/s" % (k, MB, v, nbytes / v / 1e6) for k, v in res.items())
if world > 1:
    gathered = [None] * world
    dist.all_gather_object(gathered, (line, {k: nbytes / v / 1e6 for k, v in res.items()}))
    if rank == 0:
        for l, _ in gathered:
            print(l)
        for k in res:
            print("aggregate %s: %.1f GB/s over %d concurrent ranks" % (k, sum(g[1][k] for g in gathered), world))
    dist.destroy_process_group()
else:
    print(line)
