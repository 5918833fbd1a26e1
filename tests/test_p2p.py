"""Peer-memory all-reduce over NVLink (csrc/p2p.cu, gspn_b200/p2p.py): two processes, one GPU each.  Needs 2 GPUs (gpurun --gpus 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    try:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        from gspn_b200 import p2p, train
        grp = p2p.PeerGroup(dev, max_doubles=1024)
        rng = np.random.RandomState(100 + rank)
        ok = True
        for it in range(200):  # many dependent calls: epoch parity, no slot is overwritten while its owner reads it
            n = [1, 7, 64, 513, 1024][it % 5]
            mine = rng.randn(n) * (10.0 ** (it % 7 - 3))
            buf = torch.tensor(mine, dtype=torch.float64, device=dev)
            grp.allreduce_(buf)
            both = [torch.zeros(n, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(both, torch.tensor(mine, dtype=torch.float64))
            exp = both[0].clone()
            for r in range(1, world):
                exp = exp + both[r]  # rank order, like the kernel
            ok = ok and torch.equal(buf.cpu(), exp)
        # inside a CUDA graph: the epoch is a device-side counter
        g = torch.cuda.CUDAGraph()
        static = torch.full((16,), float(rank + 1), dtype=torch.float64, device=dev)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            with torch.cuda.graph(g, stream=s):
                grp.allreduce_(static)
        for _ in range(3):
            static.fill_(float(rank + 1))
            g.replay()
            torch.cuda.synchronize()
            ok = ok and bool((static == float(sum(range(1, world + 1)))).all())
        # SyncBN through it == through torch.distributed
        train.use_peer_moments(grp)
        s1, s2, rows = train.allreduce_moments(torch.full((5,), 1.5 * (rank + 1), dtype=torch.float64, device=dev),
                                               torch.full((5,), 2.0, dtype=torch.float64, device=dev), 10)
        ok = ok and rows == 10 * world and bool((s1 == 1.5 * sum(range(1, world + 1))).all()) and bool((s2 == 2.0 * world).all())
        train.use_peer_moments(None)
        grp.close()
        if rank == 0:
            q.put("ok" if ok else "mismatch")
        dist.destroy_process_group()
    except Exception:
        import traceback
        q.put("rank %d: %s" % (rank, traceback.format_exc()))
        raise


def test_peer_memory_allreduce_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run under gpurun --gpus 2)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        if p.is_alive():
            p.terminate()
    assert res == "ok", res
