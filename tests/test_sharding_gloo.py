"""Multi-GPU host logic on CPU: world_size-2 gloo. Scenes shard over ranks with no data-path
collective; the only exchange is the max-over-ranks of the step time."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gspn_b200 import scenes


def test_shard_scenes_partitions_exactly():
    for total in (1, 7, 8, 16, 33):
        for ws in (1, 2, 4, 8):
            spans = [scenes.shard_scenes(total, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lo, hi = scenes.shard_scenes(total, rank, world)
    # every rank builds ITS scenes only; checksum proves disjoint + complete coverage
    owned = torch.zeros(total, dtype=torch.int64)
    for s in range(lo, hi):
        xyz, _ = scenes.scannet_like_scene(s, 256)
        owned[s] = int(abs(float(xyz.sum())) * 1000) + 1
    dist.all_reduce(owned)  # test-only collective (the data path has none)
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # bench.py's max-over-ranks timing
    if rank == 0:
        q.put((owned.tolist(), float(t)))
    dist.destroy_process_group()


def test_two_rank_gloo_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    total = 5
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    owned, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = []
    for s in range(total):
        xyz, _ = scenes.scannet_like_scene(s, 256)
        expect.append(int(abs(float(xyz.sum())) * 1000) + 1)
    assert owned == expect  # each scene built by exactly one rank
    assert tmax == 2.0


def _grad_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gspn_b200.train import allreduce_gradients
    params = [torch.zeros(3, 4), torch.zeros(5), torch.zeros(2, 2)]
    for i, p in enumerate(params):
        p.grad = torch.full_like(p, float((rank + 1) * (i + 1)))
    params[2].grad = None  # a parameter without a gradient contributes zeros: the bucket has the same size on every rank
    allreduce_gradients(params)
    if rank == 0:
        q.put([None if p.grad is None else p.grad.flatten()[0].item() for p in params])
    dist.destroy_process_group()


def test_gradient_allreduce_two_ranks_gloo():
    """Config 4's only collective: one bucketed mean all-reduce of the MLP/BN gradients."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == [1.5, 3.0, 0.0]  # mean over ranks of (rank+1)*(i+1); zeros where no rank had a gradient
