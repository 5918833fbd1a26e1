"""Whole-batch batch norm across data-parallel ranks (utils/tf_util.py:530-534 normalises over the WHOLE batch).

CPU (world_size-2 gloo): the [sum, sum of squares, rows] all-reduce gives the whole-batch moments, and the gradient bucket has the
same size on every rank even when a rank has no gradient for a parameter.
GPU (two processes on cuda:0, gloo carrying CUDA tensors): the training-form MLP on a batch sharded over 2 ranks == the whole batch
on one rank -- activations, moving averages and (after allreduce_gradients) every parameter gradient."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _spawn(fn, world, *args, timeout=300):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=fn, args=(r, world, port, q) + args) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=timeout)
    for p in procs:
        p.join(timeout=60)
        if p.is_alive():
            p.terminate()
        assert p.exitcode == 0 or (isinstance(out, dict) and "error" in out)
    return out


def _moments_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gspn_b200 import train
    train.EQUAL_SHARDS = False  # unequal shards below: the row count has to travel
    g = torch.Generator().manual_seed(5)
    z = torch.randn(1000, 7, generator=g, dtype=torch.float64) * 3 + 1
    rows = [400, 600]  # unequal shards: the row count travels with the sums
    lo = sum(rows[:rank])
    mine = z[lo:lo + rows[rank]]
    s1, s2, n = train.allreduce_moments(mine.sum(0), (mine * mine).sum(0), rows[rank])
    # a parameter without gradient on rank 1 only
    p0, p1 = torch.nn.Parameter(torch.ones(3)), torch.nn.Parameter(torch.ones(2))
    p0.grad = torch.full((3,), float(rank + 1))
    if rank == 0:
        p1.grad = torch.full((2,), 4.0)
    train.allreduce_gradients([p0, p1])
    if rank == 0:
        q.put((s1.numpy(), s2.numpy(), n, z.numpy(), p0.grad.numpy(), p1.grad.numpy()))
    dist.destroy_process_group()


def test_moment_allreduce_and_gradient_bucket_gloo_cpu():
    s1, s2, n, z, g0, g1 = _spawn(_moments_worker, 2)
    assert n == 1000
    np.testing.assert_allclose(s1 / n, z.mean(0), rtol=1e-12)
    np.testing.assert_allclose(s2 / n - (s1 / n) ** 2, z.var(0), rtol=1e-10)
    np.testing.assert_allclose(g0, np.full(3, 1.5))  # mean of 1 and 2
    np.testing.assert_allclose(g1, np.full(2, 2.0))  # mean of 4 and the zeros rank 1 contributed


def _layers(rng, cin, widths, dev):
    out = []
    c = cin
    for co in widths:
        out.append({"weights": torch.tensor(rng.randn(c, co).astype(np.float32) * 0.2, device=dev, requires_grad=True),
                    "biases": torch.tensor(rng.randn(co).astype(np.float32) * 0.1, device=dev, requires_grad=True),
                    "gamma": torch.tensor((0.75 + 0.5 * rng.rand(co)).astype(np.float32), device=dev, requires_grad=True),
                    "beta": torch.tensor(rng.randn(co).astype(np.float32) * 0.1, device=dev, requires_grad=True),
                    "moving_mean": torch.zeros(co, device=dev), "moving_variance": torch.ones(co, device=dev)})
        c = co
    return out


def _syncbn_worker(rank, world, port, q):
    try:
        _syncbn_body(rank, world, port, q)
    except Exception as e:  # fail fast instead of leaving the parent on its queue
        import traceback
        q.put({"error": "rank %d: %s" % (rank, traceback.format_exc())})
        raise


def _syncbn_body(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from gspn_b200 import train
    dev = torch.device("cuda:0")
    rng = np.random.RandomState(3)
    rows, cin, widths, pool = 2048, 19, [32, 64], 32
    x_all = torch.tensor(rng.randn(rows, cin).astype(np.float32), device=dev)
    tgt = torch.tensor(rng.randn(rows // pool, widths[-1]).astype(np.float32), device=dev)
    half, halfg = rows // world, rows // pool // world
    # sharded: this rank's half, whole-batch statistics through the all-reduces
    lay = _layers(np.random.RandomState(9), cin, widths, dev)
    x = x_all[rank * half:(rank + 1) * half].clone().requires_grad_(True)
    y = train.run_mlp_train(x, lay, 0.9, pool_last=pool)
    loss = ((y - tgt[rank * halfg:(rank + 1) * halfg]) ** 2).mean()
    loss.backward()
    params = [l[k] for l in lay for k in ("weights", "biases", "gamma", "beta")]
    train.allreduce_gradients(params)
    if rank == 0:
        # the whole batch on one rank, no collective
        train.SYNC_BN = False
        ref = _layers(np.random.RandomState(9), cin, widths, dev)
        xr = x_all.clone().requires_grad_(True)
        yr = train.run_mlp_train(xr, ref, 0.9, pool_last=pool)
        ((yr - tgt) ** 2).mean().backward()
        res = {"y": (y.detach().cpu().numpy(), yr[:halfg].detach().cpu().numpy()),
               "dx": (x.grad.cpu().numpy() / world, xr.grad[:half].cpu().numpy())}  # the sharded loss is a mean over half the groups
        for i, (a, b) in enumerate(zip(lay, ref)):
            for k in ("weights", "biases", "gamma", "beta"):
                res["%d/%s" % (i, k)] = (a[k].grad.cpu().numpy(), b[k].grad.cpu().numpy())
            for k in ("moving_mean", "moving_variance"):
                res["%d/%s" % (i, k)] = (a[k].cpu().numpy(), b[k].cpu().numpy())
        q.put(res)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.gpu
def test_sharded_batch_norm_equals_whole_batch(cuda):
    res = _spawn(_syncbn_worker, 2, timeout=120)
    assert "error" not in res, res["error"]
    for name, (got, exp) in res.items():
        np.testing.assert_allclose(got, exp, rtol=2e-3, atol=2e-5, err_msg=name)
