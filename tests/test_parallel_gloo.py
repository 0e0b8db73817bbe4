"""CPU, world_size 2 over gloo: ray-shard bookkeeping + the single gradient all-reduce."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from loopy_slam_b200.parallel import GradAllReducer, shard_bounds, shard_rays


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        R = 101
        rays = torch.randn(R, 3)
        w = torch.nn.Parameter(torch.randn(3, 4))
        feats = torch.nn.Parameter(torch.randn(10, 4))
        unused = torch.nn.Parameter(torch.zeros(5))
        (mine,) = shard_rays([rays], rank, world)
        loss = ((mine @ w).sum(0) * feats).sum()          # a *sum* over rays, like the reference losses
        loss.backward()
        red = GradAllReducer([w, feats, unused])
        nbytes = red.allreduce_()
        # reference: the unsharded gradient
        w2 = w.detach().clone().requires_grad_(True)
        f2 = feats.detach().clone().requires_grad_(True)
        ((rays @ w2).sum(0) * f2).sum().backward()
        ok = torch.allclose(w.grad, w2.grad, rtol=1e-5, atol=1e-5) and torch.allclose(feats.grad, f2.grad, rtol=1e-5, atol=1e-5)
        ok = ok and unused.grad is not None and float(unused.grad.abs().sum()) == 0 and nbytes == (12 + 40 + 5) * 4
        # fast path: gradients that are views of ONE flat buffer (what the fused backward returns) travel in a
        # single all-reduce of that buffer; a parameter outside it still takes the generic path
        flat = torch.zeros(16 + 40)
        a = torch.nn.Parameter(torch.randn(3, 4))
        b = torch.nn.Parameter(torch.randn(10, 4))
        c = torch.nn.Parameter(torch.randn(7))
        a.grad = flat[:12].view(3, 4)
        b.grad = flat[16:56].view(10, 4)
        a.grad.fill_(rank + 1.0)
        b.grad.fill_(10.0 * (rank + 1))
        c.grad = torch.full((7,), float(rank + 1))
        nb = GradAllReducer([a, b, c]).allreduce_(flat)
        tot = world * (world + 1) / 2
        ok = ok and bool((a.grad == tot).all()) and bool((b.grad == 10 * tot).all()) and bool((c.grad == tot).all())
        ok = ok and a.grad.data_ptr() == flat.data_ptr() and nb == (56 + 7) * 4
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_exactly():
    for n in (0, 1, 7, 4992, 5000):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_sharded_sum_plus_allreduce_equals_full_gradient():
    world = 2
    port = _free_port()
    mgr = mp.get_context('spawn').Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0] and out[1]
