"""CPU, world_size 2 over gloo: the bucketed gradient exchange (GradArena) reproduces DDP's all-reduce(mean)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from multimodal_mvd_seg_b200.ddp import GradArena
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.randn(s)) for s in [(7, 3), (5,), (64, 9), (11,), (3, 3, 3)]]
        arena = GradArena(params, bucket_bytes=1024)   # several buckets
        assert len(arena.buckets) >= 2
        # arena is laid out in backward (reverse) order
        assert arena.offset[id(params[-1])][0] == 0
        g_local = {id(p): torch.full(p.shape, float(rank + 1)) * (i + 1) for i, p in enumerate(params)}
        arena.begin_step()
        # "backward": last parameter first; parameter 1 never receives a gradient
        for i in reversed(range(len(params))):
            if i == 1:
                continue
            v = arena.view_for(params[i])
            v.copy_(g_local[id(params[i])])
            arena.on_params_ready([params[i]])
        arena.finish()
        arena.attach_grads()
        ok = True
        for i, p in enumerate(params):
            want = torch.zeros(p.shape) if i == 1 else torch.full(p.shape, float(sum(range(1, world + 1)) * (i + 1)))
            ok = ok and torch.equal(p.grad, want)
        # second step reuses the arena: stale values of the skipped parameter must not leak
        arena.begin_step()
        for i in reversed(range(len(params))):
            arena.view_for(params[i]).copy_(g_local[id(params[i])])
            arena.on_params_ready([params[i]])
        arena.finish()
        ok = ok and torch.equal(arena.view_for(params[1]), torch.full((5,), float(sum(range(1, world + 1)) * 2)))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_grad_arena_allreduce_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(2))
    assert res == {0: True, 1: True}
