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


def _epoch_worker(rank, world, port, q):
    """every rank aggregates its own step outputs; the logged values must equal the single-process result over the union
    (MVDTrainer.py:990-993, 1071-1088: all_gather_object of losses and tp / fp / fn)."""
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        import numpy as np
        from multimodal_mvd_seg_b200 import trainer as T

        def outputs(r):
            rng = np.random.default_rng(50 + r)
            tr = [{'loss': np.array(rng.normal(1.0, 0.1), dtype=np.float32)} for _ in range(3)]
            va = [{'loss': np.array(rng.normal(0.7, 0.1), dtype=np.float32),
                   'tp_hard': rng.integers(1, 100, 3).astype(np.float64), 'fp_hard': rng.integers(1, 50, 3).astype(np.float64),
                   'fn_hard': rng.integers(1, 50, 3).astype(np.float64)} for _ in range(3)]
            return tr, va

        class Stand:
            pass
        me = Stand()
        me.is_ddp, me.current_epoch, me.logger = True, 0, T.nnUNetLogger()
        tr, va = outputs(rank)
        T.nnUNetTrainer.on_train_epoch_end(me, tr)
        T.nnUNetTrainer.on_validation_epoch_end(me, va)
        ref = Stand()
        ref.is_ddp, ref.current_epoch, ref.logger = False, 0, T.nnUNetLogger()
        all_tr, all_va = [], []
        for r in range(world):
            a, b = outputs(r)
            all_tr += a
            all_va += b
        T.nnUNetTrainer.on_train_epoch_end(ref, all_tr)
        T.nnUNetTrainer.on_validation_epoch_end(ref, all_va)
        ok = True
        for k in ('train_losses', 'val_losses', 'mean_fg_dice'):
            ok = ok and abs(float(me.logger.my_fantastic_logging[k][0]) - float(ref.logger.my_fantastic_logging[k][0])) < 1e-6
        ok = ok and np.allclose(me.logger.my_fantastic_logging['dice_per_class_or_region'][0],
                                ref.logger.my_fantastic_logging['dice_per_class_or_region'][0], rtol=0, atol=1e-12)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_epoch_end_hooks_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_epoch_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(2))
    assert res == {0: True, 1: True}


def _broadcast_worker(rank, world, port, q):
    """ranks seed differently (the reference's run_training never seeds); after the start-up synchronisation every
    replica holds rank 0's weights, and a few data-parallel SGD steps on the arena's summed gradients keep them equal."""
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        from multimodal_mvd_seg_b200.ddp import GradArena, broadcast_parameters, replicas_identical
        torch.manual_seed(1000 + rank)
        nets = [torch.nn.Sequential(torch.nn.Conv3d(2, 4, 3), torch.nn.InstanceNorm3d(4, affine=True)),
                torch.nn.Linear(5, 3)]
        before = replicas_identical(nets)
        broadcast_parameters(nets, src=0)
        after = replicas_identical(nets)
        params = [p for n in nets for p in n.parameters()]
        arena = GradArena(params, bucket_bytes=256)
        opt = torch.optim.SGD(params, 0.1, momentum=0.9, nesterov=True)
        for step in range(3):
            arena.begin_step()
            g = torch.Generator().manual_seed(7 * rank + step)      # every rank sees different data
            for p in reversed(params):
                arena.view_for(p).copy_(torch.randn(p.shape, generator=g))
                arena.on_params_ready([p])
            arena.finish()
            arena.attach_grads()
            for p in params:
                p.grad = p.grad / world
            opt.step()
        q.put((rank, (bool(before), bool(after), bool(replicas_identical(nets)))))
    finally:
        dist.destroy_process_group()


def test_replicas_start_and_stay_identical_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_broadcast_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=5) for _ in range(2))
    # differently seeded before the broadcast, identical after it and after the steps
    assert res == {0: (False, True, True), 1: (False, True, True)}
