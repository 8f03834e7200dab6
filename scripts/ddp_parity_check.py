#!/usr/bin/env python
"""Multi-GPU numerical check of the batch-sharded step (run under torchrun, one rank per GPU, NCCL):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 \
      scripts/ddp_parity_check.py [--workload cfg2|cfg4] [--patch 64]

  1. ranks seed DIFFERENTLY (the reference's run_training never seeds): after trainer.initialize() every replica must
     hold rank 0's weights (what DDP's constructor guarantees, MVDTrainer.py:236-238);
  2. K steps on different data per rank: replicas still bit-identical (same all-reduced gradient, same update);
  3. one step at world N on the split batch (2 patches per rank) equals (a) the same shares stepped one after the other
     in ONE process with the gradients summed by hand (identical kernel plans: update within 5e-3) and (b) one step of
     a single-process trainer on the UNION batch of 2N patches (other kernel plans -> other bf16 roundings, amplified by
     the ill-conditioned random-init gradient: reported, bound 0.15).  Optionally with batch_dice=True (the AllGatherGrad
     branch, ddp_allgather.py:25-48; (a) is skipped there: the Dice sums couple the ranks).
Prints one JSON line on rank 0; exit code 1 on failure.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='cfg2', choices=['cfg2', 'cfg4'])
    ap.add_argument('--patch', type=int, nargs='*', default=None, help='override the patch (default: a reduced one)')
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--batch-dice', type=int, default=0)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import multimodal_mvd_seg_b200 as m
    from multimodal_mvd_seg_b200.ddp import replicas_identical
    import oracle
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    dual = args.workload == 'cfg4'
    patch = tuple(args.patch) if args.patch else ((80, 80, 48) if dual else (64, 64, 64))
    per_gpu = 2
    plans, dj = m.make_plans(patch, batch_size=per_gpu * world, n_modalities=2, n_classes=4,
                             batch_dice=bool(args.batch_dice))
    mk = (lambda: m.MVDTrainer(plans, '3d_fullres', 0, dj, device=dev, topo_iter=3)) if dual else \
        (lambda: m.nnUNetTrainer(plans, '3d_fullres', 0, dj, device=dev))
    torch.manual_seed(1000 + rank)                  # (1) different initialisation on every rank
    tr = mk()
    tr.initialize()
    assert tr.batch_size == per_gpu
    res = {'world': world, 'workload': args.workload, 'patch': list(patch), 'batch_dice': bool(args.batch_dice)}
    res['identical_after_initialize'] = replicas_identical(tr._networks())
    strides = plans['configurations']['3d_fullres']['pool_op_kernel_sizes']
    tr.on_train_epoch_start()
    w0 = [p.detach().clone() for n in tr._networks() for p in n.parameters()]
    # (3) first: ONE step from w0 on this rank's share of a union batch
    union = oracle.make_batch(per_gpu * world, 2, patch, strides, kind='structured', seed=77)
    share = {'data': union['data'][rank * per_gpu:(rank + 1) * per_gpu],
             'target': [t[rank * per_gpu:(rank + 1) * per_gpu] for t in union['target']]}
    loss_ddp = float(tr.train_step(share)['loss'])
    w1 = [p.detach().clone() for n in tr._networks() for p in n.parameters()]
    res['identical_after_split_step'] = replicas_identical(tr._networks())
    # (2) more steps on different data per rank
    for i in range(args.steps):
        b = oracle.make_batch(per_gpu, 2, patch, strides, kind='structured', seed=500 + 31 * rank + i)
        tr.train_step(b)
    res['identical_after_steps'] = replicas_identical(tr._networks())
    ok = res['identical_after_initialize'] and res['identical_after_split_step'] and res['identical_after_steps']
    if rank == 0:
        def update_err(params_new):
            num = den = 0.0
            for p, a, b in zip(params_new, w1, w0):
                num += float(((p.detach() - b) - (a - b)).double().pow(2).sum())
                den += float((p.detach() - b).double().pow(2).sum())
            return (num / max(den, 1e-300)) ** 0.5

        def fresh_solo(batch_size):
            t = mk()
            t.is_ddp = False
            t.batch_size = batch_size
            t.initialize()
            for p, w in zip([p for n in t._networks() for p in n.parameters()], w0):
                with torch.no_grad():
                    p.copy_(w)
            t.on_train_epoch_start()
            return t

        # (3a) the SAME arithmetic in one process: every rank's share stepped with the per-GPU batch size (identical
        # kernel plans), gradients summed by hand, one optimiser step with grad_scale = 1/world  (not with batch_dice:
        # the Dice sums of the shares are coupled through the all-gather there)
        emu = fresh_solo(per_gpu)
        total = None
        for r in range(world if not args.batch_dice else 0):
            sh = {'data': union['data'][r * per_gpu:(r + 1) * per_gpu],
                  'target': [t[r * per_gpu:(r + 1) * per_gpu] for t in union['target']]}
            data, target = emu._to_device(sh)
            out = emu._step_forward(data)
            l, _ = emu._loss(out, target)
            l.backward()
            if getattr(emu, '_net2_stream', None) is not None:
                torch.cuda.current_stream().wait_stream(emu._net2_stream)
            for a in emu._arenas:
                a.finish()
            flats = [a.flat.clone() for a in emu._arenas]
            total = flats if total is None else [x + y for x, y in zip(total, flats)]
        if total is not None:
            for a, f in zip(emu._arenas, total):
                a.flat.copy_(f)
                a.attach_grads()
            emu.optimizer.step(grad_scale=1.0 / world)
            res['split_vs_emulated_update_rel_err'] = update_err([p for n in emu._networks() for p in n.parameters()])
        else:
            res['split_vs_emulated_update_rel_err'] = 0.0
        del emu
        # (3b) ONE step on the union batch (2N patches in one process).  Same mathematics, but the kernels run with other
        # plans (tile shapes / reduction orders depend on the batch), i.e. other bf16 rounding decisions, and the
        # whole-network gradient at random initialisation amplifies such 1-ulp differences (the same step repeated
        # twice differs by ~2e-4 from atomics order alone): reported, with a loose sanity bound
        solo = fresh_solo(per_gpu * world)
        loss_solo = float(solo.train_step(union)['loss'])
        res['split_vs_union_update_rel_err'] = update_err([p for n in solo._networks() for p in n.parameters()])
        res['loss_rank0_share'] = loss_ddp
        res['loss_union'] = loss_solo
        ok = ok and res['split_vs_emulated_update_rel_err'] < 5e-3 and res['split_vs_union_update_rel_err'] < 0.15
        res['ok'] = bool(ok)
        print(json.dumps(res), flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    torch.cuda.synchronize()
    dist.barrier()
    sys.stdout.flush()
    os._exit(0 if int(flag) == 1 else 1)


if __name__ == '__main__':
    main()
