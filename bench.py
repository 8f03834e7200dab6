#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: 3d_fullres train patches/sec (128^3, 2-modality).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg4]

ours      : one "step" = one full training step of the hot path (H2D excluded for `value`, included for `e2e`):
            PlainConvUNet fwd, deep-supervision Dice+CE, bwd, gradient exchange (N>1), clip+SGD-nesterov, all through
            libmvdseg kernels.  Workload at N=1 = BASELINE.json configs[1] (cfg2: one 2-channel net, 128^3, batch 2);
            cfg3 / cfg4 are the dual-network mutual-distillation (+clDice) steps.  N>1: batch-sharded weak scaling
            (2 patches per GPU), one process per GPU under torchrun, NCCL all-reduce of gradients overlapped with bwd.
reference : the reference's CPU implementation of the same path (the oracle port: the reference itself cannot be
            imported) on the host cores, each step a bounded sample (1 patch, 64^3 crop, same network).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (patch, dual-network?, topo_iter)
    'cfg2': ((128, 128, 128), False, None),
    'cfg3': ((128, 128, 128), True, None),
    'cfg4': ((160, 160, 96), True, 3),
}
PER_GPU_BATCH = 2
N_CLASSES = 4


def conv_flops_per_step(patch, batch, cin, dual):
    """algorithmic conv FLOPs of fwd+bwd (2*B*Vout*Cout*Cin*taps per pass; dgrad = wgrad = fprop,
    the stem needs no dgrad; transposed convs one tap per output voxel)."""
    import oracle
    topo = oracle.topology_for_patch(patch)
    feats, strides = topo['features_per_stage'], topo['strides']
    sizes, cur = [], list(patch)
    for s in strides:
        cur = [c // k for c, k in zip(cur, s)]
        sizes.append(list(cur))
    vol = lambda sz: sz[0] * sz[1] * sz[2]
    fwd, no_dgrad = 0.0, 0.0
    c_prev = cin
    for i, f in enumerate(feats):
        a = 2.0 * batch * vol(sizes[i]) * f * c_prev * 27
        if i == 0:
            no_dgrad = a
        fwd += a + 2.0 * batch * vol(sizes[i]) * f * f * 27
        c_prev = f
    for s in range(1, len(feats)):
        below, skip, sz = feats[-s], feats[-(s + 1)], sizes[-(s + 1)]
        fwd += 2.0 * batch * vol(sz) * below * skip                 # transposed conv
        fwd += 2.0 * batch * vol(sz) * skip * (2 * skip) * 27 + 2.0 * batch * vol(sz) * skip * skip * 27
        fwd += 2.0 * batch * vol(sz) * N_CLASSES * skip             # 1x1x1 head
    total = 3.0 * fwd - no_dgrad
    return total * (2 if dual else 1)


def sample_clocks_start():
    f = tempfile.NamedTemporaryFile(prefix='clocks_', suffix='.csv', delete=False)
    f.close()
    q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')
    try:
        p = subprocess.Popen(['nvidia-smi', f'--query-gpu={q}', '--format=csv,noheader,nounits', '-lms', '100',
                              '-i', os.environ.get('LOCAL_RANK', '0')], stdout=open(f.name, 'w'),
                             stderr=subprocess.DEVNULL)
    except Exception:
        return None, f.name
    return p, f.name


def sample_clocks_stop(p, path):
    out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
    if p is not None:
        p.terminate()
        try:
            p.wait(5)
        except Exception:
            p.kill()
    try:
        rows = [r.split(',') for r in open(path).read().strip().split('\n') if r.strip()]
        sm = sorted(float(r[1]) for r in rows)
        out['sm_mhz'] = sm[len(sm) // 2]
        out['sm_max_mhz'] = float(rows[0][2])
        out['power_w_max'] = max(float(r[3]) for r in rows)
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for i, n in enumerate(names):
            if any('Active' in r[5 + i] and 'Not' not in r[5 + i] for r in rows):
                out['reasons'].append(n)
        out['samples'] = len(rows)
    except Exception:
        pass
    try:
        os.unlink(path)
    except OSError:
        pass
    return out


def reference_step_time(steps, warmup, dual, topo_iter, threads=None):
    """the oracle port on the host cores; each step = 1 patch of a 64^3 crop through the SAME network as the workload
    (6 stages 32..320, deep supervision, DC+CE [+KL +clDice], backward, clip+SGD)."""
    import torch
    import oracle
    torch.set_num_threads(threads or os.cpu_count())   # as the reference CLI does for -device cpu (run_training.py:391-395)
    full_patch = (128, 128, 128)
    topo = oracle.topology_for_patch(full_patch)
    crop = (64, 64, 64)
    nets = [oracle.PlainConvUNet(1 if dual else 2, num_classes=N_CLASSES, **topo) for _ in range(2 if dual else 1)]
    for i, n in enumerate(nets):
        torch.manual_seed(i)
        n.apply(oracle.InitWeights_He(1e-2))
    params = [p for n in nets for p in n.parameters()]
    opt = torch.optim.SGD(params, 1e-2, weight_decay=3e-5, momentum=0.99, nesterov=True)
    batch = oracle.make_batch(1, 2, crop, topo['strides'], kind='rand')
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        if dual:
            l, _ = oracle.mvd_step_loss(nets[0], nets[1], batch['data'], batch['target'], topo_iter=topo_iter)
        else:
            l, _ = oracle.single_net_step_loss(nets[0], batch['data'], batch['target'])
        l.backward()
        torch.nn.utils.clip_grad_norm_(params, 12)
        opt.step()
        float(l.detach())
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    frac = (crop[0] * crop[1] * crop[2]) / float(full_patch[0] * full_patch[1] * full_patch[2])
    return times, frac, torch.get_num_threads()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg2', choices=list(WORKLOADS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--per-layer', action='store_true', help='print a per-layer conv timing table to stderr')
    ap.add_argument('--no-graph', action='store_true', help='launch every kernel eagerly instead of replaying a CUDA graph')
    ap.add_argument('--split-graph', type=int, default=1, help='1: forward and backward as two CUDA graphs so that the H2D copy of the targets overlaps the forward pass (affects e2e only)')
    args = ap.parse_args()
    # stdout carries exactly ONE line (the JSON): libraries that write to fd 1 on their own (NCCL prints its version
    # banner there) are pointed at stderr for the whole run; emit() writes the result to the real stdout
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(obj) + '\n').encode())
    args.warmup = max(args.warmup, 3) if args.impl == 'ours' else args.warmup
    patch, dual, topo_iter = WORKLOADS[args.workload]
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    metric = '3d_fullres train patches/sec (128^3, 2-modality)'
    wl_name = {'cfg2': 'cfg2: PlainConvUNet(2ch) training step, 128^3 patches, batch 2/GPU, DC+CE deep supervision',
               'cfg3': 'cfg3: dual-network mutual-distillation step (+KL), 128^3, batch 2/GPU',
               'cfg4': 'cfg4: dual-network mutual distillation + soft-clDice, 160x160x96, batch 2/GPU'}[args.workload]

    if args.impl == 'reference':
        if rank != 0:
            return
        steps = max(1, args.steps)
        times, frac, threads = reference_step_time(steps, max(0, min(args.warmup, 2)), dual, topo_iter)
        ms = 1e3 * sum(times) / len(times)
        val = frac / (ms / 1e3)
        line = {'metric': metric, 'value': val, 'unit': 'patches/s', 'n_gpus': args.gpus, 'steps': steps,
                'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'impl': 'reference',
                'config': {'workload': wl_name, 'global_batch': PER_GPU_BATCH * args.gpus},
                'cpu_baseline': {'value': val, 'unit': 'patches/s', 'cores': threads, 'kind': 'port',
                                 'sample': 'each step = 1 patch, 64^3 crop (1/8 of a 128^3 patch), same 6-stage '
                                           'network, fp32 CPU (oracle port of the reference step)'},
                'e2e': {'value': val, 'unit': 'patches/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        emit(line)
        return

    import torch
    import torch.distributed as dist
    import multimodal_mvd_seg_b200 as m
    import oracle
    assert torch.cuda.is_available(), 'bench.py --impl ours needs a GPU: the product path has no CPU fallback'
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    n_gpus = world

    plans, dj = m.make_plans(patch, batch_size=PER_GPU_BATCH * n_gpus, n_modalities=2, n_classes=N_CLASSES)
    if dual:
        tr = m.MVDTrainer(plans, '3d_fullres', 0, dj, device=dev, topo_iter=topo_iter)
    else:
        tr = m.nnUNetTrainer(plans, '3d_fullres', 0, dj, device=dev)
    torch.manual_seed(0)
    tr.initialize()
    assert tr.batch_size == PER_GPU_BATCH
    strides = plans['configurations']['3d_fullres']['pool_op_kernel_sizes']
    host = oracle.make_batch(PER_GPU_BATCH, 2, patch, strides, max_label=N_CLASSES - 1, seed=1234 + rank, kind='rand')
    host = {'data': host['data'].pin_memory(), 'target': [t.pin_memory() for t in host['target']]}
    resident = {'data': host['data'].to(dev), 'target': [t.to(dev) for t in host['target']]}
    h2d = host['data'].numel() * 4 + sum(t.numel() * 4 for t in host['target'])
    tr.on_train_epoch_start()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- instrumented eager pass: every conv launch bracketed by CUDA events (the live per-kernel roofline numbers)
    # and the launch count of one step; doubles as warm-up
    for _ in range(2):
        tr.train_step_async(resident)
    barrier()
    timer = m.ops.ConvTimer()
    m.ops.set_conv_timer(timer)
    m.lib.reset_launch_count()
    n_instr = max(2, min(args.steps, 5))
    for _ in range(n_instr):
        tr.train_step_async(resident)
    barrier()
    launches_per_step = m.lib.launch_count() / n_instr
    conv = timer.summary()
    per_layer = timer.per_layer() if args.per_layer else None
    m.ops.set_conv_timer(None)
    for d in conv.values():
        d['ms'] /= n_instr
        d['flops'] /= n_instr
        d['launches'] /= n_instr
    if per_layer:
        for d in per_layer.values():
            d['ms'] /= n_instr
            d['flops'] /= n_instr

    # ---- warm-up of the timed configuration (CUDA-graph capture happens here)
    tr.use_cuda_graph = not args.no_graph
    tr.split_graph = bool(args.split_graph)
    tr.graph_warmup_steps = 0
    for _ in range(args.warmup):
        tr.train_step_async(resident)
    barrier()

    # ---- device-resident timed region (value)
    clk_p, clk_f = sample_clocks_start() if rank == 0 else (None, None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        tr.train_step_async(resident)
    e1.record()
    barrier()
    launches = launches_per_step * args.steps
    clocks = sample_clocks_stop(clk_p, clk_f) if rank == 0 else None
    ms_total = e0.elapsed_time(e1)

    # ---- end-to-end through the public API: trainer.train_step(host batch) -> {'loss': np.ndarray}
    for b in tr.prefetching([dict(host) for _ in range(3)]):   # untimed: staging buffers, copy stream, first replays
        tr.train_step(b)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, args.steps // 2)
    # the training loop a user writes: host batches (pinned, as nnU-Net's augmenter hands them over) go through
    # trainer.prefetching(), which uploads batch i+1 underneath step i; every step still pays its own H2D + loss D2H
    for b in tr.prefetching([dict(host) for _ in range(e2e_steps)]):
        out = tr.train_step(b)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    loss_val = float(out['loss'])

    t = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_s = float(t[0]), float(t[1])
    def finish():
        # leave without tearing NCCL down: destroy_process_group() can block behind the communicator references a
        # captured CUDA graph holds; all ranks meet at a barrier first so nobody exits under a peer's collective
        sys.stdout.flush()
        sys.stderr.flush()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            os._exit(0)

    if rank != 0:
        finish()
        return

    ms_per_step = ms_total / args.steps
    value = PER_GPU_BATCH * n_gpus / (ms_per_step / 1e3)
    e2e_val = PER_GPU_BATCH * n_gpus / (e2e_s / e2e_steps)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    peak_tf = float(peaks.get('bf16_tflops_sustained', 1400.0))
    peak_src = 'MEASURED_PEAKS.json bf16_tflops_sustained (kernel timed inside a long step)' if peaks else \
        'fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)'
    tot_flops = sum(d['flops'] for d in conv.values())     # per step
    tot_ms = sum(d['ms'] for d in conv.values())           # per step
    achieved = tot_flops / (tot_ms / 1e3) / 1e12 if tot_ms > 0 else 0.0
    roofline = {'bound': 'tensor', 'achieved': achieved, 'peak': peak_tf, 'unit': 'TFLOP/s',
                'frac': achieved / peak_tf, 'traffic': None, 'peak_source': peak_src,
                'kernel': 'conv3d fprop+dgrad+wgrad (all conv launches of the step)',
                'conv_share_of_step': tot_ms / ms_per_step,
                'timing': 'CUDA events around every conv launch in an instrumented eager pass of the same step',
                'per_pass': {k: {'tflops': d['flops'] / (d['ms'] / 1e3) / 1e12, 'ms_per_step': d['ms'],
                                 'launches_per_step': d['launches']} for k, d in conv.items()},
                'algorithmic_conv_gflop_per_step': conv_flops_per_step(patch, PER_GPU_BATCH, 1 if dual else 2, dual) / 1e9}
    try:   # DRAM bytes of the step's largest kernel from the committed ncu --set full capture (per launch)
        tr_ = json.load(open(os.path.join(ROOT, 'profiles', 'r1_ncu_traffic.json')))
        roofline['traffic'] = tr_['dram_bytes_per_launch']
        roofline['traffic_kernel'] = tr_['kernel']
        roofline['traffic_algorithmic_bytes'] = tr_['algorithmic_bytes_per_launch']
    except Exception:
        pass
    line = {'metric': metric, 'value': value, 'unit': 'patches/s', 'n_gpus': n_gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': wl_name, 'global_batch': PER_GPU_BATCH * n_gpus, 'patch': list(patch),
                       'parallelism': f'dp{n_gpus}', 'cuda_graph': bool(tr.use_cuda_graph), 'l2': 'inputs larger than L2 (GBs of activations per step)',
                       'voxels_per_s': value * patch[0] * patch[1] * patch[2]},
            'e2e': {'value': e2e_val, 'unit': 'patches/s', 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 4,
                    'steps': e2e_steps},
            'gpu_launches': int(launches), 'gpu_launches_per_step': launches / args.steps,
            'roofline': roofline, 'clocks': clocks, 'loss': loss_val}
    if n_gpus == 1 and not args.no_cpu_baseline:
        times, frac, threads = reference_step_time(30, 1, dual, topo_iter)   # ~10-12 s of host work
        s = sum(times) / len(times)
        line['cpu_baseline'] = {'value': frac / s, 'unit': 'patches/s', 'cores': threads, 'kind': 'port',
                                'sample': '30 steps of 1 patch, 64^3 crop (1/8 of a 128^3 patch) through the same '
                                          'network on the host cores (oracle port, fp32)', 'seconds': sum(times)}
    if per_layer:
        rows = sorted(per_layer.items(), key=lambda kv: -kv[1]['ms'])
        for (kind, tag), d in rows:
            print(f'{kind:6s} {tag:44s} {d["ms"]:8.3f} ms/step '
                  f'{d["flops"] / (d["ms"] / 1e3) / 1e12:8.1f} TFLOP/s', file=sys.stderr)
    emit(line)
    finish()


if __name__ == '__main__':
    main()
