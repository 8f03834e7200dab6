#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: 3d_fullres train patches/sec (128^3, 2-modality).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg4]

ours      : one "step" = one full training step of the hot path (H2D excluded for `value`, included for `e2e`):
            PlainConvUNet fwd, deep-supervision Dice+CE, bwd, gradient exchange (N>1), clip+SGD-nesterov, all through
            libmvdseg kernels.  Workload at N=1 = BASELINE.json configs[1] (cfg2: one 2-channel net, 128^3, batch 2);
            cfg3 / cfg4 (the dual-network mutual-distillation steps, +clDice) are measured in the same run and reported
            under `extra_workloads` (at N>1: cfg4 per GPU = BASELINE.json configs[4], "cfg5").  N>1: batch-sharded weak
            scaling (2 patches per GPU), one process per GPU under torchrun, NCCL all-reduce of gradients overlapped
            with backward; the line carries `ddp_weights_identical` (parameter checksums all-gathered after the run).
reference : the reference's CPU implementation of the same path (the oracle port: the reference itself cannot be
            imported) on the host cores, each step one FULL step of the same workload configuration
            (bounded number of steps).
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
WL_NAME = {'cfg2': 'cfg2: PlainConvUNet(2ch) training step, 128^3 patches, batch 2/GPU, DC+CE deep supervision',
           'cfg3': 'cfg3: dual-network mutual-distillation step (+KL), 128^3, batch 2/GPU',
           'cfg4': 'cfg4: dual-network mutual distillation + soft-clDice(iter 3), 160x160x96, batch 2/GPU'}
PER_GPU_BATCH = 2
N_CLASSES = 4
METRIC = '3d_fullres train patches/sec (128^3, 2-modality)'


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


# ----------------------------------------------------------------------------------------------------------------
# CPU legs: the oracle port of the reference step on the host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_port_step_times(patch, batch, dual, topo_iter, steps, warmup, threads=None):
    """`steps` full training steps of one configuration on the host (fp32, no autocast: the reference disables
    autocast on CPU, MVDTrainer.py:894): network(s) of the patch's own topology, deep-supervision DC+CE [+KL +clDice],
    backward, clip_grad_norm_(12), SGD nesterov.  Returns (list of seconds per step, threads used)."""
    import torch
    import oracle
    torch.set_num_threads(threads or os.cpu_count())   # as the reference CLI does for -device cpu (run_training.py:391-395)
    topo = oracle.topology_for_patch(patch)
    nets = [oracle.PlainConvUNet(1 if dual else 2, num_classes=N_CLASSES, **topo) for _ in range(2 if dual else 1)]
    for i, n in enumerate(nets):
        torch.manual_seed(i)
        n.apply(oracle.InitWeights_He(1e-2))
    params = [p for n in nets for p in n.parameters()]
    opt = torch.optim.SGD(params, 1e-2, weight_decay=3e-5, momentum=0.99, nesterov=True)
    b = oracle.make_batch(batch, 2, patch, topo['strides'], kind='rand')
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        if dual:
            l, _ = oracle.mvd_step_loss(nets[0], nets[1], b['data'], b['target'], topo_iter=topo_iter)
        else:
            l, _ = oracle.single_net_step_loss(nets[0], b['data'], b['target'])
        l.backward()
        torch.nn.utils.clip_grad_norm_(params, 12)
        opt.step()
        float(l.detach())
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return times, torch.get_num_threads()


def gpu_augmentation_rate(patch, dev, batches=20):
    """the batch-production side.  The transform chain of MVDTrainer.py:700-765 on the GPU
    (multimodal_mvd_seg_b200/augment.py) over synthetic raw patches of the loader's initial patch size (1.25 x the patch),
    once with the reference's own probabilities (what a training run sees on average) and once with every transform
    switched on; includes the deep-supervision targets.  Device-resident inputs, CUDA events."""
    import numpy as np
    import torch
    from multimodal_mvd_seg_b200.augment import GpuAugmenter, sample_parameters
    raw = tuple(int(round(p * 1.25 / 8) * 8) for p in patch)
    data = torch.randn((PER_GPU_BATCH, 2) + raw, device=dev)
    seg = (torch.rand((PER_GPU_BATCH, 1) + raw, device=dev) * N_CLASSES).floor()
    scales = [[1.0 / 2 ** i] * 3 for i in range(5)]
    aug = GpuAugmenter(patch, N_CLASSES, deep_supervision_scales=scales, seed=0)
    rng = np.random.default_rng(0)
    drawn = [sample_parameters(rng, PER_GPU_BATCH, 2) for _ in range(batches)]
    full = sample_parameters(rng, PER_GPU_BATCH, 2)
    full['mode'][:] = 1
    for b in range(PER_GPU_BATCH):
        full['mat'][b] = np.eye(3) * 1.1
    full['noise_sigma'][:] = 0.05; full['blur_sigma'][:] = 0.8; full['brightness'][:] = 1.1; full['contrast'][:] = 0.9
    full['lowres_zoom'][:] = 0.7; full['gamma_inv'][:] = 1.2; full['gamma'][:] = 0.8; full['flips'][:] = 1
    out = {}
    for name, plist in (('reference_probabilities', drawn), ('all_transforms_on', [full] * 5)):
        aug(data, seg, params=plist[0])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for prm in plist:
            aug(data, seg, params=prm)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / len(plist)
        out[name] = {'ms_per_batch': ms, 'patches_per_s': PER_GPU_BATCH / (ms / 1e3), 'batches': len(plist)}
    out['what'] = (f'GPU augmentation chain (SpatialTransform order 3 / per-label order 1, noise, blur, brightness, contrast, '
                   f'SimulateLowResolution, 2 x gamma, mirror, DS targets) on {PER_GPU_BATCH} x 2 x {raw} raw patches -> {patch}; '
                   'context for the batch-production row, not part of the timed step')
    return out


def gpu_torch_context(patch, dual, topo_iter, dev, steps=5, warmup=2):
    """CONTEXT ONLY (not an arm of the comparison): the oracle modules -- stock torch.nn Conv3d / InstanceNorm3d /
    ConvTranspose3d, i.e. ATen + cuDNN -- stepping the same workload on the same GPU under bf16 autocast with
    cudnn.benchmark, fwd + loss + bwd + clip + SGD.  SURVEY.md section 0 fact 6: this is the library-level competitor."""
    import torch
    import oracle
    prev = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True      # run_training.py:245-247
    try:
        topo = oracle.topology_for_patch(patch)
        nets = [oracle.PlainConvUNet(1 if dual else 2, num_classes=N_CLASSES, **topo).to(dev) for _ in range(2 if dual else 1)]
        for i, n in enumerate(nets):
            torch.manual_seed(i)
            n.apply(oracle.InitWeights_He(1e-2))
        params = [p for n in nets for p in n.parameters()]
        opt = torch.optim.SGD(params, 1e-2, weight_decay=3e-5, momentum=0.99, nesterov=True)
        b = oracle.make_batch(PER_GPU_BATCH, 2, patch, topo['strides'], kind='rand')
        data, target = b['data'].to(dev), [t.to(dev) for t in b['target']]

        def step():
            opt.zero_grad(set_to_none=True)
            if dual:
                l, _ = oracle.mvd_step_loss(nets[0], nets[1], data, target, topo_iter=topo_iter, autocast_bf16=True)
            else:
                l, _ = oracle.single_net_step_loss(nets[0], data, target, autocast_bf16=True)
            l.backward()
            torch.nn.utils.clip_grad_norm_(params, 12)
            opt.step()
            return l
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        del nets, params, opt, data, target
        torch.cuda.empty_cache()
        return {'what': 'CONTEXT, not the reference arm: oracle modules (torch.nn -> ATen/cuDNN) on this GPU, '
                        'torch.autocast(bf16), cudnn.benchmark=True, eager, device-resident inputs',
                'ms_per_step': ms, 'value': PER_GPU_BATCH / (ms / 1e3), 'unit': 'patches/s', 'steps': steps}
    except Exception as e:      # context only: never fail the bench over it
        return {'unavailable': f'{type(e).__name__}: {e}'[:200]}
    finally:
        torch.backends.cudnn.benchmark = prev


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def measure_workload(name, args, rank, world, dev, steps, with_e2e=True, with_clocks=False):
    """builds the trainer of one workload, runs (i) an instrumented eager pass (CUDA events around every conv launch
    and every HBM-bound launch; launch count), (ii) the device-resident timed region, (iii) the end-to-end region
    through trainer.train_step(host batch).  Returns a dict of raw measurements (max over ranks where timed)."""
    import torch
    import torch.distributed as dist
    import multimodal_mvd_seg_b200 as m
    import oracle
    patch, dual, topo_iter = WORKLOADS[name]
    n_gpus = world
    plans, dj = m.make_plans(patch, batch_size=PER_GPU_BATCH * n_gpus, n_modalities=2, n_classes=N_CLASSES)
    if dual:
        tr = m.MVDTrainer(plans, '3d_fullres', 0, dj, device=dev, topo_iter=topo_iter)
    else:
        tr = m.nnUNetTrainer(plans, '3d_fullres', 0, dj, device=dev)
    torch.manual_seed(1234 + rank)      # ranks initialise differently; initialize() synchronises them (rank 0's weights)
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

    # ---- instrumented eager pass (doubles as warm-up)
    for _ in range(2):
        tr.train_step_async(resident)
    barrier()
    conc = getattr(tr, 'concurrent_networks', None)
    if conc:       # per-kernel event timing wants one stream: the two networks run back to back in this pass only
        tr.concurrent_networks = False
    timer = m.ops.ConvTimer()
    m.ops.set_conv_timer(timer)
    m.lib.reset_launch_count()
    m.lib.reset_fallback_count()
    n_instr = max(2, min(steps, 5))
    for _ in range(n_instr):
        tr.train_step_async(resident)
    barrier()
    launches_per_step = m.lib.launch_count() / n_instr
    fallbacks = int(m.lib.fallback_count())
    # ---- gradient exchange timeline (N > 1): eager steps with the overlap ON (no per-kernel timer), CUDA events around
    # every bucket's all-reduce on the comm stream and around the compute stream's final wait for the exchange
    ddp_timeline = None
    if world > 1:
        m.ops.set_conv_timer(None)
        for a in tr._arenas:
            a.arm_timeline(True)
        for _ in range(3):
            tr.train_step_async(resident)
        barrier()
        per_arena = [a.timeline_summary() for a in tr._arenas]
        for a in tr._arenas:
            a.arm_timeline(False)
        steps_rec = [list(x) for x in zip(*[p for p in per_arena if p])]
        if steps_rec:
            last = steps_rec[-1]       # the last recorded step, all arenas (networks)
            ddp_timeline = dict(arenas=len(last), buckets=sum(r['buckets'] for r in last),
                                bytes_per_step=sum(r['bytes'] for r in last),
                                allreduce_ms_sum=sum(r['allreduce_ms_sum'] for r in last),
                                exposed_ms=sum(r['exposed_ms'] for r in last),
                                first_allreduce_after_step_start_ms=min(r['first_allreduce_after_step_start_ms'] for r in last),
                                note='eager step with backward overlap on; exposed = compute stream waiting for the '
                                     'exchange after backward (rank 0)')
        m.ops.set_conv_timer(timer)
    m.ops.set_conv_timer(None)
    conv = timer.summary()
    mem = timer.mem_summary()
    per_layer = timer.per_layer()
    for d in list(conv.values()) + list(mem.values()):
        for k in d:
            d[k] /= n_instr
    if per_layer:
        for d in per_layer.values():
            d['ms'] /= n_instr
            d['flops'] /= n_instr
            d['launches'] /= n_instr

    if conc:
        tr.concurrent_networks = True
    # ---- warm-up of the timed configuration (CUDA-graph capture happens here)
    tr.use_cuda_graph = not args.no_graph
    tr.split_graph = bool(args.split_graph)
    tr.graph_warmup_steps = 0
    for _ in range(max(args.warmup, 3)):
        tr.train_step_async(resident)
    barrier()

    # ---- device-resident timed region (value)
    clk_p, clk_f = sample_clocks_start() if (rank == 0 and with_clocks) else (None, None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        tr.train_step_async(resident)
    e1.record()
    barrier()
    clocks = sample_clocks_stop(clk_p, clk_f) if (rank == 0 and with_clocks) else None
    ms_total = e0.elapsed_time(e1)

    # ---- end-to-end through the public API: trainer.train_step(host batch) -> {'loss': np.ndarray}
    e2e_s, e2e_steps, loss_val = 0.0, 0, None
    if with_e2e:
        for b in tr.prefetching([dict(host) for _ in range(3)]):   # untimed: staging buffers, copy stream, first replays
            tr.train_step(b)
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(3, steps // 2)
        # the training loop a user writes: host batches (pinned, as nnU-Net's augmenter hands them over) go through
        # trainer.prefetching(), which uploads batch i+1 underneath step i; every step pays its own H2D + loss D2H
        for b in tr.prefetching([dict(host) for _ in range(e2e_steps)]):
            out = tr.train_step(b)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        loss_val = float(out['loss'])

    t = torch.tensor([ms_total, e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_s = float(t[0]), float(t[1])
    identical = None
    if world > 1:
        from multimodal_mvd_seg_b200.ddp import replicas_identical
        identical = bool(replicas_identical(tr._networks()))
    res = dict(name=name, patch=patch, dual=dual, ms_per_step=ms_total / steps, steps=steps, e2e_s=e2e_s,
               e2e_steps=e2e_steps, h2d=h2d, launches_per_step=launches_per_step, fallbacks=fallbacks, conv=conv,
               mem=mem, per_layer=per_layer, clocks=clocks, loss=loss_val, cuda_graph=bool(tr.use_cuda_graph),
               ddp_weights_identical=identical, ddp_timeline=ddp_timeline)
    # release this workload's device memory before the next one is built
    del tr, resident, host
    m.ops.set_grad_allocator(None)
    m.ops.clear_param_grad_ready_hooks()
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    return res


def conv_roofline(r, peaks, clocks):
    """tensor-core roofline of all conv launches of one step.  Denominator: the BURST bf16 peak when the SM clock
    sampled during the timed region stayed at >= 95 % of its maximum (the step does not pull the GPU into the power-capped
    regime the sustained figure was measured in), else the sustained peak; both fractions are reported."""
    burst = float(peaks.get('bf16_tflops', 1675.0))
    sustained = float(peaks.get('bf16_tflops_sustained', 1400.0))
    conv = r['conv']
    tot_flops = sum(d['flops'] for d in conv.values())
    tot_ms = sum(d['ms'] for d in conv.values())
    achieved = tot_flops / (tot_ms / 1e3) / 1e12 if tot_ms > 0 else 0.0
    at_full_clock = bool(clocks and clocks.get('sm_mhz') and clocks.get('sm_max_mhz') and
                         clocks['sm_mhz'] >= 0.95 * clocks['sm_max_mhz'])
    use_burst = at_full_clock or clocks is None or clocks.get('sm_mhz') is None
    peak = burst if use_burst else sustained
    src = 'MEASURED_PEAKS.json' if peaks else 'fallback of B200_PROFILING.md'
    return {'bound': 'tensor', 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s', 'frac': achieved / peak,
            'traffic': None,
            'peak_source': f'{src} bf16_tflops ({"burst: SM clock >= 95 % of max during the timed region" if use_burst else "sustained: SM clock below 95 % of max during the timed region"})',
            'frac_vs_burst': achieved / burst, 'frac_vs_sustained': achieved / sustained,
            'peak_burst': burst, 'peak_sustained': sustained,
            'kernel': 'conv3d fprop+dgrad+wgrad (all conv launches of the step)',
            'conv_share_of_step': tot_ms / r['ms_per_step'],
            'timing': 'CUDA events around every conv launch in an instrumented eager pass of the same step; a 0.1 ms busy-wait kernel queued ahead of each bracket keeps the host launch path out of the interval',
            'per_pass': {k: {'tflops': d['flops'] / (d['ms'] / 1e3) / 1e12, 'ms_per_step': d['ms'],
                             'launches_per_step': d['launches']} for k, d in conv.items()},
            'algorithmic_conv_gflop_per_step': conv_flops_per_step(r['patch'], PER_GPU_BATCH, 1 if r['dual'] else 2,
                                                                   r['dual']) / 1e9,
            'dominant_kernel': dominant_kernel(r, burst, sustained)}


def dominant_kernel(r, burst, sustained):
    """the single kernel with the largest share of the step: wgrad_halo_kernel = the weight gradients of all 3x3x3 /
    stride-1 layers (profiles/r2_step_profile_cfg2.txt: 19 % of the in-step kernel time, 16 launches per network).
    Algorithmic FLOPs of those launches / their event-timed duration in the instrumented pass."""
    pl = r.get('per_layer') or {}
    sel = [d for (kind, tag), d in pl.items() if kind == 'wgrad' and 'k333 s111' in tag and not tag.startswith('stem')]
    ms = sum(d['ms'] for d in sel)
    if not sel or ms <= 0:
        return None
    fl = sum(d['flops'] for d in sel)
    tf = fl / (ms / 1e3) / 1e12
    return {'name': 'wgrad_halo_kernel', 'achieved': tf, 'unit': 'TFLOP/s', 'frac_vs_burst': tf / burst,
            'frac_vs_sustained': tf / sustained, 'ms_per_step': ms, 'launches_per_step': sum(d['launches'] for d in sel),
            'gflop_per_step': fl / 1e9, 'share_of_step': ms / r['ms_per_step'],
            'ncu': 'profiles/r2_ncu_wgrad_halo.txt (tensor pipe 65 %, DRAM 557 MB for 537 MB algorithmic at 32->32, 2 x 128^3)'}


def hbm_roofline(r, peaks):
    """HBM-bound kernel families of the step: algorithmic bytes / event-timed duration, against
    the measured copy bandwidth."""
    peak = float(peaks.get('hbm_gbs', 6500.0))
    out = {}
    for name, d in sorted(r['mem'].items()):
        gbs = d['bytes'] / (d['ms'] / 1e3) / 1e9 if d['ms'] > 0 else 0.0
        out[name] = {'achieved': gbs, 'frac': gbs / peak, 'ms_per_step': d['ms'], 'launches_per_step': d['launches'],
                     'algorithmic_mb_per_step': d['bytes'] / 1e6}
    return {'bound': 'hbm', 'peak': peak, 'unit': 'GB/s',
            'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if peaks else 'fallback of B200_PROFILING.md',
            'timing': 'CUDA events around every launch of the family in the instrumented eager pass, each behind a 0.1 ms '
                      'busy-wait kernel (device time only; short launches still include their pipeline fill)', 'kernels': out}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg2', choices=list(WORKLOADS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extra', action='store_true', help='skip extra_workloads (cfg3/cfg4) and gpu_torch_context')
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

    if args.impl == 'reference':
        if rank != 0:
            return
        # bounded: full steps of the SAME configuration (2 x 128^3 through the 6-stage network is ~15-25 s per step on
        # the box's cores), at most 3 timed + 1 warm-up so that the run ends within a few minutes
        steps = max(1, min(args.steps, 3 if not dual else 1))
        warm = max(0, min(args.warmup, 1))
        times, threads = cpu_port_step_times(patch, PER_GPU_BATCH, dual, topo_iter, steps, warm)
        ms = 1e3 * sum(times) / len(times)
        val = PER_GPU_BATCH / (ms / 1e3)
        line = {'metric': METRIC, 'value': val, 'unit': 'patches/s', 'n_gpus': args.gpus, 'steps': steps,
                'warmup': warm, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'impl': 'reference',
                'config': {'workload': WL_NAME[args.workload], 'global_batch': PER_GPU_BATCH * args.gpus,
                           'patch': list(patch), 'parallelism': f'dp{args.gpus}'},
                'cpu_baseline': {'value': val, 'unit': 'patches/s', 'cores': threads, 'kind': 'port',
                                 'sample': f'{steps} full training step(s) of the same configuration (batch '
                                           f'{PER_GPU_BATCH}, {patch[0]}x{patch[1]}x{patch[2]}, same network) after {warm} warm-up, fp32 on '
                                           'the host cores (oracle port of the reference step; the reference itself '
                                           'cannot be imported); a bounded sample of the workload: one per-GPU batch '
                                           'per step whatever --gpus says (patches/s is a rate)', 'seconds': sum(times)},
                'e2e': {'value': val, 'unit': 'patches/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
        emit(line)
        return

    import torch
    import torch.distributed as dist
    import multimodal_mvd_seg_b200 as m  # noqa: F401
    assert torch.cuda.is_available(), 'bench.py --impl ours needs a GPU: the product path has no CPU fallback'
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    n_gpus = world

    main_r = measure_workload(args.workload, args, rank, world, dev, args.steps, with_e2e=True, with_clocks=True)
    extras = {}
    if not args.no_extra:
        # the other BASELINE.json configurations, measured in the same run (shorter): at N=1 cfg3 + cfg4, at N>1 the
        # full distillation + topology step batch-sharded over the N GPUs (configs[4], "cfg5" = cfg4 per GPU)
        names = [n for n in (('cfg3', 'cfg4') if world == 1 else ('cfg4',)) if n != args.workload]
        for n in names:
            try:
                extras[n] = measure_workload(n, args, rank, world, dev, max(5, min(args.steps, 10)), with_e2e=True)
            except Exception as e:      # a failing extra must not take the headline line down with it
                extras[n] = {'error': f'{type(e).__name__}: {e}'[:300]}

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

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    r = main_r
    ms_per_step = r['ms_per_step']
    value = PER_GPU_BATCH * n_gpus / (ms_per_step / 1e3)
    e2e_val = PER_GPU_BATCH * n_gpus / (r['e2e_s'] / r['e2e_steps'])
    roofline = conv_roofline(r, peaks, r['clocks'])
    try:   # DRAM bytes of the step's dominant kernel from the committed `ncu --set full` capture (per launch); cannot be
        # measured without the profiler, so the source file is named next to it
        for fn in ('r2_ncu_traffic.json', 'r1_ncu_traffic.json'):
            pth = os.path.join(ROOT, 'profiles', fn)
            if os.path.exists(pth):
                tr_ = json.load(open(pth))
                roofline['traffic'] = tr_['dram_bytes_per_launch']
                roofline['traffic_kernel'] = tr_['kernel']
                roofline['traffic_algorithmic_bytes'] = tr_['algorithmic_bytes_per_launch']
                roofline['traffic_source'] = f'profiles/{fn} (ncu --set full capture, not measured in this run)'
                break
    except Exception:
        pass
    line = {'metric': METRIC, 'value': value, 'unit': 'patches/s', 'n_gpus': n_gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
            'config': {'workload': WL_NAME[args.workload], 'global_batch': PER_GPU_BATCH * n_gpus, 'patch': list(patch),
                       'parallelism': f'dp{n_gpus}', 'cuda_graph': r['cuda_graph'],
                       'l2': 'inputs larger than L2 (GBs of activations per step)',
                       'voxels_per_s': value * patch[0] * patch[1] * patch[2]},
            'e2e': {'value': e2e_val, 'unit': 'patches/s', 'h2d_bytes_per_step': r['h2d'], 'd2h_bytes_per_step': 4,
                    'steps': r['e2e_steps']},
            'gpu_launches': int(r['launches_per_step'] * args.steps), 'gpu_launches_per_step': r['launches_per_step'],
            'generic_fallbacks_per_step': r['fallbacks'],
            'roofline': roofline, 'roofline_hbm': hbm_roofline(r, peaks), 'clocks': r['clocks'], 'loss': r['loss']}
    if n_gpus > 1:
        line['ddp_weights_identical'] = r['ddp_weights_identical']
        line['ddp_exchange'] = r['ddp_timeline']
    if extras:
        ex = {}
        for n, e in extras.items():
            key = n if n_gpus == 1 else 'cfg5'
            if 'error' in e:
                ex[key] = e
                continue
            rf = conv_roofline(e, peaks, r['clocks'])
            ex[key] = {'workload': WL_NAME[n] + ('' if n_gpus == 1 else f', batch-sharded over {n_gpus} GPUs (BASELINE.json configs[4])'),
                       'ms_per_step': e['ms_per_step'], 'value': PER_GPU_BATCH * n_gpus / (e['ms_per_step'] / 1e3),
                       'unit': 'patches/s', 'steps': e['steps'],
                       'e2e_value': PER_GPU_BATCH * n_gpus / (e['e2e_s'] / e['e2e_steps']),
                       'h2d_bytes_per_step': e['h2d'], 'gpu_launches_per_step': e['launches_per_step'],
                       'generic_fallbacks_per_step': e['fallbacks'], 'conv_tflops': rf['achieved'],
                       'conv_frac_vs_burst': rf['frac_vs_burst'], 'conv_share_of_step': rf['conv_share_of_step'],
                       'roofline_hbm': {k: {'achieved': v['achieved'], 'frac': v['frac'], 'ms_per_step': v['ms_per_step']}
                                        for k, v in hbm_roofline(e, peaks)['kernels'].items()},
                       'loss': e['loss']}
            if n_gpus > 1:
                ex[key]['ddp_weights_identical'] = e['ddp_weights_identical']
                ex[key]['ddp_exchange'] = e['ddp_timeline']
        line['extra_workloads'] = ex
    if n_gpus == 1 and not args.no_extra:
        line['gpu_torch_context'] = gpu_torch_context(patch, dual, topo_iter, dev)
        try:
            line['gpu_augmentation'] = gpu_augmentation_rate(patch, dev)
        except Exception as e:      # context only: never lose the bench line over it
            line['gpu_augmentation'] = {'error': repr(e)}
    if n_gpus == 1 and not args.no_cpu_baseline:
        # (i) the survey's CPU case (BASELINE.md section 4 / SURVEY.md 8d): cfg-1 = 5-stage net, 1 x 2 x 64^3, best of 3
        # after one warm-up; (ii) ONE full step of the benchmark configuration itself so that the GPU/CPU comparison is
        # on the same config (the `value`)
        t1, threads = cpu_port_step_times((64, 64, 64), 1, False, None, 3, 1)
        tfull, _ = cpu_port_step_times(patch, PER_GPU_BATCH, dual, topo_iter, 1, 0)
        line['cpu_baseline'] = {'value': PER_GPU_BATCH / tfull[0], 'unit': 'patches/s', 'cores': threads, 'kind': 'port',
                                'same_config': True,
                                'sample': f'1 full training step of the same configuration (batch {PER_GPU_BATCH}, '
                                          f'{patch[0]}x{patch[1]}x{patch[2]}) on the host cores (oracle port, fp32, no warm-up step)',
                                'seconds': tfull[0] + sum(t1),
                                'cfg1': {'what': 'BASELINE.json configs[0]: 5-stage PlainConvUNet, 1 x 2 x 64^3, DC+CE, '
                                                 'fwd+bwd+clip+SGD on CPU, best of 3 after 1 warm-up',
                                         'seconds_per_step': min(t1), 'patches_64cubed_per_s': 1.0 / min(t1)}}
    if r['per_layer'] and args.per_layer:
        rows = sorted(r['per_layer'].items(), key=lambda kv: -kv[1]['ms'])
        for (kind, tag), d in rows:
            print(f'{kind:6s} {tag:44s} {d["ms"]:8.3f} ms/step '
                  f'{d["flops"] / (d["ms"] / 1e3) / 1e12:8.1f} TFLOP/s', file=sys.stderr)
    emit(line)
    finish()


if __name__ == '__main__':
    main()
