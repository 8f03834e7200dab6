"""Per-kernel time table of one eager training step (CUPTI via torch.profiler; no ncu replay, so it costs seconds).
Durations are in-step (warm caches, real predecessors), summed per kernel name over STEPS steps.
usage: step_profile.py [cfg2|cfg3|cfg4] [steps=3]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import multimodal_mvd_seg_b200 as m
import oracle
from bench import WORKLOADS, PER_GPU_BATCH, N_CLASSES

wl = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
patch, dual, topo_iter = WORKLOADS[wl]
dev = torch.device('cuda:0')
plans, dj = m.make_plans(patch, batch_size=PER_GPU_BATCH, n_modalities=2, n_classes=N_CLASSES)
tr = (m.MVDTrainer(plans, '3d_fullres', 0, dj, device=dev, topo_iter=topo_iter) if dual
      else m.nnUNetTrainer(plans, '3d_fullres', 0, dj, device=dev))
torch.manual_seed(0)
tr.initialize()
strides = plans['configurations']['3d_fullres']['pool_op_kernel_sizes']
host = oracle.make_batch(PER_GPU_BATCH, 2, patch, strides, max_label=N_CLASSES - 1, seed=1234, kind='rand')
res = {'data': host['data'].to(dev), 'target': [t.to(dev) for t in host['target']]}
tr.on_train_epoch_start()
for _ in range(3):
    tr.train_step_async(res)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(steps):
        tr.train_step_async(res)
    torch.cuda.synchronize()
rows = {}
for e in prof.events():
    if e.device_type.name != 'CUDA':
        continue
    d = rows.setdefault(e.name, [0.0, 0])
    d[0] += e.device_time if hasattr(e, 'device_time') else e.cuda_time
    d[1] += 1
tot = sum(v[0] for v in rows.values())
print(f'{wl}: {tot / steps / 1e3:.3f} ms of kernel time per step, {sum(v[1] for v in rows.values()) / steps:.0f} launches/step')
for name, (us, n) in sorted(rows.items(), key=lambda kv: -kv[1][0]):
    print(f'{us / steps / 1e3:9.3f} ms {100 * us / tot:5.1f}%  {n / steps:6.1f}x  {name[:110]}')
