"""Soak run: many graph-replayed training steps of one workload on random batches (rare races / hangs / NaNs show up here,
not in the short tests).  usage: soak.py [cfg2|cfg3|cfg4] [steps=1000]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_mvd_seg_b200 as m
import oracle
from bench import WORKLOADS, PER_GPU_BATCH, N_CLASSES

wl = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
patch, dual, topo_iter = WORKLOADS[wl]
dev = torch.device('cuda:0')
plans, dj = m.make_plans(patch, batch_size=PER_GPU_BATCH, n_modalities=2, n_classes=N_CLASSES)
tr = (m.MVDTrainer(plans, '3d_fullres', 0, dj, device=dev, topo_iter=topo_iter) if dual
      else m.nnUNetTrainer(plans, '3d_fullres', 0, dj, device=dev))
torch.manual_seed(0)
tr.initialize()
tr.use_cuda_graph = True
strides = plans['configurations']['3d_fullres']['pool_op_kernel_sizes']
batches = []
for s in range(4):
    host = oracle.make_batch(PER_GPU_BATCH, 2, patch, strides, max_label=N_CLASSES - 1, seed=100 + s, kind='structured')
    batches.append({'data': host['data'].to(dev), 'target': [t.to(dev) for t in host['target']]})
tr.on_train_epoch_start()
t0 = time.time()
losses = []
for i in range(steps):
    l = tr.train_step_async(batches[i % 4])
    if i % 50 == 0 or i == steps - 1:
        v = float(l)
        losses.append(v)
        assert v == v and abs(v) < 1e4, (i, v)
torch.cuda.synchronize()
for n in ([tr.network] + ([tr.network2] if dual else [])):
    for k, p in n.named_parameters():
        assert bool(torch.isfinite(p).all()), k
print(f'{wl}: {steps} steps in {time.time() - t0:.1f} s, loss {losses[0]:.4f} -> {losses[-1]:.4f} (min {min(losses):.4f}), all parameters finite')
