import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import multimodal_mvd_seg_b200 as m
import oracle
from bench import WORKLOADS, PER_GPU_BATCH, N_CLASSES
wl = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
patch, dual, topo_iter = WORKLOADS[wl]
patch = (64, 64, 64) if wl == 'cfg2' else (64, 64, 32)
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
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    tr.train_step_async(res)
    torch.cuda.synchronize()
for e in prof.events():
    if e.name in ('aten::fill_', 'aten::zero_', 'aten::zeros', 'aten::ones_like', 'aten::zeros_like', 'aten::add_', 'aten::add', 'aten::copy_', 'aten::mul', 'aten::sum', 'aten::cat'):
        print(e.name, [s for s in (e.stack or [])][:6], e.input_shapes if hasattr(e, 'input_shapes') else '')
