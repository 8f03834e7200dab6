"""single process: gradient of one step on a batch of 4 vs the mean of the gradients of its two halves (same weights)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_mvd_seg_b200 as m
import oracle

dev = torch.device('cuda:0')
patch = (64, 64, 64)
strides = None


def grads_for(batch, bs):
    plans, dj = m.make_plans(patch, batch_size=bs, n_modalities=2, n_classes=4)
    torch.manual_seed(0)
    tr = m.nnUNetTrainer(plans, '3d_fullres', 0, dj, device=dev)
    tr.initialize()
    data, target = tr._to_device(batch)
    out = tr._step_forward(data)
    l, _ = tr._loss(out, target)
    l.backward()
    for a in tr._arenas:
        a.finish()
        a.attach_grads()
    torch.cuda.synchronize()
    g = {n: p.grad.detach().clone() for n, p in tr.network.named_parameters() if p.grad is not None}
    w = {n: p.detach().clone() for n, p in tr.network.named_parameters()}
    return g, w, float(l)


plans, dj = m.make_plans(patch, batch_size=4)
strides = plans['configurations']['3d_fullres']['pool_op_kernel_sizes']
union = oracle.make_batch(4, 2, patch, strides, kind='structured', seed=77)
half = lambda r: {'data': union['data'][2 * r:2 * r + 2], 'target': [t[2 * r:2 * r + 2] for t in union['target']]}
gu, wu, lu = grads_for(union, 4)
g0, w0, l0 = grads_for(half(0), 2)
g1, w1, l1 = grads_for(half(1), 2)
assert all(torch.equal(wu[k], w0[k]) for k in wu)
print('losses', lu, (l0 + l1) / 2)
rows = []
num = den = 0.0
for k in gu:
    a = gu[k].double()
    b = ((g0[k] + g1[k]) / 2).double()
    e = float((a - b).norm() / a.norm().clamp_min(1e-30))
    num += float((a - b).pow(2).sum()); den += float(a.pow(2).sum())
    rows.append((e, k, float(a.norm())))
rows.sort(reverse=True)
print('total rel err', (num / den) ** 0.5)
for e, k, n in rows[:25]:
    print(f'{e:.4e}  |g|={n:.3e}  {k}')
# determinism of the same computation twice
g0b, _, _ = grads_for(half(0), 2)
print('repeat rel err', max(float((g0[k].double() - g0b[k].double()).norm() / g0[k].double().norm().clamp_min(1e-30)) for k in g0))
