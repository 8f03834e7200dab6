"""GPU diagnostic (not a test): logits / gradient agreement between our two conv paths and the oracle in bf16-autocast
and in fp32, to separate implementation error from the bf16 noise floor."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_mvd_seg_b200 as m
import oracle

def rel(a, b):
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))

dev = 'cuda:0'
patch = tuple(int(v) for v in (sys.argv[1].split(',') if len(sys.argv) > 1 else (32, 32, 32)))
cin, B = 2, 2
topo = oracle.topology_for_patch(patch)
ref = oracle.build_plain_conv_unet(cin, 4, patch, seed=0).to(dev)
net = m.PlainConvUNet(cin, num_classes=4, **topo).to(dev)
net.load_state_dict(ref.state_dict())
batch = oracle.make_batch(B, cin, patch, topo['strides'], kind='structured')
data = batch['data'].to(dev); target = [t.to(dev) for t in batch['target']]
w = m.deep_supervision_weights(len(topo['strides']) - 1)
mk = lambda mod: mod.DeepSupervisionWrapper(mod.DC_and_CE_loss({'batch_dice': False, 'smooth': 1e-5, 'do_bg': False, 'ddp': False}, {}, weight_ce=1, weight_dice=1, ignore_label=None, dice_class=mod.MemoryEfficientSoftDiceLoss), w)

res = {}
for algo in ('generic', 'auto'):
    m.ops.set_conv_algo(algo)
    net.zero_grad(set_to_none=True)
    out = net(data); l = mk(m)(out, target); l.backward()
    res[algo] = ([o.detach().float() for o in out], {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}, float(l))
m.ops.set_conv_algo('auto')
for name, ac in (('ref_bf16', True), ('ref_fp32', False)):
    ref.zero_grad(set_to_none=True)
    if ac:
        with torch.autocast('cuda', dtype=torch.bfloat16):
            out = ref(data); l = mk(oracle)(out, target)
    else:
        out = ref(data); l = mk(oracle)(out, target)
    l.backward()
    res[name] = ([o.detach().float() for o in out], {n: p.grad.clone() for n, p in ref.named_parameters() if p.grad is not None}, float(l))

names = list(res)
print('losses', {k: v[2] for k, v in res.items()})
for i in range(len(names)):
    for j in range(i + 1, len(names)):
        a, b = res[names[i]], res[names[j]]
        lr = [rel(x, y) for x, y in zip(a[0], b[0])]
        ag = float((a[0][0].argmax(1) == b[0][0].argmax(1)).float().mean())
        gr = {n: rel(a[1][n], b[1][n]) for n in a[1] if n in b[1] and not (n.endswith('.conv.bias') and 'stages' in n)}
        worst = sorted(gr.items(), key=lambda kv: -kv[1])[:4]
        print(f'{names[i]:9s} vs {names[j]:9s}: logits rel {["%.4f" % v for v in lr]} argmax agree {ag:.5f} '
              f'grad rel median {sorted(gr.values())[len(gr)//2]:.4f} worst {[(n[-40:], "%.4f" % v) for n, v in worst]}')
# margin-aware agreement vs fp32 truth
truth = res['ref_fp32'][0][0]
top2 = truth.topk(2, dim=1).values
margin = (top2[:, 0] - top2[:, 1])
for k in names[:-1]:
    pred = res[k][0][0].argmax(1)
    for thr in (0.0, 0.01, 0.02, 0.05):
        msk = margin > thr
        print(f'{k}: margin>{thr}: frac voxels {float(msk.float().mean()):.4f} agree {float((pred == truth.argmax(1))[msk].float().mean()):.5f}')
