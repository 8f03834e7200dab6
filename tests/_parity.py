"""Shared GPU-parity helpers (imported by the -m gpu tests): the block-by-block teacher-forced comparison of every
ConvDropoutNormReLU / ConvTranspose3d / head of the network against the oracle's own tensors."""
import torch

TOL = 2e-2


def rel_err(got, want):
    got, want = got.detach().double().flatten(), want.detach().double().flatten()
    return float((got - want).norm() / want.norm().clamp_min(1e-30))


def build_pair(m, oracle, cin, patch, seed=0, dev='cuda:0'):
    """oracle network (He init + non-trivial affine parameters so that dgamma / dbeta are exercised) and the CUDA
    network holding the same weights."""
    topo = oracle.topology_for_patch(patch)
    ref = oracle.build_plain_conv_unet(cin, 4, patch, seed=seed).to(dev)
    g = torch.Generator().manual_seed(seed + 100)
    for n, p in ref.named_parameters():
        if '.norm.weight' in n and 'all_modules' not in n:
            p.data.copy_((1 + 0.2 * torch.randn(p.shape, generator=g)).to(dev))
        if '.norm.bias' in n and 'all_modules' not in n:
            p.data.copy_((0.1 * torch.randn(p.shape, generator=g)).to(dev))
    net = m.PlainConvUNet(cin, num_classes=4, **topo).to(dev)
    net.load_state_dict(ref.state_dict())
    return net, ref, topo


def ds_loss(mod, n_out):
    return mod.DeepSupervisionWrapper(
        mod.DC_and_CE_loss({'batch_dice': False, 'smooth': 1e-5, 'do_bg': False, 'ddp': False}, {}, weight_ce=1,
                           weight_dice=1, ignore_label=None, dice_class=mod.MemoryEfficientSoftDiceLoss),
        mod.deep_supervision_weights(n_out))


def blockwise_teacher_forced(m, oracle, patch, cin, B, dev='cuda:0', seed=0):
    """Every block of the network is fed the ORACLE's own bf16 input and output-gradient tensors, captured with hooks
    during a bf16-autocast fwd/bwd of the whole oracle network on the GPU, and must reproduce the oracle's output,
    input gradient and parameter gradients within 2e-2 (norm-wise).  Returns (blocks checked, {name.kind: rel err})."""
    from multimodal_mvd_seg_b200 import ops
    net, ref, topo = build_pair(m, oracle, cin, patch, seed=seed, dev=dev)
    batch = oracle.make_batch(B, cin, patch, topo['strides'], kind='structured')
    data = batch['data'].to(dev)
    target = [t.to(dev) for t in batch['target']]
    rec = {}

    def hook(name):
        def f(mod, inp, out):
            r = rec.setdefault(name, {})
            r['x'] = inp[0].detach()
            r['y'] = out.detach().clone()
            out.register_hook(lambda g: r.__setitem__('gy', g.detach().clone()))
            if inp[0].requires_grad:
                inp[0].register_hook(lambda g: r.__setitem__('gx', g.detach().clone()))
        return f

    handles = []
    for name, mod in ref.named_modules():
        if name.startswith('decoder.encoder'):
            continue
        if isinstance(mod, (oracle.ConvDropoutNormReLU, torch.nn.ConvTranspose3d)) or '.seg_layers.' in name:
            handles.append(mod.register_forward_hook(hook(name)))
    with torch.autocast('cuda', dtype=torch.bfloat16):
        out_ref = ref(data)
        l_ref = ds_loss(oracle, len(out_ref))(out_ref, target)
    l_ref.backward()
    for h in handles:
        h.remove()
    del out_ref, l_ref
    ref_mods = dict(ref.named_modules())
    ours = dict(net.named_modules())
    ref_grads = {n: p.grad.detach().clone() for n, p in ref.named_parameters() if p.grad is not None}
    checked, errs_all, floors = 0, {}, {}
    # the last decoder block and its head also run FUSED in the product (InstanceNorm + LeakyReLU folded into the head,
    # csrc/norm_head.cu): keep their records for a second, joint check below
    n_dec = len(net.decoder.stages)
    last_blk = f'decoder.stages.{n_dec - 1}.convs.{len(net.decoder.stages[-1].convs) - 1}'
    last_head = f'decoder.seg_layers.{n_dec - 1}'
    fused_rec = (dict(rec[last_blk]), dict(rec[last_head])) if last_blk in rec and 'gy' in rec.get(last_head, {}) else None
    for name in list(rec):
        r = rec.pop(name)
        if 'gy' not in r:      # zero-weighted deep-supervision head: no gradient reaches it
            continue
        mod, rmod = ours[name], ref_mods[name]
        # a tensor hook reports the TOTAL gradient of a tensor; where the block's input has other consumers too
        # (stage outputs feeding both the next stage and the skip / a head and the next up-convolution) the block's
        # own input gradient cannot be isolated here -- those dgrads are covered by test_kernels_gpu.py
        multi = ('.seg_layers.' in name or '.transpconvs.' in name or
                 (name.startswith('encoder.stages.') and name.endswith('.convs.0') and not name.startswith('encoder.stages.0.')))
        if multi:
            r.pop('gx', None)
        x = ops.to_cl_view(r['x'].to(torch.bfloat16)).detach().requires_grad_('gx' in r)
        gy = ops.to_cl_view(r['gy'].to(torch.bfloat16))
        for p in mod.parameters():
            p.grad = None
        if isinstance(mod, m.ConvDropoutNormReLU):
            y = mod.forward_cl(x)
            pnames = ['conv.weight', 'norm.weight', 'norm.bias']
        elif isinstance(mod, torch.nn.ConvTranspose3d):
            y = ops.ConvTransposeFn.apply(x, mod.weight, mod.bias, tuple(mod.stride), None, None)
            pnames = ['weight', 'bias']
        else:
            y = ops.HeadFn.apply(x, mod.weight, mod.bias, None)
            pnames = ['weight', 'bias']
        y.backward(gy)
        got = {'out': ops.ncdhw_view(y).float()}
        want = {'out': r['y'].float()}
        if 'gx' in r:
            got['gx'], want['gx'] = ops.ncdhw_view(x.grad).float(), r['gx'].float()
        own, rown = dict(mod.named_parameters()), dict(rmod.named_parameters())
        for pn in pnames:
            got[pn], want[pn] = own[pn].grad.float(), ref_grads[f'{name}.{pn}'].float()
        # the same block of the ORACLE evaluated in fp32 on the identical (bf16-valued) inputs: how far the reference's
        # own bf16-autocast evaluation is from exact arithmetic for each of these tensors (its rounding-noise floor)
        for p in rmod.parameters():
            p.grad = None
        x32 = r['x'].float().detach().requires_grad_('gx' in r)
        y32 = rmod(x32)
        y32.backward(r['gy'].float())
        truth = {'out': y32.detach()}
        if 'gx' in r:
            truth['gx'] = x32.grad
        for pn in pnames:
            truth[pn] = rown[pn].grad.float()
        for k in got:
            errs_all[f'{name}.{k}'] = rel_err(got[k], want[k])
            floors[f'{name}.{k}'] = (rel_err(got[k], truth[k]), rel_err(want[k], truth[k]))
        checked += 1
        del x, gy, y, r, x32, y32, got, want, truth
    if fused_rec is not None and ops.head_fusion_ok(net.decoder.stages[-1].output_channels, ours[last_head]):
        rb, rh = fused_rec
        blk, head, rblk, rhead = ours[last_blk], ours[last_head], ref_mods[last_blk], ref_mods[last_head]
        for p in list(blk.parameters()) + list(head.parameters()):
            p.grad = None
        x = ops.to_cl_view(rb['x'].to(torch.bfloat16)).detach().requires_grad_(True)
        logits = blk.forward_cl(x, head=head)
        logits.backward(ops.to_cl_view(rh['gy'].to(torch.bfloat16)))
        got = {'out': ops.ncdhw_view(logits).float(), 'gx': ops.ncdhw_view(x.grad).float(),
               'conv.weight': blk.conv.weight.grad.float(), 'norm.weight': blk.norm.weight.grad.float(),
               'norm.bias': blk.norm.bias.grad.float(), 'head.weight': head.weight.grad.float(),
               'head.bias': head.bias.grad.float()}
        want = {'out': rh['y'].float(), 'gx': rb['gx'].float(),
                'conv.weight': ref_grads[f'{last_blk}.conv.weight'].float(), 'norm.weight': ref_grads[f'{last_blk}.norm.weight'].float(),
                'norm.bias': ref_grads[f'{last_blk}.norm.bias'].float(), 'head.weight': ref_grads[f'{last_head}.weight'].float(),
                'head.bias': ref_grads[f'{last_head}.bias'].float()}
        for p in list(rblk.parameters()) + list(rhead.parameters()):
            p.grad = None
        x32 = rb['x'].float().detach().requires_grad_(True)
        y32 = rhead(rblk(x32))
        y32.backward(rh['gy'].float())
        truth = {'out': y32.detach(), 'gx': x32.grad, 'conv.weight': rblk.conv.weight.grad.float(),
                 'norm.weight': rblk.norm.weight.grad.float(), 'norm.bias': rblk.norm.bias.grad.float(),
                 'head.weight': rhead.weight.grad.float(), 'head.bias': rhead.bias.grad.float()}
        for k in got:
            errs_all[f'fused[{last_blk}+head].{k}'] = rel_err(got[k], want[k])
            floors[f'fused[{last_blk}+head].{k}'] = (rel_err(got[k], truth[k]), rel_err(want[k], truth[k]))
        checked += 1
    return checked, errs_all, floors


def out_of_tolerance(errs, floors, tol=TOL):
    """tensors that miss the tolerance against the bf16-autocast oracle AND are further from the oracle's exact (fp32)
    evaluation of the same block on the same inputs than the bf16 oracle itself is (+10 %): i.e. genuinely worse than
    the reference's own rounding noise.  A weight gradient that sums millions of cancelling bf16-rounded products (the
    last decoder convolution in front of the rank-4 head gradient) carries a few per cent of such noise in BOTH
    implementations."""
    bad = []
    for k, e in errs.items():
        if e < tol:
            continue
        ours32, ref32 = floors[k]
        if not ours32 <= 1.1 * ref32:
            bad.append(f'{k}: {e:.4f} vs bf16 oracle; {ours32:.4f} vs fp32 oracle (bf16 oracle itself: {ref32:.4f})')
    return bad
