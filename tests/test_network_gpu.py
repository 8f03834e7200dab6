"""GPU parity of the assembled hot path against the oracle run on the same device under
torch.autocast('cuda', dtype=torch.bfloat16) with identical weights and inputs:
logits and gradients within 2e-2 relative (norm-wise), argmax masks identical on >= 99.9 % of voxels,
losses within 2e-2 (they are fp32 reductions of logits that themselves carry the bf16 tolerance)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 2e-2


def rel_err(got, want):
    got, want = got.double().flatten(), want.double().flatten()
    return float((got - want).norm() / want.norm().clamp_min(1e-30))


def _build_pair(m, oracle, cin, patch, seed=0, dev='cuda:0'):
    topo = oracle.topology_for_patch(patch)
    ref = oracle.build_plain_conv_unet(cin, 4, patch, seed=seed).to(dev)
    # non-trivial affine parameters so that dgamma/dbeta paths are exercised
    g = torch.Generator().manual_seed(seed + 100)
    for n, p in ref.named_parameters():
        if '.norm.weight' in n and 'all_modules' not in n:
            p.data.copy_((1 + 0.2 * torch.randn(p.shape, generator=g)).to(dev))
        if '.norm.bias' in n and 'all_modules' not in n:
            p.data.copy_((0.1 * torch.randn(p.shape, generator=g)).to(dev))
    net = m.PlainConvUNet(cin, num_classes=4, **topo).to(dev)
    net.load_state_dict(ref.state_dict())
    return net, ref, topo


def _check_param_grads(net, ref):
    named_ref = dict(ref.named_parameters())
    worst = 0.0
    for n, p in net.named_parameters():
        gr = named_ref[n].grad
        if gr is None or float(gr.abs().max()) == 0.0:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, n
            continue
        assert p.grad is not None, n
        if n.endswith('.conv.bias') and 'stages' in n:
            # a bias in front of InstanceNorm has an exactly-zero true gradient: both sides hold rounding noise.
            wn = n[:-len('bias')] + 'weight'
            scale = float(named_ref[wn].grad.abs().max())
            assert float(p.grad.abs().max()) <= 0.05 * max(scale, 1e-6) + 1e-3, n
            continue
        e = rel_err(p.grad, gr)
        worst = max(worst, e)
        assert e < TOL, f'{n}: gradient rel err {e:.4f}'
    return worst


@pytest.mark.parametrize('patch,cin,B', [((32, 32, 32), 2, 2), ((40, 40, 24), 1, 1)])
def test_unet_forward_backward_parity(patch, cin, B):
    import multimodal_mvd_seg_b200 as m
    import oracle
    dev = 'cuda:0'
    net, ref, topo = _build_pair(m, oracle, cin, patch)
    batch = oracle.make_batch(B, cin, patch, topo['strides'], kind='structured')
    data = batch['data'].to(dev)
    target = [t.to(dev) for t in batch['target']]
    out = net(data)
    with torch.autocast('cuda', dtype=torch.bfloat16):
        out_ref = ref(data)
    assert len(out) == len(out_ref)
    for a, b in zip(out, out_ref):
        assert tuple(a.shape) == tuple(b.shape) and a.dtype == torch.bfloat16
        assert rel_err(a.float(), b.float()) < TOL
    agree = float((out[0].argmax(1) == out_ref[0].argmax(1)).float().mean())
    assert agree >= 0.999, agree
    # loss + backward
    ds_w = m.deep_supervision_weights(len(out))
    mk = lambda mod: mod.DeepSupervisionWrapper(
        mod.DC_and_CE_loss({'batch_dice': False, 'smooth': 1e-5, 'do_bg': False, 'ddp': False}, {}, weight_ce=1,
                           weight_dice=1, ignore_label=None, dice_class=mod.MemoryEfficientSoftDiceLoss), ds_w)
    l = mk(m)(out, target)
    with torch.autocast('cuda', dtype=torch.bfloat16):
        l_ref = mk(oracle)(out_ref, target)
    assert abs(float(l) - float(l_ref)) <= TOL * max(1.0, abs(float(l_ref)))
    l.backward()
    l_ref.backward()
    worst = _check_param_grads(net, ref)
    print(f'patch {patch}: loss {float(l):.5f} vs {float(l_ref):.5f}, argmax agreement {agree:.5f}, '
          f'worst param-grad rel err {worst:.4f}')


def test_deep_supervision_switch_and_eval():
    import multimodal_mvd_seg_b200 as m
    import oracle
    net, ref, topo = _build_pair(m, oracle, 2, (32, 32, 32))
    x = torch.randn(1, 2, 32, 32, 32, device='cuda:0')
    net.decoder.deep_supervision = False
    ref.decoder.deep_supervision = False
    with torch.no_grad():
        y = net(x)
        with torch.autocast('cuda', dtype=torch.bfloat16):
            yr = ref(x)
    assert isinstance(y, torch.Tensor) and tuple(y.shape) == (1, 4, 32, 32, 32)
    assert rel_err(y.float(), yr.float()) < TOL


def test_trainer_step_matches_oracle_step():
    import multimodal_mvd_seg_b200 as m
    import oracle
    dev = torch.device('cuda:0')
    patch = (32, 32, 32)
    plans, dj = m.make_plans(patch, batch_size=2, n_modalities=2, n_classes=4)
    tr = m.nnUNetTrainer(plans, '3d_fullres', 0, dj, device=dev)
    tr.initialize()
    topo = oracle.topology_for_patch(patch)
    ref = oracle.build_plain_conv_unet(2, 4, patch, seed=0).to(dev)
    tr.network.load_state_dict(ref.state_dict())
    batch = oracle.make_batch(2, 2, patch, topo['strides'], kind='structured')
    p0 = [p.detach().clone() for p in ref.parameters()]
    tr.on_train_epoch_start()
    before = m.lib.launch_count()
    out = tr.train_step(batch)
    assert m.lib.launch_count() - before > 50
    assert isinstance(out['loss'], np.ndarray)
    # oracle step: autocast fwd/bwd, clip 12, SGD nesterov (nnUNetTrainer.py:906-924)
    data, target = batch['data'].to(dev), [t.to(dev) for t in batch['target']]
    l_ref, _ = oracle.single_net_step_loss(ref, data, target, autocast_bf16=True)
    l_ref.backward()
    params = list(ref.parameters())
    oracle.sgd_nesterov_clip_step([p.data for p in params], [p.grad for p in params], [None] * len(params), lr=1e-2)
    assert abs(float(out['loss']) - float(l_ref)) <= TOL * max(1.0, abs(float(l_ref)))
    num = den = 0.0
    for p_new, p_ref_new, p_old in zip(tr.network.parameters(), params, p0):
        num += float(((p_new.detach() - p_old) - (p_ref_new.detach() - p_old)).double().pow(2).sum())
        den += float((p_ref_new.detach() - p_old).double().pow(2).sum())
    assert (num / den) ** 0.5 < TOL, (num / den) ** 0.5
    # validation_step contract
    v = tr.validation_step(batch)
    assert set(v) == {'loss', 'tp_hard', 'fp_hard', 'fn_hard'} and v['tp_hard'].shape == (3,)
    # a second step runs (momentum buffers, arena reuse)
    out2 = tr.train_step(batch)
    assert np.isfinite(out2['loss'])


@pytest.mark.parametrize('vessel_only', [False, True])
def test_mvd_step_matches_oracle(vessel_only):
    import multimodal_mvd_seg_b200 as m
    import oracle
    dev = torch.device('cuda:0')
    patch = (32, 32, 32)
    plans, dj = m.make_plans(patch, batch_size=2, n_modalities=2, n_classes=4)
    tr = m.MVDTrainer(plans, '3d_fullres', 0, dj, device=dev, topo_iter=3, kl_vessel_only=vessel_only)
    tr.initialize()
    topo = oracle.topology_for_patch(patch)
    r1 = oracle.build_plain_conv_unet(1, 4, patch, seed=0).to(dev)
    r2 = oracle.build_plain_conv_unet(1, 4, patch, seed=1).to(dev)
    tr.network.load_state_dict(r1.state_dict())
    tr.network2.load_state_dict(r2.state_dict())
    batch = oracle.make_batch(2, 2, patch, topo['strides'], kind='structured')
    data, target = batch['data'].to(dev), [t.to(dev) for t in batch['target']]
    l, _ = tr._forward_loss(data, target)
    l.backward()
    l_ref, parts = oracle.mvd_step_loss(r1, r2, data, target, lambda1=0.5, lambda3=1.0, T=1.0, topo_iter=3,
                                        kl_vessel_only=vessel_only, autocast_bf16=True)
    l_ref.backward()
    assert abs(float(l) - float(l_ref)) <= TOL * max(1.0, abs(float(l_ref)))
    assert abs(float(tr.last_terms['mutual']) - float(parts['mutual'])) <= TOL * max(abs(float(parts['mutual'])), 1e-2)
    _check_param_grads(tr.network, r1)
    _check_param_grads(tr.network2, r2)


def test_checkpoint_roundtrip(tmp_path):
    import multimodal_mvd_seg_b200 as m
    dev = torch.device('cuda:0')
    plans, dj = m.make_plans((16, 16, 16), batch_size=1)
    tr = m.nnUNetTrainer(plans, '3d_fullres', 0, dj, device=dev)
    tr.initialize()
    f = str(tmp_path / 'checkpoint_final.pth')
    tr.save_checkpoint(f)
    ck = torch.load(f, weights_only=False)
    assert {'network_weights', 'optimizer_state', 'grad_scaler_state', 'logging', '_best_ema', 'current_epoch',
            'init_args', 'trainer_name', 'inference_allowed_mirroring_axes'} <= set(ck)
    ck['network_weights'] = {'module.' + k: v for k, v in ck['network_weights'].items()}   # a DDP-saved checkpoint
    tr2 = m.nnUNetTrainer(plans, '3d_fullres', 0, dj, device=dev)
    tr2.load_checkpoint(ck)
    for a, b in zip(tr.network.parameters(), tr2.network.parameters()):
        assert torch.equal(a, b)
