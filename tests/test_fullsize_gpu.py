"""Full-size checks at BASELINE.json's configuration sizes (cfg-2: 2 x 2 x 128^3; cfg-4 volume 2 x 160 x 160 x 96).

The kernel-by-kernel and block-by-block parity tests run at sizes the CPU oracle finishes in seconds.  Here the same
kernels run at the sizes the benchmark is quoted on, checked (i) directly against the oracle executed on the GPU under
bf16 autocast (it only needs PyTorch there), and (ii) through size-independent EXACT properties: scaling an operand by a
power of two scales a convolution's output by exactly that power (fp32 accumulation and bf16 rounding commute with it),
min/max morphology commutes with it, and the streaming reductions must reproduce the sums of what was stored."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF = torch.bfloat16
DEV = 'cuda:0'


def rel_err(got, want):
    got, want = got.detach().double().flatten(), want.detach().double().flatten()
    return float((got - want).norm() / want.norm().clamp_min(1e-30))


def test_cfg2_forward_and_loss_match_oracle_bf16():
    import multimodal_mvd_seg_b200 as m
    import oracle
    patch, B = (128, 128, 128), 2
    topo = oracle.topology_for_patch(patch)
    ref = oracle.build_plain_conv_unet(2, 4, patch, seed=0).to(DEV)
    net = m.PlainConvUNet(2, num_classes=4, **topo).to(DEV)
    net.load_state_dict(ref.state_dict())
    batch = oracle.make_batch(B, 2, patch, topo['strides'], kind='structured')
    data = batch['data'].to(DEV)
    target = [t.to(DEV) for t in batch['target']]
    mk = lambda mod, n: mod.DeepSupervisionWrapper(
        mod.DC_and_CE_loss({'batch_dice': False, 'smooth': 1e-5, 'do_bg': False, 'ddp': False}, {}, weight_ce=1,
                           weight_dice=1, ignore_label=None, dice_class=mod.MemoryEfficientSoftDiceLoss),
        mod.deep_supervision_weights(n))
    with torch.no_grad():
        out = net(data)
        l = mk(m, len(out))(out, target)
        with torch.autocast('cuda', dtype=BF):
            out_ref = ref(data)
        l_ref = mk(oracle, len(out_ref))([o.float() for o in out_ref], target)
        # the loss kernels on the SAME logits: fp32 reductions, 1e-5
        l_same = mk(oracle, len(out))([o.float() for o in out], target)
    assert len(out) == len(out_ref) == 5
    for a, b in zip(out, out_ref):
        assert tuple(a.shape) == tuple(b.shape)
        assert rel_err(a.float(), b.float()) < 2e-2
    assert abs(float(l) - float(l_same)) <= 1e-5 * max(1.0, abs(float(l_same)))
    assert abs(float(l) - float(l_ref)) <= 1e-3 * max(1.0, abs(float(l_ref)))
    agree = float((out[0].argmax(1) == out_ref[0].argmax(1)).float().mean())
    truth = out_ref[0].float()
    top2 = truth.topk(2, dim=1).values
    decisive = (top2[:, 0] - top2[:, 1]) > 2e-2 * float(truth.abs().max())
    assert float((out[0].argmax(1) == truth.argmax(1))[decisive].float().mean()) >= 0.999
    print(f'cfg-2 full size: loss {float(l):.5f} vs {float(l_ref):.5f}; argmax agreement with the bf16 reference {agree:.5f}')


@pytest.mark.parametrize('cin,cout,E,s', [(32, 32, 128, 1), (64, 32, 128, 1), (32, 64, 128, 2), (320, 320, 8, 1)])
def test_conv_power_of_two_scaling_is_exact(cin, cout, E, s):
    """fprop / dgrad: bit-exact; wgrad: bit-exact in the deterministic reduction mode, 1e-5 with vector atomics."""
    import multimodal_mvd_seg_b200 as m
    ops = m.ops
    B = 2
    g = torch.Generator().manual_seed(3)
    geom = ops.ConvGeom((3,) * 3, (s,) * 3, (1,) * 3)
    Eo = geom.out_size((E, E, E))[0]
    x = torch.randn((B, E, E, E, cin), generator=g).to(BF).to(DEV)
    dy = torch.randn((B, Eo, Eo, Eo, cout), generator=g).to(BF).to(DEV)
    w = (torch.randn((cout, cin, 3, 3, 3), generator=g) / np.sqrt(cin * 27)).to(DEV)
    wf, wd = ops.pack_weights(w)
    y1, y2 = torch.empty_like(dy), torch.empty_like(dy)
    ops.conv_fprop(geom, x, y1, wf)
    ops.conv_fprop(geom, x * 4, y2, wf)
    assert torch.equal(y2.float(), y1.float() * 4)
    dx1, dx2 = torch.empty_like(x), torch.empty_like(x)
    ops.conv_dgrad(geom, dx1, dy, wd)
    ops.conv_dgrad(geom, dx2, dy * 0.5, wd)
    assert torch.equal(dx2.float(), dx1.float() * 0.5)
    dw1, dw2 = torch.empty_like(w), torch.empty_like(w)
    ops.conv_wgrad(geom, x, dy, dw1)
    ops.conv_wgrad(geom, x * 2, dy, dw2)
    assert rel_err(dw2, dw1 * 2) < 1e-5
    m.lib.set_deterministic(1)
    try:
        ops.conv_wgrad(geom, x, dy, dw1)
        ops.conv_wgrad(geom, x * 2, dy, dw2)
    finally:
        m.lib.set_deterministic(0)
    assert torch.equal(dw2, dw1 * 2)
    # and against the fp32 reference on a sub-volume of the same tensors (first 16 output planes)
    d = min(Eo, 16)
    xs = x[:, :min(E, d * s + 2)].float().permute(0, 4, 1, 2, 3)
    ys = F.conv3d(xs, w.to(BF).float(), None, stride=s, padding=1)
    assert rel_err(y1[:, :d - 1].float().permute(0, 4, 1, 2, 3), ys[:, :, :d - 1]) < 1e-2


def test_instance_norm_full_size_statistics_and_roundtrip():
    """2 x 128^3 x 32: the fused conv-epilogue sums equal the sums of the stored bf16 tensor; the normalised output has
    zero mean / unit variance per (b, c) and the backward of a constant upstream gradient is ~0 (the IN null space)."""
    import multimodal_mvd_seg_b200 as m
    ops = m.ops
    lib = m.lib
    B, E, C = 2, 128, 32
    g = torch.Generator().manual_seed(4)
    x = torch.randn((B, E, E, E, C), generator=g).to(BF).to(DEV)
    w = (torch.randn((C, C, 3, 3, 3), generator=g) / np.sqrt(C * 27)).to(DEV)
    wf, _ = ops.pack_weights(w)
    geom = ops.ConvGeom((3,) * 3, (1,) * 3, (1,) * 3)
    y = torch.empty_like(x)
    stats = torch.zeros((B, C, 2), dtype=torch.float64, device=DEV)
    ops.conv_fprop(geom, x, y, wf, stats=stats)
    yd = y.double().reshape(B, -1, C)
    np.testing.assert_allclose(stats[..., 0].cpu(), yd.sum(1).cpu(), rtol=1e-5, atol=5e-2)
    np.testing.assert_allclose(stats[..., 1].cpu(), (yd * yd).sum(1).cpu(), rtol=1e-5)
    stats2 = torch.zeros_like(stats)
    st = torch.cuda.current_stream().cuda_stream
    V = E ** 3
    lib.inorm_stats(y.data_ptr(), C, B, V, C, stats2.data_ptr(), st)
    np.testing.assert_allclose(stats2.cpu(), stats.cpu(), rtol=1e-5, atol=5e-2)
    gamma, beta = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    z = torch.empty_like(y)
    lib.inorm_lrelu_fwd(y.data_ptr(), C, z.data_ptr(), C, stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), B, V, C,
                        1e-5, 1.0, st)      # slope 1: plain InstanceNorm
    zd = z.double().reshape(B, -1, C)
    assert float(zd.mean(1).abs().max()) < 2e-3 and float((zd.var(1, unbiased=False) - 1).abs().max()) < 5e-3


def test_morphology_scaling_is_exact_at_cfg4_size():
    import multimodal_mvd_seg_b200 as m
    g = torch.Generator().manual_seed(5)
    x = torch.rand((2, 1, 160, 160, 96), generator=g).to(DEV)
    for fn in (m.soft_erode, m.soft_dilate, m.soft_open):
        assert torch.equal(fn(x * 2), fn(x) * 2)
    sk = m.soft_skel(x, 3)
    assert float(sk.min()) >= 0.0 and float(sk.max()) <= 1.0 and tuple(sk.shape) == tuple(x.shape)
    assert torch.equal(m.soft_skel(x.flip(2), 3), sk.flip(2))        # mirror equivariance along D


# ----------------------------------------------------------------------------------------------------------------
# Backward parity at BASELINE.json's sizes, against the oracle run on the GPU under bf16 autocast (external check of
# dgrad / wgrad / InstanceNorm-backward / head-backward at the sizes the benchmark is quoted on)
# ----------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('patch,cin', [((128, 128, 128), 2), ((160, 160, 96), 1)], ids=['cfg2', 'cfg4'])
def test_blockwise_teacher_forced_parity_full_size(patch, cin):
    """cfg-2: 2 x 2 x 128^3 (6 stages, bottleneck 4^3); cfg-3/4 network: 2 x 1 x 160 x 160 x 96 (last stride (2,2,1),
    bottleneck 5 x 5 x 6).  Output, input gradient and all parameter gradients of all 27 blocks <= 2e-2."""
    import multimodal_mvd_seg_b200 as m
    import oracle
    from _parity import blockwise_teacher_forced, out_of_tolerance
    m.lib.reset_fallback_count()
    checked, errs, floors = blockwise_teacher_forced(m, oracle, patch, cin, 2)
    top = sorted(errs.items(), key=lambda kv: -kv[1])[:8]
    over = {k: (e, *floors[k]) for k, e in errs.items() if not e < 2e-2}
    print(f'{patch}: {checked} blocks, {len(errs)} tensors, {len(errs) - len(over)} within 2e-2 of the bf16 oracle; '
          'largest: ' + ', '.join(f'{k} {v:.4f}' for k, v in top))
    for k, (e, o32, r32) in over.items():
        print(f'  {k}: {e:.4f} vs bf16 oracle | ours vs fp32 oracle {o32:.4f} | bf16 oracle vs fp32 oracle {r32:.4f}')
    assert checked >= 26, checked
    assert len(over) <= 4, over             # the noise-floor clause below is for isolated ill-conditioned sums only
    bad = out_of_tolerance(errs, floors)
    assert not bad, bad
    # every layer of the benchmark configurations is covered by a tcgen05 kernel: no silent CUDA-core fallback
    assert m.lib.fallback_count() == 0, m.lib.fallback_count()


@pytest.mark.parametrize('topo_iter', [3, 10])
def test_cfg4_dual_net_losses_and_dlogits_match_oracle(topo_iter):
    """cfg-4: two 1-channel networks at 2 x 160 x 160 x 96, L(out1) + L(out2) + 0.5 KL + 1.0 clDice(iter).
    Logits of both networks <= 2e-2 against the oracle networks under bf16 autocast; then, ON IDENTICAL LOGITS, the
    total loss and its `seg` / `mutual` / `topo` terms <= 1e-5 (fp32 reductions) and d(term)/d(logits) <= 2e-2 for the
    deep-supervision loss (all active scales), the KL (both networks) and the clDice term."""
    import multimodal_mvd_seg_b200 as m
    import oracle
    patch, B, c = (160, 160, 96), 2, 2
    topo = oracle.topology_for_patch(patch)
    assert tuple(topo['strides'][-1]) == (2, 2, 1)
    refs = [oracle.build_plain_conv_unet(1, 4, patch, seed=s).to(DEV) for s in (0, 1)]
    nets = []
    for r in refs:
        n = m.PlainConvUNet(1, num_classes=4, **topo).to(DEV)
        n.load_state_dict(r.state_dict())
        nets.append(n)
    batch = oracle.make_batch(B, 2, patch, topo['strides'], kind='structured')
    data = batch['data'].to(DEV)
    target = [t.to(DEV) for t in batch['target']]
    m.lib.reset_fallback_count()
    with torch.no_grad():
        outs = [nets[i](data[:, i:i + 1]) for i in range(2)]
        with torch.autocast('cuda', dtype=BF):
            outs_ref = [refs[i](data[:, i:i + 1]) for i in range(2)]
        outs_32 = [refs[i](data[:, i:i + 1]) for i in range(2)]
    assert m.lib.fallback_count() == 0
    rep = []
    for ni, (o, orf, o32) in enumerate(zip(outs, outs_ref, outs_32)):
        assert len(o) == len(orf) == 5
        for si, (a, b, c32) in enumerate(zip(o, orf, o32)):
            assert tuple(a.shape) == tuple(b.shape)
            rep.append((ni, si, rel_err(a.float(), b.float()), rel_err(a.float(), c32), rel_err(b.float(), c32)))
    print('cfg-4 logits (net, scale, ours vs bf16 oracle, ours vs fp32 oracle, bf16 oracle vs fp32 oracle): ' +
          '; '.join(f'{ni}/{si} {e1:.4f} {e2:.4f} {e3:.4f}' for ni, si, e1, e2, e3 in rep))
    for ni, si, e1, e2, e3 in rep:
        # 2e-2 against the bf16-autocast oracle; at this size and random initialisation the reference's own bf16
        # evaluation sits 2.0-4.7e-2 from its fp32 evaluation (e3, growing with depth): where 2e-2 is missed (the
        # zero-weighted deepest scale) we must be well inside that floor, and never further from exact arithmetic
        # than the bf16 reference is
        assert e1 < 2e-2 or e1 < 0.75 * e3, (ni, si, e1, e3)
        assert e2 <= 1.1 * e3, (ni, si, e2, e3)
        if si < 2:
            assert e1 < 2e-2, (ni, si, e1)
    del outs_ref, outs_32, refs
    n_sc = len(outs[0])

    def terms(mod, L1, L2, f32):
        ds = mod.DeepSupervisionWrapper(
            mod.DC_and_CE_loss({'batch_dice': False, 'smooth': 1e-5, 'do_bg': False, 'ddp': False}, {}, weight_ce=1,
                               weight_dice=1, ignore_label=None, dice_class=mod.MemoryEfficientSoftDiceLoss),
            mod.deep_supervision_weights(n_sc))
        seg = ds(L1, target) + ds(L2, target)
        mutual = mod.distill_kl(L1[0], L2[0], 1.0)
        gt = (target[0].long() == c).float()
        if f32:
            prob = torch.softmax(L1[0], 1)[:, c:c + 1]
        else:
            prob = mod.softmax_channel(L1[0], c)
        topo_l = mod.soft_cldice(iter_=topo_iter, smooth=1.)(gt, prob)
        return seg, mutual, topo_l

    leaves = lambda f32: [[(o.detach().float() if f32 else o.detach().clone()).requires_grad_() for o in outs[i]]
                          for i in range(2)]
    ours, want = leaves(False), leaves(True)
    got_terms, want_terms = terms(m, ours[0], ours[1], False), terms(oracle, want[0], want[1], True)
    for name, g, w in zip(('seg', 'mutual', 'topo'), got_terms, want_terms):
        assert abs(float(g) - float(w)) <= 1e-5 * max(1.0, abs(float(w))), (name, float(g), float(w))
    tot_g = got_terms[0] + 0.5 * got_terms[1] + 1.0 * got_terms[2]
    tot_w = want_terms[0] + 0.5 * want_terms[1] + 1.0 * want_terms[2]
    assert abs(float(tot_g) - float(tot_w)) <= 1e-5 * max(1.0, abs(float(tot_w)))
    report = {}
    for ti, name in enumerate(('seg', 'mutual', 'topo')):
        for L in ours + want:
            for t in L:
                t.grad = None
        got_terms[ti].backward(retain_graph=True)
        want_terms[ti].backward(retain_graph=True)
        for ni in range(2):
            for si in range(n_sc):
                gw = want[ni][si].grad
                gg = ours[ni][si].grad
                if gw is None or float(gw.abs().max()) == 0.0:
                    assert gg is None or float(gg.float().abs().max()) == 0.0, (name, ni, si)
                    continue
                assert gg is not None, (name, ni, si)
                report[f'{name}.net{ni + 1}.scale{si}'] = rel_err(gg.float(), gw)
    print(f'cfg-4 iter {topo_iter}: total {float(tot_g):.6f} vs {float(tot_w):.6f}; dlogits ' +
          ', '.join(f'{k} {v:.4f}' for k, v in report.items()))
    assert {'seg.net1.scale0', 'seg.net2.scale3', 'mutual.net1.scale0', 'mutual.net2.scale0', 'topo.net1.scale0'} <= set(report)
    bad = {k: v for k, v in report.items() if not v < 2e-2}
    assert not bad, bad


def test_cfg2_argmax_agreement_after_training_full_size():
    """north_star: argmax masks identical on >= 99.9 % of voxels -- asserted RAW (no decisiveness filter) at cfg-2's
    size on weights that have been trained for a while (at random initialisation the four logits are near-tied and the
    reference itself in bf16 disagrees with its own fp32 run on ~1 % of voxels)."""
    import multimodal_mvd_seg_b200 as m
    import oracle
    dev = torch.device(DEV)
    patch = (128, 128, 128)
    plans, dj = m.make_plans(patch, batch_size=2, n_modalities=2, n_classes=4)
    tr = m.nnUNetTrainer(plans, '3d_fullres', 0, dj, device=dev)
    torch.manual_seed(0)
    tr.initialize()
    topo = oracle.topology_for_patch(patch)
    batch = oracle.make_batch(2, 2, patch, topo['strides'], kind='structured')
    lab = batch['target'][0]
    batch['data'] = batch['data'] * 0.3 + torch.cat([(lab == 2).float() * 2 + (lab == 1).float(),
                                                      (lab == 3).float() * 2 - (lab == 1).float()], 1)
    tr.on_train_epoch_start()
    first = float(tr.train_step(batch)['loss'])
    for _ in range(60):
        last = float(tr.train_step(batch)['loss'])
    assert last < first - 0.3, (first, last)
    ref = oracle.build_plain_conv_unet(2, 4, patch, seed=0).to(dev)
    ref.load_state_dict(tr.network.state_dict())
    data = batch['data'].to(dev)
    with torch.no_grad():
        out = tr.network(data)
        with torch.autocast('cuda', dtype=BF):
            out_ref = ref(data)
    assert rel_err(out[0].float(), out_ref[0].float()) < 2e-2
    agree = float((out[0].argmax(1) == out_ref[0].argmax(1)).float().mean())
    print(f'cfg-2 full size after 61 steps: loss {first:.3f} -> {last:.3f}; raw argmax agreement {agree:.5f}')
    assert agree >= 0.999, agree
