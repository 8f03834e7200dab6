"""CPU: pins oracle/ against fixtures produced by the reference's own source files (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

import oracle


@pytest.mark.parametrize('vol', ['smooth', 'ties'])
@pytest.mark.parametrize('fn', ['soft_erode', 'soft_dilate', 'soft_open'])
def test_morphology_matches_reference(golden_skel, vol, fn):
    g = golden_skel
    x = torch.from_numpy(g[f'{vol}.{fn}.in']).clone().requires_grad_(True)
    y = getattr(oracle, fn)(x)
    assert np.array_equal(y.detach().numpy(), g[f'{vol}.{fn}.out'])
    (y * torch.from_numpy(g[f'{vol}.{fn}.w'])).sum().backward()
    assert np.array_equal(x.grad.numpy(), g[f'{vol}.{fn}.grad'])


@pytest.mark.parametrize('vol', ['smooth', 'ties'])
@pytest.mark.parametrize('it', [0, 1, 3])
def test_soft_skel_matches_reference(golden_skel, vol, it):
    g = golden_skel
    x = torch.from_numpy(g[f'{vol}.soft_erode.in']).clone().requires_grad_(True)
    y = oracle.soft_skel(x, it)
    assert np.array_equal(y.detach().numpy(), g[f'{vol}.soft_skel{it}.out'])
    (y * torch.from_numpy(g[f'{vol}.soft_skel{it}.w'])).sum().backward()
    np.testing.assert_allclose(x.grad.numpy(), g[f'{vol}.soft_skel{it}.grad'], rtol=0, atol=1e-6)


def test_robust_ce(golden_misc):
    g = golden_misc
    x = torch.from_numpy(g['ce.logits']).clone().requires_grad_(True)
    l = oracle.RobustCrossEntropyLoss()(x, torch.from_numpy(g['ce.target']))
    l.backward()
    np.testing.assert_allclose(l.detach().numpy(), g['ce.loss'], rtol=1e-6)
    np.testing.assert_allclose(x.grad.numpy(), g['ce.grad'], rtol=1e-5, atol=1e-8)


@pytest.mark.parametrize('tag', ['c4_T1', 'c4_T2', 'c1_T1'])
def test_distill_kl(golden_misc, tag):
    g = golden_misc
    a = torch.from_numpy(g[f'kl.{tag}.ys']).clone().requires_grad_(True)
    b = torch.from_numpy(g[f'kl.{tag}.yt']).clone().requires_grad_(True)
    l = oracle.distill_kl(a, b, float(g[f'kl.{tag}.T']))
    l.backward()
    np.testing.assert_allclose(l.detach().numpy(), g[f'kl.{tag}.loss'], rtol=1e-6)
    np.testing.assert_allclose(a.grad.numpy(), g[f'kl.{tag}.gs'], rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(b.grad.numpy(), g[f'kl.{tag}.gt'], rtol=1e-5, atol=1e-9)


def test_polylr(golden_misc):
    g = golden_misc
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1e-2)
    sch = oracle.PolyLRScheduler(opt, 1e-2, 1000)
    for e, lr in zip(g['polylr.epochs'], g['polylr.lrs']):
        sch.step(int(e))
        assert opt.param_groups[0]['lr'] == pytest.approx(float(lr), rel=1e-12)


def test_he_init(golden_misc):
    g = golden_misc
    torch.manual_seed(0)
    conv = torch.nn.Conv3d(3, 5, 3)
    tconv = torch.nn.ConvTranspose3d(5, 3, 2, 2)
    torch.nn.Sequential(conv, tconv).apply(oracle.InitWeights_He(1e-2))
    assert np.array_equal(conv.weight.detach().numpy(), g['he.conv_w'])
    assert np.array_equal(conv.bias.detach().numpy(), g['he.conv_b'])
    assert np.array_equal(tconv.weight.detach().numpy(), g['he.tconv_w'])


def test_sum_tensor(golden_misc):
    g = golden_misc
    assert np.array_equal(oracle.sum_tensor(torch.from_numpy(g['sum.in']), (0, 2, 3)).numpy(), g['sum.out'])


@pytest.mark.parametrize('tag', ['128', '64', '160', '32'])
def test_topology(golden_misc, tag):
    g = golden_misc
    topo = oracle.topology_for_patch(tuple(int(i) for i in g[f'topo.{tag}.patch']))
    assert np.array_equal(np.array(topo['strides']), g[f'topo.{tag}.pool'])
    assert np.array_equal(np.array(topo['kernel_sizes']), g[f'topo.{tag}.convk'])
    assert topo['features_per_stage'] == [min(32 * 2 ** i, 320) for i in range(len(topo['strides']))]


def test_network_shapes_and_keys():
    net = oracle.build_plain_conv_unet(2, 4, (32, 32, 32))
    keys = list(net.state_dict().keys())
    for k in ('encoder.stages.0.0.convs.0.conv.weight', 'encoder.stages.0.0.convs.0.all_modules.1.bias',
              'decoder.encoder.stages.1.0.convs.1.norm.weight', 'decoder.transpconvs.0.weight',
              'decoder.stages.0.convs.0.conv.weight', 'decoder.seg_layers.2.bias'):
        assert k in keys
    out = net(torch.zeros(1, 2, 32, 32, 32))
    assert [tuple(o.shape) for o in out] == [(1, 4, 32, 32, 32), (1, 4, 16, 16, 16), (1, 4, 8, 8, 8)]
    # parameter count of the cfg-2 network
    topo = oracle.topology_for_patch((128, 128, 128))
    n = oracle.PlainConvUNet(2, num_classes=4, **topo)
    assert sum(p.numel() for p in n.parameters()) == 31198068


def test_ds_weights_and_step_helpers():
    np.testing.assert_allclose(oracle.deep_supervision_weights(5), [8 / 15, 4 / 15, 2 / 15, 1 / 15, 0])
    # sgd_nesterov_clip_step reproduces clip_grad_norm_ + torch.optim.SGD(nesterov)
    torch.manual_seed(1)
    p_ref = [torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(7))]
    p_mine = [p.detach().clone() for p in p_ref]
    opt = torch.optim.SGD(p_ref, 1e-2, weight_decay=3e-5, momentum=0.99, nesterov=True)
    bufs = [None, None]
    for it in range(3):
        grads = [torch.randn_like(p) * 30 for p in p_ref]
        for p, g in zip(p_ref, grads):
            p.grad = g.clone()
        torch.nn.utils.clip_grad_norm_(p_ref, 12)
        opt.step()
        _, bufs = oracle.sgd_nesterov_clip_step(p_mine, grads, bufs, 1e-2)
        for a, b in zip(p_ref, p_mine):
            np.testing.assert_allclose(a.detach().numpy(), b.numpy(), rtol=1e-5, atol=1e-7)


# ------------------------------------------------------------------------------------------------------------------
# deep-supervision target transform (oracle/ds_targets.py): the restated third-party resampling (scipy.ndimage.zoom,
# what skimage.transform.resize(order=0) calls) against the closed form the CUDA kernel implements
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('shape,new', [((16, 16, 12), (8, 8, 6)), ((20, 20, 12), (10, 10, 12)), ((9, 7, 5), (4, 4, 2)),
                                       ((12, 10, 6), (3, 2, 2)), ((5, 6, 7), (5, 3, 7))])
def test_ds_resize_segmentation_closed_form(shape, new):
    from oracle import ds_targets
    rng = np.random.default_rng(7)
    seg = rng.integers(-1, 4, size=shape).astype(np.int16)
    a = ds_targets.resize_segmentation(seg, new, 0)
    b = ds_targets.resize_segmentation_closed_form(seg, new)
    assert a.dtype == seg.dtype and np.array_equal(a, b)
    if all(o == 2 * n for o, n in zip(shape, new)):     # factor-2 pyramid: the odd voxel of every pair
        assert np.array_equal(a, seg[1::2, 1::2, 1::2])


def test_ds_transform_oracle_contract():
    from oracle import ds_targets
    rng = np.random.default_rng(8)
    seg = rng.integers(0, 4, size=(2, 1, 16, 16, 12)).astype(np.float32)
    scales = [[1, 1, 1], [0.5, 0.5, 0.5], [0.25, 0.25, 0.5]]
    out = ds_targets.DownsampleSegForDSTransform2(scales, 0, input_key='target', output_key='target')(target=seg)['target']
    assert out[0] is seg                                           # all-ones scale: the input object itself (:43-44)
    assert out[1].shape == (2, 1, 8, 8, 6) and out[2].shape == (2, 1, 4, 4, 6)
    assert np.array_equal(out[1], seg[:, :, 1::2, 1::2, 1::2])
