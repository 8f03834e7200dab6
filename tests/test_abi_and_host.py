"""CPU: the C-ABI library loads and exports every symbol include/mvdseg.h declares; host-side logic."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, 'include', 'mvdseg.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(mvd_[a-z0-9_]+)\s*\(', txt)))


def test_library_exports_every_declared_symbol():
    import multimodal_mvd_seg_b200 as m
    cdll = ctypes.CDLL(m.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(cdll, s), f'{s} declared in include/mvdseg.h but not exported'
    from multimodal_mvd_seg_b200._lib import exported_symbols
    assert sorted(exported_symbols()) == syms, 'ctypes binding table and header disagree'
    assert m.lib.version() == 100


def test_bad_arguments_return_error_not_crash():
    import multimodal_mvd_seg_b200 as m
    with pytest.raises(m.MvdError, match='bad arguments'):
        m.lib.inorm_stats(None, 8, 1, 8, 8, None, None)
    with pytest.raises(m.MvdError):
        m.lib.dice_ce_fwd(None, 4, None, 1, 8, 4, None, None)


def test_product_path_refuses_cpu_tensors():
    import multimodal_mvd_seg_b200 as m
    net = m.PlainConvUNet(1, 2, [8, 16], kernel_sizes=3, strides=[1, 2], n_conv_per_stage=1, num_classes=2,
                          n_conv_per_stage_decoder=1)
    with pytest.raises(m.MvdError, match='CUDA'):
        net(torch.zeros(1, 1, 8, 8, 8))
    with pytest.raises(m.MvdError, match='CUDA'):
        m.distill_kl(torch.zeros(1, 2, 4, 4, 4), torch.zeros(1, 2, 4, 4, 4))
    with pytest.raises(RuntimeError, match='CUDA'):
        plans, dj = m.make_plans((16, 16, 16))
        m.nnUNetTrainer(plans, '3d_fullres', 0, dj, device=torch.device('cpu'))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, 'multimodal_mvd_seg_b200')
    for fn in os.listdir(pkg):
        if fn.endswith('.py'):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r'^\s*(import|from)\s+oracle', src, flags=re.M), fn


def test_state_dict_keys_match_oracle_network():
    import multimodal_mvd_seg_b200 as m
    import oracle
    topo = oracle.topology_for_patch((32, 32, 32))
    a = oracle.PlainConvUNet(2, num_classes=4, **topo)
    b = m.PlainConvUNet(2, num_classes=4, **topo)
    sa, sb = a.state_dict(), b.state_dict()
    assert list(sa.keys()) == list(sb.keys())
    assert all(sa[k].shape == sb[k].shape for k in sa)
    b.load_state_dict(sa)
    assert b.compute_conv_feature_map_size((32, 32, 32)) == a.compute_conv_feature_map_size((32, 32, 32))


def test_make_plans_matches_reference_topology(golden_misc):
    import multimodal_mvd_seg_b200 as m
    for tag in ('128', '64', '160', '32'):
        patch = tuple(int(i) for i in golden_misc[f'topo.{tag}.patch'])
        plans, _ = m.make_plans(patch)
        cfg = plans['configurations']['3d_fullres']
        assert np.array_equal(np.array(cfg['pool_op_kernel_sizes']), golden_misc[f'topo.{tag}.pool'])
        assert np.array_equal(np.array(cfg['conv_kernel_sizes']), golden_misc[f'topo.{tag}.convk'])


def test_split_batch_for_rank():
    import multimodal_mvd_seg_b200 as m
    # global batch 16 over 8 ranks, oversample 0.33 (MVDTrainer.py:316-361)
    got = [m.split_batch_for_rank(16, 8, r, 0.33) for r in range(8)]
    assert [g[0] for g in got] == [2] * 8
    assert [g[1] for g in got][:5] == [0.0] * 5 and got[7][1] == 1.0 and 0 < got[5][1] < 1
    # uneven split: 5 over 2 -> 3, 2
    assert [m.split_batch_for_rank(5, 2, r)[0] for r in range(2)] == [3, 2]
    with pytest.raises(AssertionError):
        m.split_batch_for_rank(1, 2, 0)


def test_polylr_and_ds_weights(golden_misc):
    import multimodal_mvd_seg_b200 as m
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1e-2)
    sch = m.PolyLRScheduler(opt, 1e-2, 1000)
    for e, lr in zip(golden_misc['polylr.epochs'], golden_misc['polylr.lrs']):
        sch.step(int(e))
        assert opt.param_groups[0]['lr'] == pytest.approx(float(lr), rel=1e-12)
    np.testing.assert_allclose(m.deep_supervision_weights(5), [8 / 15, 4 / 15, 2 / 15, 1 / 15, 0])


def test_he_init_product(golden_misc):
    import multimodal_mvd_seg_b200 as m
    torch.manual_seed(0)
    conv = torch.nn.Conv3d(3, 5, 3)
    tconv = torch.nn.ConvTranspose3d(5, 3, 2, 2)
    torch.nn.Sequential(conv, tconv).apply(m.InitWeights_He(1e-2))
    assert np.array_equal(conv.weight.detach().numpy(), golden_misc['he.conv_w'])
    assert np.array_equal(tconv.weight.detach().numpy(), golden_misc['he.tconv_w'])


def test_ds_targets_host_logic_and_cpu_refusal():
    """shape rule of DownsampleSegForDSTransform2 (deep_supervision_donwsampling.py:46-49: float shape x scale, np.round
    = half to even) shared with the oracle, argument errors through the C ABI, and no CPU path."""
    import multimodal_mvd_seg_b200 as m
    from multimodal_mvd_seg_b200 import ds_targets as prod
    from oracle import ds_targets as orc
    for shape, s in [((2, 1, 16, 16, 12), [0.5, 0.5, 0.5]), ((1, 1, 5, 7, 9), [0.5, 0.5, 0.5]),
                     ((1, 2, 20, 20, 12), [0.25, 0.25, 1]), ((1, 1, 10, 6, 3), [0.25, 0.25, 0.5])]:
        seg = np.zeros(shape, dtype=np.int16)
        want = orc.DownsampleSegForDSTransform2([s], 0)(seg=seg)['seg'][0].shape
        assert prod._new_shape(shape, [2, 3, 4], s) == tuple(want)
    assert prod._new_shape((1, 1, 5, 7, 9), [2, 3, 4], [0.5] * 3) == (1, 1, 2, 4, 4)      # 2.5 -> 2, 3.5 -> 4, 4.5 -> 4
    with pytest.raises(RuntimeError, match='CUDA'):
        m.downsample_seg_for_ds(torch.zeros(1, 1, 4, 4, 4), [[0.5, 0.5, 0.5]])
    with pytest.raises(NotImplementedError):
        m.DownsampleSegForDSTransform2([[1, 1, 1]], order=3)
    with pytest.raises(m.MvdError, match='bad arguments'):
        m.lib.downsample_seg_nearest(None, 1, 4, 4, 4, 1, None, None, None)
