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


def test_epoch_hooks_match_reference_fixture():
    """collate_outputs, nnUNetLogger.log (incl. the EMA) and the on_train_epoch_end / on_validation_epoch_end aggregation
    against tests/golden/epoch_hooks.npz, produced by executing the reference's own texts (make_golden_epoch.py).  The
    hooks only touch self.logger / is_ddp / current_epoch, so they run on a stand-in object (the trainer needs CUDA)."""
    import importlib.util
    import multimodal_mvd_seg_b200 as m
    from multimodal_mvd_seg_b200 import trainer as T
    g = np.load(os.path.join(ROOT, 'tests', 'golden', 'epoch_hooks.npz'))
    spec = importlib.util.spec_from_file_location('mk_epoch', os.path.join(ROOT, 'tests', 'golden', 'make_golden_epoch.py'))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)            # only its synthetic_outputs() is used; nothing under /root/reference is read

    class Stand:
        is_ddp = False
    me = Stand()
    me.logger = T.nnUNetLogger()
    n_epochs, n_steps, n_fg = (int(v) for v in g['meta'])
    for ep in range(n_epochs):
        me.current_epoch = ep
        train, val = mk.synthetic_outputs(100 + ep, n_steps, n_fg)
        if ep == 2:
            for v in val:
                v['tp_hard'][1] = v['fp_hard'][1] = v['fn_hard'][1] = 0
        T.nnUNetTrainer.on_train_epoch_end(me, train)
        T.nnUNetTrainer.on_validation_epoch_end(me, val)
    log = me.logger.my_fantastic_logging
    for k in ('train_losses', 'val_losses', 'mean_fg_dice', 'ema_fg_dice'):
        np.testing.assert_array_equal(np.array(log[k], dtype=np.float64), g[f'log.{k}'])
    np.testing.assert_array_equal(np.array(log['dice_per_class_or_region'], dtype=np.float64),
                                  g['log.dice_per_class_or_region'])
    assert np.isnan(g['log.dice_per_class_or_region'][2, 1])          # the empty class of epoch 2 stays nan
    mixed = [{'s': 1.5, 'a': np.arange(3.0) + i, 'l': [i, i + 1]} for i in range(3)]
    c = T.collate_outputs(mixed)
    np.testing.assert_array_equal(np.array(c['s']), g['collate.s'])
    np.testing.assert_array_equal(c['a'], g['collate.a'])
    np.testing.assert_array_equal(np.array(c['l']), g['collate.l'])
    with pytest.raises(ValueError):
        T.collate_outputs([{'x': (1, 2)}])
    # logger contract: one entry per epoch, re-logging an epoch overwrites
    lg = T.nnUNetLogger()
    lg.log('mean_fg_dice', 0.5, 0)
    lg.log('mean_fg_dice', 0.7, 1)
    assert lg.my_fantastic_logging['ema_fg_dice'] == [0.5, 0.5 * 0.9 + 0.1 * 0.7]
    lg.log('mean_fg_dice', 0.9, 1)
    assert lg.my_fantastic_logging['mean_fg_dice'] == [0.5, 0.9]
    with pytest.raises(AssertionError):
        lg.log('unknown_key', 1.0, 0)


def test_run_training_control_flow(tmp_path):
    """the epoch loop (MVDTrainer.py:1323-1345) on a stub trainer: step counts, batch order through the generator handed to
    prefetching(), logger bookkeeping, checkpoint cadence (save_every, best EMA, final)."""
    from multimodal_mvd_seg_b200 import trainer as T

    class Stub(T.nnUNetTrainer):
        def __init__(self, out):
            self.is_ddp, self.local_rank, self.was_initialized = False, 0, True
            self.num_epochs, self.num_iterations_per_epoch, self.num_val_iterations_per_epoch = 3, 4, 2
            self.current_epoch, self.save_every, self._best_ema = 0, 2, None
            self.logger = T.nnUNetLogger()
            self.output_folder = out
            self.optimizer = type('O', (), {'param_groups': [{'lr': 0.01}]})()
            self.lr_scheduler = type('S', (), {'step': lambda self_, e: None})()
            self.calls, self.saved = [], []
            self.dataloader_train, self.dataloader_val = iter(range(1000)), iter(range(1000, 2000))

        def _networks(self):
            return []

        def set_deep_supervision_enabled(self, enabled):
            self.calls.append(('ds', enabled))

        def prefetching(self, batches):
            yield from batches

        def train_step(self, b):
            self.calls.append(('train', b))
            return {'loss': np.array(1.0 + 0.1 * b, dtype=np.float32)}

        def validation_step(self, b):
            self.calls.append(('val', b))
            k = float(self.current_epoch + 1)
            return {'loss': np.array(0.5, dtype=np.float32), 'tp_hard': np.array([10.0 * k, 5.0]),
                    'fp_hard': np.array([1.0, 5.0]), 'fn_hard': np.array([1.0, 5.0])}

        def save_checkpoint(self, filename):
            self.saved.append((self.current_epoch, os.path.basename(filename)))

    tr = Stub(str(tmp_path / 'run'))
    tr.run_training()
    assert tr.current_epoch == 3 and os.path.isdir(tr.output_folder)
    assert [c[1] for c in tr.calls if c[0] == 'train'] == list(range(12))
    assert [c[1] for c in tr.calls if c[0] == 'val'] == list(range(1000, 1006))
    assert tr.calls[0] == ('ds', True)
    log = tr.logger.my_fantastic_logging
    assert all(len(log[k]) == 3 for k in log)
    np.testing.assert_allclose(log['train_losses'], [1.15, 1.55, 1.95], rtol=1e-6)
    d0 = 2 * 20.0 / (2 * 20.0 + 2 + 2)
    assert abs(log['dice_per_class_or_region'][0][0] - d0) < 1e-12 and abs(log['dice_per_class_or_region'][0][1] - 0.5) < 1e-12
    # pseudo-Dice rises every epoch -> a new best each time; 'latest' at (epoch + 1) % save_every == 0 except the last epoch
    assert tr.saved == [(0, 'checkpoint_best.pth'), (1, 'checkpoint_latest.pth'), (1, 'checkpoint_best.pth'),
                        (2, 'checkpoint_best.pth'), (3, 'checkpoint_final.pth')]
    tr2 = Stub(None)
    tr2.dataloader_train = None
    with pytest.raises(RuntimeError, match='dataloader_train'):
        tr2.run_training()


def test_load_pretrained_weights_drop_in(tmp_path):
    """our PlainConvUNet under the REFERENCE's own run/load_pretrained_weights.py (executed from /root/reference when the
    tree is present: build container only) and under the package's mirror of it: everything but the '.seg_layers.' heads
    is transferred, a shape mismatch outside the heads is refused."""
    import importlib.util
    import multimodal_mvd_seg_b200 as m

    def build(seed, classes=3):
        torch.manual_seed(seed)
        net = m.PlainConvUNet(2, 3, [8, 16, 32], kernel_sizes=3, strides=[1, 2, 2], n_conv_per_stage=2, num_classes=classes,
                              n_conv_per_stage_decoder=2, deep_supervision=True)
        net.apply(m.InitWeights_He(1e-2))
        for n, p in net.named_parameters():      # non-default norm parameters and biases
            if p.dim() == 1:
                p.data.normal_(0.5, 0.1)
        return net
    src = build(0)
    f = str(tmp_path / 'pre.pth')
    torch.save({'network_weights': src.state_dict()}, f)
    loaders = [('mirror', m.load_pretrained_weights)]
    ref_file = '/root/reference/nnUNet/nnunetv2/run/load_pretrained_weights.py'
    if os.path.exists(ref_file):
        spec = importlib.util.spec_from_file_location('ref_lpw', ref_file)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        loaders.append(('reference', mod.load_pretrained_weights))
    results = {}
    for name, fn in loaders:
        dst = build(1)
        before = {k: v.clone() for k, v in dst.state_dict().items()}
        fn(dst, f)
        after = dst.state_dict()
        for k, v in after.items():
            if '.seg_layers.' in k:
                assert torch.equal(v, before[k]), (name, k)
            else:
                assert torch.equal(v, src.state_dict()[k]), (name, k)
        results[name] = {k: v.clone() for k, v in after.items()}
        # different number of classes: only the heads differ -> still loads
        fn(build(2, classes=5), f)
    if 'reference' in results:
        assert all(torch.equal(results['mirror'][k], results['reference'][k]) for k in results['mirror'])
    wide = m.PlainConvUNet(2, 3, [8, 16, 48], kernel_sizes=3, strides=[1, 2, 2], n_conv_per_stage=2, num_classes=3,
                           n_conv_per_stage_decoder=2, deep_supervision=True)
    with pytest.raises(AssertionError, match='shape'):
        m.load_pretrained_weights(wide, f)


def test_deep_supervision_scales_match_reference_text():
    """_get_deep_supervision_scales against the reference's own method text (nnUNetTrainer.py:296-302), executed on a
    stand-in when /root/reference is present; the closed form is asserted either way."""
    import ast
    import textwrap
    import multimodal_mvd_seg_b200 as m
    from multimodal_mvd_seg_b200 import trainer as T

    class Stand:
        pass
    for patch in [(128, 128, 128), (160, 160, 96), (64, 64, 64)]:
        plans, dj = m.make_plans(patch)
        me = Stand()
        me.configuration_manager = T.PlansManager(plans).get_configuration('3d_fullres')
        me.enable_deep_supervision = True
        ours = T.nnUNetTrainer._get_deep_supervision_scales(me)
        pools = np.vstack(me.configuration_manager.pool_op_kernel_sizes)
        assert len(ours) == len(pools) - 1 and ours[0] == [1.0, 1.0, 1.0]
        np.testing.assert_array_equal(np.array(ours), (1 / np.cumprod(pools, axis=0))[:-1])
        ref_file = '/root/reference/nnUNet/nnunetv2/training/nnUNetTrainer/nnUNetTrainer.py'
        if os.path.exists(ref_file):
            src = open(ref_file).read()
            node = next(n for c in ast.parse(src).body if isinstance(c, ast.ClassDef) and c.name == 'nnUNetTrainer'
                        for n in c.body if isinstance(n, ast.FunctionDef) and n.name == '_get_deep_supervision_scales')
            ns = dict(np=np)
            exec(textwrap.dedent(ast.get_source_segment(src, node)), ns)
            assert ns['_get_deep_supervision_scales'](me) == ours
            me.enable_deep_supervision = False
            assert ns['_get_deep_supervision_scales'](me) is None
        me.enable_deep_supervision = False
        assert T.nnUNetTrainer._get_deep_supervision_scales(me) is None


def test_split_batch_for_rank_matches_reference_text():
    """split_batch_for_rank against ContrastiveTrainer._set_batch_size_and_oversample (MVDTrainer.py:316-361), executed
    from the reference file with a stand-in torch.distributed (build container only)."""
    import ast
    import contextlib
    import io
    import textwrap
    import multimodal_mvd_seg_b200 as m
    ref_file = '/root/reference/nnUNet/nnunetv2/training/nnUNetTrainer/MVDTrainer.py'
    if not os.path.exists(ref_file):
        pytest.skip('reference tree not present')
    src = open(ref_file).read()
    node = next(n for c in ast.parse(src).body if isinstance(c, ast.ClassDef) and c.name == 'ContrastiveTrainer'
                for n in c.body if isinstance(n, ast.FunctionDef) and n.name == '_set_batch_size_and_oversample')

    class Dist:
        def __init__(self, world, rank):
            self.world, self.rank = world, rank

        def get_world_size(self):
            return self.world

        def get_rank(self):
            return self.rank

    class Stand:
        pass
    for global_bs in (2, 5, 8, 16, 17):
        for world in range(1, min(global_bs, 8) + 1):
            for over in (0.33, 0.0, 1.0, 0.5):
                for rank in range(world):
                    me = Stand()
                    me.is_ddp = True
                    me.oversample_foreground_percent = over
                    me.configuration_manager = Stand()
                    me.configuration_manager.batch_size = global_bs
                    ns = dict(np=np, dist=Dist(world, rank))
                    exec(textwrap.dedent(ast.get_source_segment(src, node)), ns)
                    with contextlib.redirect_stdout(io.StringIO()), np.errstate(all='ignore'):
                        ns['_set_batch_size_and_oversample'](me)
                        bs, ov = m.split_batch_for_rank(global_bs, world, rank, over)
                    assert int(bs) == int(me.batch_size), (global_bs, world, rank, over)
                    ref_ov = float(me.oversample_foreground_percent)
                    # a rank left without samples (8 over 5 ranks -> 2,2,2,2,0) gets 0/0 in the reference as well
                    assert (np.isnan(ov) and np.isnan(ref_ov)) or abs(float(ov) - ref_ov) < 1e-12, (global_bs, world, rank, over)


def test_optimizer_state_interchanges_with_torch_sgd():
    """`optimizer_state` of a checkpoint goes both ways between SGDNesterovClip and the reference's
    torch.optim.SGD(lr, weight_decay=3e-5, momentum=0.99, nesterov=True) (MVDTrainer.py:482-486, 1138, 1180)."""
    import multimodal_mvd_seg_b200 as m
    mk = lambda: [torch.nn.Parameter(torch.randn(4, 3)), torch.nn.Parameter(torch.randn(5))]
    p_ref, p_mine = mk(), mk()
    ref = torch.optim.SGD(p_ref, 1e-2, weight_decay=3e-5, momentum=0.99, nesterov=True)
    for p in p_ref:
        p.grad = torch.randn_like(p)
    ref.step()                                    # creates the momentum buffers
    mine = m.SGDNesterovClip(p_mine, 5e-3, weight_decay=3e-5, momentum=0.99, nesterov=True, max_norm=12.0)
    assert set(mine.param_groups[0]) == set(ref.param_groups[0])          # same group keys: nothing to drop or to miss
    sd_ref = ref.state_dict()
    mine.load_state_dict(sd_ref)                  # reference checkpoint -> ours
    assert mine.max_norm == 12.0 and mine.param_groups[0]['lr'] == 1e-2
    for a, b in zip(p_mine, p_ref):
        assert torch.equal(mine.state[a]['momentum_buffer'], ref.state[b]['momentum_buffer'])
    ref2 = torch.optim.SGD(mk(), 1.0, momentum=0.5)
    ref2.load_state_dict(mine.state_dict())       # ours -> reference
    ref2.param_groups[0]['params'][0].grad = torch.zeros(4, 3)
    ref2.param_groups[0]['params'][1].grad = torch.zeros(5)
    ref2.step()                                   # would raise KeyError('dampening') on a group with missing keys
    assert ref2.param_groups[0]['nesterov'] and ref2.param_groups[0]['momentum'] == 0.99
    # a round-1 checkpoint of this package carried max_norm inside the groups: still loads
    old = mine.state_dict()
    old['param_groups'][0]['max_norm'] = 12.0
    mine.load_state_dict(old)
    assert 'max_norm' not in mine.param_groups[0]
