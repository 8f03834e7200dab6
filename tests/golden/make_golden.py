"""Generates tests/golden/*.npz by running the REFERENCE's own source files (importable ones) from /root/reference.

Run in the build container only (the reference tree does not exist on the GPU box):
    python tests/golden/make_golden.py
The fixtures pin oracle/ (tests/test_oracle_golden.py) and, through it and directly, the CUDA kernels.
Files executed (nnUNet/nnunetv2/...):
  training/loss/soft_skeleton.py        soft_erode, soft_dilate, soft_open, soft_skel (+ autograd gradients)
  training/loss/robust_ce_loss.py       RobustCrossEntropyLoss
  training/loss/other_loss.py:51-64     distill_kl (function text exec'd: the module imports `lightly`, absent)
  training/lr_scheduler/polylr.py       PolyLRScheduler
  utilities/network_initialization.py   InitWeights_He
  utilities/tensor_utilities.py         sum_tensor
  experiment_planning/experiment_planners/network_topology.py   get_pool_and_conv_props
  inference/sliding_window_prediction.py:10-58   compute_gaussian, compute_steps_for_sliding_window (function texts
                                                 exec'd: the module imports acvl_utils, absent)  -> inference.npz
"""
import importlib.util
import os
import sys
import warnings

import numpy as np
import torch
import torch.nn.functional as F

REF = '/root/reference/nnUNet/nnunetv2'
OUT = os.path.dirname(os.path.abspath(__file__))


def load(rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def golden_inference():
    """exec the two pure functions of inference/sliding_window_prediction.py and record their outputs."""
    import re
    from functools import lru_cache
    from typing import Union, Tuple, List
    from scipy.ndimage import gaussian_filter
    src = open(os.path.join(REF, 'inference/sliding_window_prediction.py')).read()
    a = src.index('@lru_cache(maxsize=2)')
    b = src.index("if __name__ == '__main__':")
    ns = dict(np=np, torch=torch, lru_cache=lru_cache, Union=Union, Tuple=Tuple, List=List, gaussian_filter=gaussian_filter)
    exec(src[a:b], ns)
    out = {}
    for i, tile in enumerate([(8, 8, 8), (12, 10, 6), (32, 20, 16), (5, 7, 9)]):
        g = ns['compute_gaussian'](tile, sigma_scale=1. / 8, value_scaling_factor=1000, dtype=torch.float32,
                                   device=torch.device('cpu'))
        out[f'gauss{i}.tile'] = np.array(tile)
        out[f'gauss{i}.out'] = g.numpy()
    cases = [((40, 48, 37), (32, 32, 32), 0.5), ((110, 64, 70), (64, 64, 64), 0.5), ((32, 32, 32), (32, 32, 32), 0.5),
             ((100, 90, 33), (32, 48, 32), 0.25), ((65, 130, 64), (64, 64, 64), 1.0)]
    for i, (img, tile, step) in enumerate(cases):
        st = ns['compute_steps_for_sliding_window'](img, tile, step)
        out[f'steps{i}.img'] = np.array(img)
        out[f'steps{i}.tile'] = np.array(tile)
        out[f'steps{i}.step'] = np.array(step)
        for ax in range(3):
            out[f'steps{i}.ax{ax}'] = np.array(st[ax])
    np.savez_compressed(os.path.join(OUT, 'inference.npz'), **out)
    print('wrote inference.npz', len(out), 'arrays')


def main():
    golden_inference()
    torch.manual_seed(20261018)
    sk = load('training/loss/soft_skeleton.py', 'ref_soft_skeleton')
    g = torch.Generator().manual_seed(7)
    out = {}
    # two volumes: smooth random (few ties) and coarsely quantised (many ties, incl. exact 0/1 as bf16-saturated
    # probabilities produce) -- exercises the tie rules of max_pool3d / torch.min / relu backward
    vols = {
        'smooth': torch.rand((2, 1, 9, 11, 10), generator=g),
        'ties': torch.round(torch.rand((1, 1, 8, 7, 12), generator=g) * 4) / 4,
    }
    for name, v in vols.items():
        for fn_name in ('soft_erode', 'soft_dilate', 'soft_open'):
            x = v.clone().requires_grad_(True)
            y = getattr(sk, fn_name)(x)
            w = torch.rand(y.shape, generator=g)
            (y * w).sum().backward()
            out[f'{name}.{fn_name}.in'] = v.numpy()
            out[f'{name}.{fn_name}.out'] = y.detach().numpy()
            out[f'{name}.{fn_name}.w'] = w.numpy()
            out[f'{name}.{fn_name}.grad'] = x.grad.numpy()
        for it in (0, 1, 3):
            x = v.clone().requires_grad_(True)
            y = sk.soft_skel(x, it)
            w = torch.rand(y.shape, generator=g)
            (y * w).sum().backward()
            out[f'{name}.soft_skel{it}.out'] = y.detach().numpy()
            out[f'{name}.soft_skel{it}.w'] = w.numpy()
            out[f'{name}.soft_skel{it}.grad'] = x.grad.numpy()
    np.savez_compressed(os.path.join(OUT, 'soft_skeleton.npz'), **out)

    # ---- RobustCrossEntropyLoss
    ce = load('training/loss/robust_ce_loss.py', 'ref_robust_ce')
    logits = torch.randn((2, 4, 5, 6, 7), generator=g)
    target = torch.randint(0, 4, (2, 1, 5, 6, 7), generator=g).float()
    x = logits.clone().requires_grad_(True)
    l = ce.RobustCrossEntropyLoss()(x, target)
    l.backward()
    misc = {'ce.logits': logits.numpy(), 'ce.target': target.numpy(), 'ce.loss': l.detach().numpy(),
            'ce.grad': x.grad.numpy()}

    # ---- distill_kl: exec the function text (lines 51-64 of other_loss.py), dropping nothing
    src = open(os.path.join(REF, 'training/loss/other_loss.py')).read().split('\n')
    start = next(i for i, s in enumerate(src) if s.startswith('def distill_kl'))
    end = next(i for i in range(start + 1, len(src)) if src[i].startswith('def '))
    ns = {'torch': torch, 'F': F}
    exec('\n'.join(src[start:end]), ns)
    for tag, C, T in (('c4_T1', 4, 1.0), ('c4_T2', 4, 2.0), ('c1_T1', 1, 1.0)):
        ys = torch.randn((2, C, 4, 5, 6), generator=g)
        yt = torch.randn((2, C, 4, 5, 6), generator=g)
        a, b = ys.clone().requires_grad_(True), yt.clone().requires_grad_(True)
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            l = ns['distill_kl'](None, a, b, T)   # the reference signature carries a stray `self`
        l.backward()
        misc.update({f'kl.{tag}.ys': ys.numpy(), f'kl.{tag}.yt': yt.numpy(), f'kl.{tag}.T': np.float32(T),
                     f'kl.{tag}.loss': l.detach().numpy(), f'kl.{tag}.gs': a.grad.numpy(), f'kl.{tag}.gt': b.grad.numpy()})

    # ---- PolyLR
    poly = load('training/lr_scheduler/polylr.py', 'ref_polylr')
    opt = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1))], lr=1e-2)
    # the reference ctor passes a `verbose` positional that torch 2.11's LRScheduler no longer takes
    # (polylr.py:11); build the object by hand and run the reference's own step()
    sch = object.__new__(poly.PolyLRScheduler)
    sch.optimizer, sch.initial_lr, sch.max_steps, sch.exponent, sch.ctr = opt, 1e-2, 1000, 0.9, 0
    lrs = []
    for e in (0, 1, 10, 500, 999):
        sch.step(e)
        lrs.append(opt.param_groups[0]['lr'])
    misc['polylr.epochs'] = np.array([0, 1, 10, 500, 999])
    misc['polylr.lrs'] = np.array(lrs, dtype=np.float64)

    # ---- InitWeights_He
    init = load('utilities/network_initialization.py', 'ref_init')
    torch.manual_seed(0)
    conv = torch.nn.Conv3d(3, 5, 3)
    tconv = torch.nn.ConvTranspose3d(5, 3, 2, 2)
    seq = torch.nn.Sequential(conv, tconv)
    seq.apply(init.InitWeights_He(1e-2))
    misc['he.conv_w'] = conv.weight.detach().numpy()
    misc['he.conv_b'] = conv.bias.detach().numpy()
    misc['he.tconv_w'] = tconv.weight.detach().numpy()

    # ---- sum_tensor
    tu = load('utilities/tensor_utilities.py', 'ref_tu')
    t = torch.randn((2, 3, 4, 5), generator=g)
    misc['sum.in'] = t.numpy()
    misc['sum.out'] = tu.sum_tensor(t, (0, 2, 3)).numpy()

    # ---- topology
    topo = load('experiment_planning/experiment_planners/network_topology.py', 'ref_topo')
    for tag, patch in (('128', (128, 128, 128)), ('64', (64, 64, 64)), ('160', (160, 160, 96)), ('32', (32, 32, 32))):
        npa, pool, convk, ps, div = topo.get_pool_and_conv_props((1.0, 1.0, 1.0), patch, 4, 999999)
        misc[f'topo.{tag}.patch'] = np.array(patch)
        misc[f'topo.{tag}.pool'] = np.array(pool)
        misc[f'topo.{tag}.convk'] = np.array(convk)
        misc[f'topo.{tag}.num_pool'] = np.array(npa)
    np.savez_compressed(os.path.join(OUT, 'misc.npz'), **misc)
    print('wrote', os.listdir(OUT))


if __name__ == '__main__':
    if not os.path.isdir(REF):
        sys.exit('reference tree not mounted; fixtures can only be regenerated in the build container')
    main()
