"""Generates tests/golden/epoch_hooks.npz from the REFERENCE's own epoch-level host code (build container only).

Executed (nnUNet/nnunetv2/...):
  utilities/collate_outputs.py:6-24                       collate_outputs (module imported as is)
  training/logging/nnunet_logger.py:9-52                  nnUNetLogger.__init__ / .log (class text exec'd without the plotting
                                                          methods: the module imports matplotlib / seaborn / batchgenerators)
  training/nnUNetTrainer/MVDTrainer.py:987-997, 1065-1097 on_train_epoch_end / on_validation_epoch_end of class ContrastiveTrainer (method texts exec'd as
                                                          free functions on a stand-in `self` with is_ddp = False)
The fixture pins multimodal_mvd_seg_b200.trainer.{collate_outputs, nnUNetLogger, on_train_epoch_end,
on_validation_epoch_end} (tests/test_abi_and_host.py::test_epoch_hooks_match_reference_fixture)."""
import ast
import importlib.util
import os
import textwrap
from typing import List

import numpy as np

REF = '/root/reference/nnUNet/nnunetv2'
OUT = os.path.dirname(os.path.abspath(__file__))


def _load(rel, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _class_without(src, cls, drop):
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            node.body = [n for n in node.body if not (isinstance(n, ast.FunctionDef) and n.name in drop)]
            return ast.unparse(node)
    raise KeyError(cls)


def _method_text(src, cls, name):
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == cls:
            for n in node.body:
                if isinstance(n, ast.FunctionDef) and n.name == name:
                    return textwrap.dedent(ast.get_source_segment(src, n))
    raise KeyError(name)


def synthetic_outputs(seed, n_steps, n_fg):
    rng = np.random.default_rng(seed)
    train = [{'loss': np.array(rng.normal(1.0, 0.2), dtype=np.float32)} for _ in range(n_steps)]
    val = []
    for _ in range(n_steps):
        tp = rng.integers(0, 5000, size=n_fg).astype(np.float64)
        fp = rng.integers(0, 3000, size=n_fg).astype(np.float64)
        fn = rng.integers(0, 3000, size=n_fg).astype(np.float64)
        val.append({'loss': np.array(rng.normal(0.8, 0.1), dtype=np.float32), 'tp_hard': tp, 'fp_hard': fp, 'fn_hard': fn})
    return train, val


def main():
    collate = _load('utilities/collate_outputs.py', 'ref_collate').collate_outputs
    ns = {}
    exec(_class_without(open(os.path.join(REF, 'training/logging/nnunet_logger.py')).read(), 'nnUNetLogger',
                        {'plot_progress_png', 'get_checkpoint', 'load_checkpoint'}), ns)
    Logger = ns['nnUNetLogger']
    tsrc = open(os.path.join(REF, 'training/nnUNetTrainer/MVDTrainer.py')).read()
    fns = dict(np=np, List=List, collate_outputs=collate, dist=None)
    exec(_method_text(tsrc, 'ContrastiveTrainer', 'on_train_epoch_end'), fns)
    exec(_method_text(tsrc, 'ContrastiveTrainer', 'on_validation_epoch_end'), fns)

    class Self:
        is_ddp = False

    me = Self()
    me.logger = Logger()
    out = {}
    n_epochs, n_steps, n_fg = 4, 5, 3
    for ep in range(n_epochs):
        me.current_epoch = ep
        train, val = synthetic_outputs(100 + ep, n_steps, n_fg)
        if ep == 2:   # a class that never occurs: 0/0 -> nan, ignored by nanmean
            for v in val:
                v['tp_hard'][1] = v['fp_hard'][1] = v['fn_hard'][1] = 0
        with np.errstate(all='ignore'):
            fns['on_train_epoch_end'](me, train)
            fns['on_validation_epoch_end'](me, val)
    for k in ('train_losses', 'val_losses', 'mean_fg_dice', 'ema_fg_dice'):
        out[f'log.{k}'] = np.array(me.logger.my_fantastic_logging[k], dtype=np.float64)
    out['log.dice_per_class_or_region'] = np.array(me.logger.my_fantastic_logging['dice_per_class_or_region'], dtype=np.float64)
    out['meta'] = np.array([n_epochs, n_steps, n_fg])
    # collate_outputs on the three value kinds it supports
    mixed = [{'s': 1.5, 'a': np.arange(3.0) + i, 'l': [i, i + 1]} for i in range(3)]
    c = collate(mixed)
    out['collate.s'] = np.array(c['s'])
    out['collate.a'] = c['a']
    out['collate.l'] = np.array(c['l'])
    np.savez_compressed(os.path.join(OUT, 'epoch_hooks.npz'), **out)
    print('wrote epoch_hooks.npz:', {k: v.shape for k, v in out.items()})


if __name__ == '__main__':
    main()
