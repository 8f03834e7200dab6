"""Loss modules with the reference's call signatures, computed by libmvdseg kernels.

  DeepSupervisionWrapper(loss, weight_factors)(list_out, list_tgt)       nnUNetTrainer.py:366-374
  DC_and_CE_loss(soft_dice_kwargs, ce_kwargs, weight_ce, weight_dice, ignore_label, dice_class)(net_output, target)
                                                                         nnUNetTrainer.py:359-361
  MemoryEfficientSoftDiceLoss(apply_nonlin, batch_dice, do_bg, smooth, ddp)(x, y, loss_mask=None)
  RobustCrossEntropyLoss(weight, ignore_index)(input, target)           training/loss/robust_ce_loss.py:6-16
  get_tp_fp_fn_tn(net_output, gt, axes, mask)                           nnUNetTrainer.py:990 (hard one-hot input)
  distill_kl(y_s, y_t, T=1)                                             training/loss/other_loss.py:51-64
  soft_erode / soft_dilate / soft_open / soft_skel(img, iter_)          training/loss/soft_skeleton.py:6-37
  soft_cldice(iter_, smooth)(y_true, y_pred)                            (clDice wrapper)
Inputs are the network's bf16 logits of logical shape [B,C,D,H,W] (channels-last views are consumed in place) and
the reference's float targets [B,1,D,H,W].  The unsupported corners of the reference API raise NotImplementedError
instead of silently falling back to PyTorch: class weights, a stand-alone loss_mask, region (BCE) training
(ignore_label is built: DC_and_CE_loss(ignore_label=...) with the label behind the last class, as nnU-Net defines it).
"""
from typing import Callable, List, Optional, Sequence

import numpy as np
import torch
from torch import nn

from . import ops


def softmax_helper_dim1(x: torch.Tensor) -> torch.Tensor:
    """utilities/helpers.py:8-9 -- marker object: the fused kernels apply the softmax themselves."""
    return torch.softmax(x, 1)


def _dice_ce(logits, target, *, w_ce, w_dice, smooth, do_bg, batch_dice, ddp, weights=None, networks=None,
             ignore_label=None):
    """logits / target: one tensor, or one list over the deep-supervision scales.  ``networks``: a list of such logits
    lists (several networks supervised by the same targets): everything goes through one autograd node."""
    if networks is None:
        single = not isinstance(logits, (list, tuple))
        if single:
            logits, target = [logits], [target]
        networks = [list(logits)]
    target = list(target)
    n = len(target)
    assert all(len(net) == n for net in networks), 'every network needs one output per target scale'
    if weights is None:
        weights = [1.0] * n
    cfg = dict(weights=[float(w) for w in weights], weight_ce=float(w_ce), weight_dice=float(w_dice),
               smooth=float(smooth), do_bg=bool(do_bg), batch_dice=bool(batch_dice), ddp=bool(ddp),
               n_nets=len(networks), ignore_label=ignore_label)
    return ops.DiceCEMultiScaleFn.apply(cfg, *[t for net in networks for t in net], *target)


class MemoryEfficientSoftDiceLoss(nn.Module):
    def __init__(self, apply_nonlin: Callable = None, batch_dice: bool = False, do_bg: bool = True,
                 smooth: float = 1., ddp: bool = True):
        super().__init__()
        self.do_bg, self.batch_dice, self.apply_nonlin, self.smooth, self.ddp = do_bg, batch_dice, apply_nonlin, smooth, ddp

    def forward(self, x, y, loss_mask=None):
        if loss_mask is not None:
            raise NotImplementedError('loss_mask (ignore label) is outside the built hot path')
        if self.apply_nonlin is not softmax_helper_dim1:
            raise NotImplementedError('the fused Dice kernel applies softmax over dim 1 (apply_nonlin=softmax_helper_dim1)')
        return _dice_ce(x, y, w_ce=0.0, w_dice=1.0, smooth=self.smooth, do_bg=self.do_bg, batch_dice=self.batch_dice,
                        ddp=self.ddp)


class RobustCrossEntropyLoss(nn.Module):
    def __init__(self, weight=None, ignore_index: int = -100, **kw):
        super().__init__()
        if weight is not None or kw.get('label_smoothing', 0):
            raise NotImplementedError('class weights / label smoothing are outside the built hot path')
        self.ignore_index = ignore_index

    def forward(self, input: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        if target.dim() == input.dim():
            assert target.shape[1] == 1
            target = target[:, 0]
        return _dice_ce(input, target, w_ce=1.0, w_dice=0.0, smooth=1.0, do_bg=True, batch_dice=False, ddp=False)


class DC_and_CE_loss(nn.Module):
    def __init__(self, soft_dice_kwargs, ce_kwargs, weight_ce=1, weight_dice=1, ignore_label=None,
                 dice_class=MemoryEfficientSoftDiceLoss):
        super().__init__()
        # ignore_label (nnUNetTrainer.py:353-361 passes label_manager.ignore_label): voxels carrying it are masked out of
        # the Dice sums (loss_mask) and of the cross-entropy mean (ignore_index) inside the fused kernels, which treat
        # every target outside [0, C) that way -- nnU-Net's ignore label is the id behind the last class
        self.weight_dice, self.weight_ce, self.ignore_label = weight_dice, weight_ce, ignore_label
        self.ce = RobustCrossEntropyLoss(**ce_kwargs)
        self.dc = dice_class(apply_nonlin=softmax_helper_dim1, **soft_dice_kwargs)

    def _kw(self):
        return dict(w_ce=self.weight_ce, w_dice=self.weight_dice, smooth=self.dc.smooth, do_bg=self.dc.do_bg,
                    batch_dice=self.dc.batch_dice, ddp=self.dc.ddp, ignore_label=self.ignore_label)

    def forward(self, net_output: torch.Tensor, target: torch.Tensor):
        return _dice_ce(net_output, target, **self._kw())


class DeepSupervisionWrapper(nn.Module):
    def __init__(self, loss, weight_factors=None):
        super().__init__()
        self.weight_factors = weight_factors
        self.loss = loss

    def forward(self, *args):
        for i in args:
            assert isinstance(i, (tuple, list)), 'all args must be either tuple or list, got %s' % type(i)
        weights = [1] * len(args[0]) if self.weight_factors is None else self.weight_factors
        if isinstance(self.loss, DC_and_CE_loss) and len(args) == 2:
            # all scales through one autograd node; zero-weight scales are skipped (they contribute 0)
            return _dice_ce(list(args[0]), list(args[1]), weights=weights, **self.loss._kw())
        l = weights[0] * self.loss(*[j[0] for j in args])
        for i, inputs in enumerate(zip(*args)):
            if i == 0:
                continue
            if weights[i] != 0:
                l = l + weights[i] * self.loss(*inputs)
        return l

    def forward_networks(self, outputs_per_network, targets):
        """sum over several networks of ``self(outputs_k, targets)`` -- ``loss(out1, tgt) + loss(out2, tgt)`` of the
        mutual-distillation step (MVDTrainer.py:925) -- in one fused forward and one fused backward launch."""
        weights = [1] * len(targets) if self.weight_factors is None else self.weight_factors
        if isinstance(self.loss, DC_and_CE_loss):
            return _dice_ce(None, list(targets), weights=weights, networks=[list(o) for o in outputs_per_network],
                            **self.loss._kw())
        total = None
        for o in outputs_per_network:
            l = self(o, targets)
            total = l if total is None else total + l
        return total


def deep_supervision_weights(n_scales: int) -> np.ndarray:
    """nnUNetTrainer.py:366-372."""
    w = np.array([1 / (2 ** i) for i in range(n_scales)])
    w[-1] = 0
    return w / w.sum()


def get_tp_fp_fn_tn(net_output, gt, axes=None, mask=None, square=False):
    """validation path only (nnUNetTrainer.py:990): net_output is the hard one-hot of argmax, so tp/fp/fn are counts.
    ``net_output`` may be the raw logits: the kernel takes the argmax itself."""
    if mask is not None or square:
        raise NotImplementedError('mask / square are outside the built hot path')
    tp, fp, fn = ops.argmax_tp_fp_fn(net_output, gt)
    total = float(np.prod([net_output.shape[0], *net_output.shape[2:]]))
    tn = total - tp - fp - fn
    return tp, fp, fn, tn


def distill_kl(y_s, y_t, T=1, upstream_grad=None):
    """other_loss.py:51-64.  ``upstream_grad`` (optional, not in the reference signature): the constant the caller
    multiplies the result with before adding it to the loss (lambda1, MVDTrainer.py:925); lets the kernel produce loss
    and gradients in a single pass (ops.KLFn)."""
    return ops.KLFn.apply(y_s, y_t, float(T), upstream_grad)


def soft_erode(img):
    return ops.SoftErodeFn.apply(img)


def soft_dilate(img):
    return ops.SoftDilateFn.apply(img)


def soft_open(img):
    return soft_dilate(soft_erode(img))


def soft_skel(img, iter_):
    return ops.SoftSkelFn.apply(img, int(iter_))


class soft_cldice(nn.Module):
    def __init__(self, iter_=3, smooth=1.):
        super().__init__()
        self.iter, self.smooth = iter_, smooth

    def forward(self, y_true, y_pred):
        return ops.SoftClDiceFn.apply(y_true, y_pred, int(self.iter), float(self.smooth))


def softmax_channel(logits, channel: int, target=None):
    """softmax(logits, 1)[:, channel:channel+1] in fp32 (input of the topological term, MVDTrainer.py:907).  With
    ``target`` ([B,1,D,H,W] class ids) returns (prob, onehot) where onehot = (target == channel).float() -- the ground
    truth channel of MVDTrainer.py:904-905 -- produced by the same kernel."""
    if target is None:
        return ops.SoftmaxChannelFn.apply(logits, int(channel))
    return ops.SoftmaxChannelFn.apply(logits, int(channel), target)
