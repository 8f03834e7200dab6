"""oracle/losses.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Restatement of the loss arithmetic on the hot path.  In-tree sources (pinned by tests/golden fixtures):
  soft_erode/soft_dilate/soft_open/soft_skel  nnunetv2/training/loss/soft_skeleton.py:6-37
  distill_kl                                  nnunetv2/training/loss/other_loss.py:51-64
  RobustCrossEntropyLoss                      nnunetv2/training/loss/robust_ce_loss.py:6-16
  sum_tensor                                  nnunetv2/utilities/tensor_utilities.py:7-15
  softmax_helper_dim1                         nnunetv2/utilities/helpers.py:8-9
Missing from the tree (PARITY UNPINNED, restated from the call sites nnUNetTrainer.py:351-375, :990 and public
nnunetv2 2.1.x -- SURVEY.md Appendix A.2):
  get_tp_fp_fn_tn, MemoryEfficientSoftDiceLoss, DC_and_CE_loss, DeepSupervisionWrapper, soft_cldice.
"""
import warnings
from typing import Callable, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn, Tensor


def softmax_helper_dim1(x: Tensor) -> Tensor:
    return torch.softmax(x, 1)


def sum_tensor(inp: Tensor, axes, keepdim: bool = False) -> Tensor:
    axes = np.unique(axes).astype(int)
    if keepdim:
        for ax in axes:
            inp = inp.sum(int(ax), keepdim=True)
    else:
        for ax in sorted(axes, reverse=True):
            inp = inp.sum(int(ax))
    return inp


def get_tp_fp_fn_tn(net_output: Tensor, gt: Tensor, axes=None, mask=None, square: bool = False):
    """Call site nnUNetTrainer.py:990 (axes=[0,2,3,4]); body per SURVEY.md A.2."""
    if axes is None:
        axes = tuple(range(2, net_output.ndim))
    with torch.no_grad():
        if net_output.ndim != gt.ndim:
            gt = gt.view((gt.shape[0], 1, *gt.shape[1:]))
        if net_output.shape == gt.shape:
            y_onehot = gt
        else:
            y_onehot = torch.zeros(net_output.shape, device=net_output.device)
            y_onehot.scatter_(1, gt.long(), 1)
    tp = net_output * y_onehot
    fp = net_output * (1 - y_onehot)
    fn = (1 - net_output) * y_onehot
    tn = (1 - net_output) * (1 - y_onehot)
    if mask is not None:
        with torch.no_grad():
            mask_here = torch.tile(mask, (1, tp.shape[1], *[1 for _ in range(2, tp.ndim)]))
        tp, fp, fn, tn = tp * mask_here, fp * mask_here, fn * mask_here, tn * mask_here
    if square:
        tp, fp, fn, tn = tp ** 2, fp ** 2, fn ** 2, tn ** 2
    if len(axes) > 0:
        tp = sum_tensor(tp, axes, keepdim=False)
        fp = sum_tensor(fp, axes, keepdim=False)
        fn = sum_tensor(fn, axes, keepdim=False)
        tn = sum_tensor(tn, axes, keepdim=False)
    return tp, fp, fn, tn


class MemoryEfficientSoftDiceLoss(nn.Module):
    """kwargs at the call site nnUNetTrainer.py:359-360: batch_dice from plans, smooth=1e-5, do_bg=False, ddp."""

    def __init__(self, apply_nonlin: Callable = None, batch_dice: bool = False, do_bg: bool = True,
                 smooth: float = 1., ddp: bool = True):
        super().__init__()
        self.do_bg = do_bg
        self.batch_dice = batch_dice
        self.apply_nonlin = apply_nonlin
        self.smooth = smooth
        self.ddp = ddp

    def forward(self, x, y, loss_mask=None):
        if self.apply_nonlin is not None:
            x = self.apply_nonlin(x)
        axes = tuple(range(2, x.ndim))
        with torch.no_grad():
            if x.ndim != y.ndim:
                y = y.view((y.shape[0], 1, *y.shape[1:]))
            if x.shape == y.shape:
                y_onehot = y
            else:
                y_onehot = torch.zeros(x.shape, device=x.device, dtype=torch.bool)
                y_onehot.scatter_(1, y.long(), 1)
            if not self.do_bg:
                y_onehot = y_onehot[:, 1:]
            sum_gt = y_onehot.sum(axes) if loss_mask is None else (y_onehot * loss_mask).sum(axes)
        if not self.do_bg:
            x = x[:, 1:]
        if loss_mask is None:
            intersect = (x * y_onehot).sum(axes)
            sum_pred = x.sum(axes)
        else:
            intersect = (x * y_onehot * loss_mask).sum(axes)
            sum_pred = (x * loss_mask).sum(axes)
        if self.batch_dice:
            if self.ddp and torch.distributed.is_available() and torch.distributed.is_initialized():
                from .step import AllGatherGrad
                intersect = AllGatherGrad.apply(intersect).sum(0)
                sum_pred = AllGatherGrad.apply(sum_pred).sum(0)
                sum_gt = AllGatherGrad.apply(sum_gt).sum(0)
            intersect = intersect.sum(0)
            sum_pred = sum_pred.sum(0)
            sum_gt = sum_gt.sum(0)
        dc = (2 * intersect + self.smooth) / (torch.clip(sum_gt + sum_pred + self.smooth, 1e-8))
        return -dc.mean()


class RobustCrossEntropyLoss(nn.CrossEntropyLoss):
    """robust_ce_loss.py:6-16."""

    def forward(self, input: Tensor, target: Tensor) -> Tensor:
        if len(target.shape) == len(input.shape):
            assert target.shape[1] == 1
            target = target[:, 0]
        return super().forward(input, target.long())


class DC_and_CE_loss(nn.Module):
    """ctor call nnUNetTrainer.py:359-361."""

    def __init__(self, soft_dice_kwargs, ce_kwargs, weight_ce=1, weight_dice=1, ignore_label=None,
                 dice_class=MemoryEfficientSoftDiceLoss):
        super().__init__()
        if ignore_label is not None:
            ce_kwargs['ignore_index'] = ignore_label
        self.weight_dice = weight_dice
        self.weight_ce = weight_ce
        self.ignore_label = ignore_label
        self.ce = RobustCrossEntropyLoss(**ce_kwargs)
        self.dc = dice_class(apply_nonlin=softmax_helper_dim1, **soft_dice_kwargs)

    def forward(self, net_output: Tensor, target: Tensor):
        if self.ignore_label is not None:
            assert target.shape[1] == 1
            mask = (target != self.ignore_label).bool()
            target_dice = torch.clone(target)
            target_dice[target == self.ignore_label] = 0
            num_fg = mask.sum()
        else:
            target_dice = target
            mask = None
        dc_loss = self.dc(net_output, target_dice, loss_mask=mask) if self.weight_dice != 0 else 0
        ce_loss = self.ce(net_output, target[:, 0].long()) \
            if self.weight_ce != 0 and (self.ignore_label is None or num_fg > 0) else 0
        return self.weight_ce * ce_loss + self.weight_dice * dc_loss


class DeepSupervisionWrapper(nn.Module):
    """built at nnUNetTrainer.py:366-374."""

    def __init__(self, loss, weight_factors=None):
        super().__init__()
        self.weight_factors = weight_factors
        self.loss = loss

    def forward(self, *args):
        for i in args:
            assert isinstance(i, (tuple, list))
        weights = [1] * len(args[0]) if self.weight_factors is None else self.weight_factors
        l = weights[0] * self.loss(*[j[0] for j in args])
        for i, inputs in enumerate(zip(*args)):
            if i == 0:
                continue
            l += weights[i] * self.loss(*inputs)
        return l


def deep_supervision_weights(n_scales: int) -> np.ndarray:
    """nnUNetTrainer.py:366-372: 1/2^i, last = 0, normalised to sum 1."""
    w = np.array([1 / (2 ** i) for i in range(n_scales)])
    w[-1] = 0
    return w / w.sum()


def distill_kl(y_s: Tensor, y_t: Tensor, T=1):
    """other_loss.py:51-64 without the stray ``self`` first parameter."""
    if y_s.shape[1] == 1:
        y_s = torch.cat([y_s, torch.zeros_like(y_s)], 1)
        y_t = torch.cat([y_t, torch.zeros_like(y_t)], 1)
    p_s = F.log_softmax(y_s / T + 1e-40, dim=1)
    p_t = F.softmax(y_t / T, dim=1)
    # reduction='mean' divides by numel and makes torch emit a UserWarning; kept verbatim
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        loss = F.kl_div(p_s, p_t, reduction='mean') * (T ** 2)
    return loss


def soft_erode(img):
    """soft_skeleton.py:6-15 (5-D branch)."""
    p1 = -F.max_pool3d(-img, (3, 1, 1), (1, 1, 1), (1, 0, 0))
    p2 = -F.max_pool3d(-img, (1, 3, 1), (1, 1, 1), (0, 1, 0))
    p3 = -F.max_pool3d(-img, (1, 1, 3), (1, 1, 1), (0, 0, 1))
    return torch.min(torch.min(p1, p2), p3)


def soft_dilate(img):
    """soft_skeleton.py:18-22."""
    return F.max_pool3d(img, (3, 3, 3), (1, 1, 1), (1, 1, 1))


def soft_open(img):
    """soft_skeleton.py:25-26."""
    return soft_dilate(soft_erode(img))


def soft_skel(img, iter_):
    """soft_skeleton.py:29-37."""
    img1 = soft_open(img)
    skel = F.relu(img - img1)
    for _ in range(iter_):
        img = soft_erode(img)
        img1 = soft_open(img)
        delta = F.relu(img - img1)
        skel = skel + F.relu(delta - skel * delta)
    return skel


class soft_cldice(nn.Module):
    """clDice wrapper around the in-tree soft_skel (wrapper not in the tree; formula SURVEY.md A.2; hard-metric
    twin nnunetv2/training/metrics/clDice_metric.py:20-36)."""

    def __init__(self, iter_=3, smooth=1.):
        super().__init__()
        self.iter = iter_
        self.smooth = smooth

    def forward(self, y_true, y_pred):
        skel_pred = soft_skel(y_pred, self.iter)
        skel_true = soft_skel(y_true, self.iter)
        tprec = (torch.sum(skel_pred * y_true) + self.smooth) / (torch.sum(skel_pred) + self.smooth)
        tsens = (torch.sum(skel_true * y_pred) + self.smooth) / (torch.sum(skel_true) + self.smooth)
        return 1. - 2.0 * (tprec * tsens) / (tprec + tsens)
