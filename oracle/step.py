"""oracle/step.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

The canonical training-step arithmetic the CUDA path is checked against.
  * spine                 nnunetv2/training/nnUNetTrainer/nnUNetTrainer.py:888-925 (base train_step)
  * MVD loss assembly     nnunetv2/training/nnUNetTrainer/MVDTrainer.py:895-925 (intent; the in-tree code references
                          undefined logits1/logits2 and self.topo_loss -- SURVEY.md section 0 fact 2 -- so the canonical
                          decisions of SURVEY.md section 8c are used: two PlainConvUNet(1 channel), KL on hi-res logits,
                          soft-clDice on softmax channel 2, total = L(out1)+L(out2)+lambda3*topo+lambda1*KL)
  * optimiser             MVDTrainer.py:482-486 (SGD lr 1e-2, wd 3e-5, momentum 0.99, nesterov), clip 12 (:978),
                          PolyLR polylr.py:4-20
  * AllGatherGrad         nnunetv2/utilities/ddp_allgather.py:25-48
"""
from typing import Any, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from .losses import (DC_and_CE_loss, DeepSupervisionWrapper, MemoryEfficientSoftDiceLoss, deep_supervision_weights,
                     distill_kl, soft_cldice)


class AllGatherGrad(torch.autograd.Function):
    """ddp_allgather.py:25-48."""

    @staticmethod
    def forward(ctx: Any, tensor: torch.Tensor, group=None) -> torch.Tensor:
        ctx.group = group
        gathered = [torch.zeros_like(tensor) for _ in range(torch.distributed.get_world_size())]
        torch.distributed.all_gather(gathered, tensor, group=group)
        return torch.stack(gathered, dim=0)

    @staticmethod
    def backward(ctx: Any, *grad_output: torch.Tensor):
        grad_output = torch.cat(grad_output)
        torch.distributed.all_reduce(grad_output, op=torch.distributed.ReduceOp.SUM, async_op=False, group=ctx.group)
        return grad_output[torch.distributed.get_rank()], None


def build_ds_loss(n_scales: int, batch_dice: bool = False, ddp: bool = False) -> nn.Module:
    """nnUNetTrainer._build_loss, non-region branch (nnUNetTrainer.py:359-374)."""
    loss = DC_and_CE_loss({'batch_dice': batch_dice, 'smooth': 1e-5, 'do_bg': False, 'ddp': ddp}, {},
                          weight_ce=1, weight_dice=1, ignore_label=None, dice_class=MemoryEfficientSoftDiceLoss)
    return DeepSupervisionWrapper(loss, deep_supervision_weights(n_scales))


def single_net_step_loss(net: nn.Module, data: torch.Tensor, target: List[torch.Tensor], autocast_bf16: bool = False):
    """output = network(data); l = loss(output, target)   (nnUNetTrainer.py:906-913)."""
    ctx = torch.autocast(data.device.type, dtype=torch.bfloat16) if autocast_bf16 else _null()
    with ctx:
        output = net(data)
        l = build_ds_loss(len(output))(output, target)
    return l, output


def mvd_step_loss(net1: nn.Module, net2: nn.Module, data: torch.Tensor, target: List[torch.Tensor],
                  lambda1: float = 0.5, lambda3: float = 1.0, T: float = 1.0, topo_iter: Optional[int] = 3,
                  vessel_class: int = 2, kl_vessel_only: bool = False, autocast_bf16: bool = False):
    """Canonical MVD loss (SURVEY.md 8c decisions 1-4).  net1 sees modality 0, net2 modality 1
    (split pattern selfattnNet.py:588-589); lambdas MVDTrainer.py:132-134."""
    ctx = torch.autocast(data.device.type, dtype=torch.bfloat16) if autocast_bf16 else _null()
    with ctx:
        out1 = net1(data[:, 0:1])
        out2 = net2(data[:, 1:2])
        ds = build_ds_loss(len(out1))
        l_seg = ds(out1, target) + ds(out2, target)
        if kl_vessel_only:   # call-site literal MVDTrainer.py:897-899 (channel 2 -> shape[1]==1 branch of distill_kl)
            mutual = distill_kl(out1[0][:, vessel_class:vessel_class + 1].float(),
                                out2[0][:, vessel_class:vessel_class + 1].float(), T)
        else:
            mutual = distill_kl(out1[0].float(), out2[0].float(), T)
        total = l_seg + lambda1 * mutual
        topo = None
        if topo_iter is not None:
            prob = torch.softmax(out1[0].float(), 1)[:, vessel_class:vessel_class + 1]
            gt = (target[0].long() == vessel_class).float()      # one-hot channel 2 (MVDTrainer.py:904-908)
            topo = soft_cldice(iter_=topo_iter, smooth=1.)(gt, prob)
            total = total + lambda3 * topo
    return total, dict(seg=l_seg, mutual=mutual, topo=topo, out1=out1, out2=out2)


class _null:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


class PolyLRScheduler:
    """polylr.py:4-20 without the _LRScheduler base (only .step(current_step) is used by the trainer,
    MVDTrainer.py:869-877)."""

    def __init__(self, optimizer, initial_lr: float, max_steps: int, exponent: float = 0.9):
        self.optimizer = optimizer
        self.initial_lr = initial_lr
        self.max_steps = max_steps
        self.exponent = exponent
        self.ctr = 0

    def step(self, current_step=None):
        if current_step is None or current_step == -1:
            current_step = self.ctr
            self.ctr += 1
        new_lr = self.initial_lr * (1 - current_step / self.max_steps) ** self.exponent
        for param_group in self.optimizer.param_groups:
            param_group['lr'] = new_lr


def sgd_nesterov_clip_step(params: Sequence[torch.Tensor], grads: Sequence[torch.Tensor],
                           bufs: Sequence[Optional[torch.Tensor]], lr: float, weight_decay: float = 3e-5,
                           momentum: float = 0.99, max_norm: float = 12.0) -> Tuple[float, list]:
    """clip_grad_norm_(params, 12) followed by torch.optim.SGD(nesterov=True) -- MVDTrainer.py:978-979, 482-484.
    Pure-tensor restatement of torch's own formulas (clip coef = max_norm/(norm+1e-6) clamped to 1; first step
    buf = g).  Returns (total_norm, new_bufs); params are updated in place."""
    total_norm = torch.linalg.vector_norm(torch.stack([torch.linalg.vector_norm(g.float(), 2) for g in grads]), 2)
    coef = torch.clamp(max_norm / (total_norm + 1e-6), max=1.0)
    new_bufs = []
    for p, g, b in zip(params, grads, bufs):
        g = g.float() * coef
        g = g + weight_decay * p
        if b is None:
            b = g.clone()
        else:
            b = momentum * b + g
        g = g + momentum * b
        p.sub_(lr * g)
        new_bufs.append(b)
    return float(total_norm), new_bufs
