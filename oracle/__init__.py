"""oracle/ -- TEST INFRASTRUCTURE ONLY.

A plain-PyTorch restatement of the reference hot path (nnU-Net v2 3d_fullres training step of
JaronTu/Multimodal_MVD_Seg: PlainConvUNet + deep-supervision DC_and_CE + mutual-distillation KL +
soft-skeleton clDice).  It is the checker the CUDA path is compared against; it is never the thing
measured or shipped.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``multimodal_mvd_seg_b200/`` imports this package.

Parity pinning:
  * PINNED against the reference's own source, imported from /root/reference in the build container and
    frozen as fixtures under tests/golden/ (generator: tests/golden/make_golden.py):
      soft_erode / soft_dilate / soft_open / soft_skel   (training/loss/soft_skeleton.py:6-37)
      RobustCrossEntropyLoss                             (training/loss/robust_ce_loss.py:6-16)
      PolyLRScheduler                                    (training/lr_scheduler/polylr.py:4-20)
      InitWeights_He                                     (utilities/network_initialization.py:4-12)
      sum_tensor, softmax_helper_dim1                    (utilities/tensor_utilities.py:7-15, helpers.py:8-9)
      get_pool_and_conv_props                            (experiment_planning/.../network_topology.py:30-105)
      distill_kl formula                                 (training/loss/other_loss.py:51-64; the module itself
                                                          cannot be imported -- `lightly` is absent -- so the
                                                          six arithmetic lines are executed from the file text)
  * PARITY UNPINNED (the module is absent from the reference tree AND its third-party home is not installed;
    restated from the call sites cited in each docstring):
      PlainConvUNet / PlainConvEncoder / UNetDecoder / StackedConvBlocks (dynamic_network_architectures>=0.2,
      un-vendored, setup.py:15), MemoryEfficientSoftDiceLoss, DC_and_CE_loss, DeepSupervisionWrapper,
      get_tp_fp_fn_tn (nnunetv2.training.loss.{dice,compound_losses,deep_supervision}, missing from the tree),
      soft_cldice wrapper (public clDice repo), and the canonical MVD step (the in-tree train_step references
      undefined names, MVDTrainer.py:897-898,920).
"""
from .network import PlainConvUNet, PlainConvEncoder, UNetDecoder, StackedConvBlocks, ConvDropoutNormReLU, \
    InitWeights_He, build_plain_conv_unet, topology_for_patch
from .losses import (softmax_helper_dim1, sum_tensor, get_tp_fp_fn_tn, MemoryEfficientSoftDiceLoss,
                     RobustCrossEntropyLoss, DC_and_CE_loss, DeepSupervisionWrapper, distill_kl,
                     soft_erode, soft_dilate, soft_open, soft_skel, soft_cldice, deep_supervision_weights)
from .step import mvd_step_loss, single_net_step_loss, sgd_nesterov_clip_step, PolyLRScheduler
from .synthetic import make_batch, structured_labels

__all__ = [n for n in dir() if not n.startswith('_')]
from . import inference
from . import ds_targets
