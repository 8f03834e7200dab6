"""multimodal_mvd_seg_b200 -- B200-native implementation of the nnU-Net v2 3d_fullres training step of
JaronTu/Multimodal_MVD_Seg (PlainConvUNet fwd/bwd, deep-supervision Dice+CE, mutual-distillation KL, soft-skeleton
clDice, clip + SGD-nesterov, batch-sharded data parallelism).  Host side: Python/PyTorch plumbing; device side:
libmvdseg.so, hand-written CUDA for sm_100a behind the C ABI of include/mvdseg.h.  No CPU fallback."""
from ._lib import LIB_PATH, MvdError, lib
from . import ops
from .network import (PlainConvUNet, PlainConvEncoder, UNetDecoder, StackedConvBlocks, ConvDropoutNormReLU,
                      InitWeights_He, get_network_from_plans, load_pretrained_weights)
from .losses import (DC_and_CE_loss, DeepSupervisionWrapper, MemoryEfficientSoftDiceLoss, RobustCrossEntropyLoss,
                     get_tp_fp_fn_tn, distill_kl, soft_erode, soft_dilate, soft_open, soft_skel, soft_cldice,
                     softmax_channel, softmax_helper_dim1, deep_supervision_weights)
from .optim import SGDNesterovClip, PolyLRScheduler
from .ddp import GradArena, split_batch_for_rank
from .trainer import nnUNetTrainer, MVDTrainer, make_plans, PlansManager, ConfigurationManager, LabelManager

__version__ = '0.1.0'
from .inference import SlidingWindowPredictor, compute_gaussian, compute_steps_for_sliding_window
from .ds_targets import DownsampleSegForDSTransform2, downsample_seg_for_ds
