"""oracle/network.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Plain-PyTorch restatement of the network the reference builds for 3d_fullres:
``get_network_from_plans`` -> ``PlainConvUNet``  (reference: nnunetv2/utilities/get_network_from_plans.py:15-92).
The class itself lives in the un-vendored third-party package ``dynamic_network_architectures`` (>=0.2,
reference setup.py:15) => PARITY UNPINNED; the wiring below follows the in-tree evidence:
  * decoder construction and forward loop   nnunetv2/training/my_network/UNetDecoder.py:35-74, 104-121
  * encoder ctor argument order              nnunetv2/training/my_network/selfattnNet.py:509-514
  * hyper-parameters (bias, InstanceNorm eps/affine, LeakyReLU inplace, He init)
                                             nnunetv2/utilities/get_network_from_plans.py:38-45, 70-92
  * topology rule                            nnunetv2/experiment_planning/experiment_planners/network_topology.py:30-105
Module / state_dict names follow upstream so checkpoints interchange.
"""
from copy import deepcopy
from typing import List, Sequence, Tuple, Union

import numpy as np
import torch
from torch import nn


class InitWeights_He(object):
    """nnunetv2/utilities/network_initialization.py:4-12 (kaiming_normal_(a=neg_slope), bias 0)."""

    def __init__(self, neg_slope: float = 1e-2):
        self.neg_slope = neg_slope

    def __call__(self, module):
        if isinstance(module, (nn.Conv3d, nn.Conv2d, nn.ConvTranspose2d, nn.ConvTranspose3d)):
            module.weight = nn.init.kaiming_normal_(module.weight, a=self.neg_slope)
            if module.bias is not None:
                module.bias = nn.init.constant_(module.bias, 0)


def _triple(v) -> Tuple[int, int, int]:
    if isinstance(v, (int, np.integer)):
        return (int(v),) * 3
    v = tuple(int(i) for i in v)
    assert len(v) == 3
    return v


class ConvDropoutNormReLU(nn.Module):
    """conv -> (dropout: None here) -> InstanceNorm3d(eps=1e-5, affine) -> LeakyReLU(0.01, inplace).
    kwargs as passed at get_network_from_plans.py:39-45."""

    def __init__(self, input_channels: int, output_channels: int, kernel_size, stride, conv_bias: bool = True):
        super().__init__()
        kernel_size = _triple(kernel_size)
        stride = _triple(stride)
        self.input_channels = input_channels
        self.output_channels = output_channels
        self.stride = stride
        self.conv = nn.Conv3d(input_channels, output_channels, kernel_size, stride,
                              padding=[(k - 1) // 2 for k in kernel_size], dilation=1, bias=conv_bias)
        self.norm = nn.InstanceNorm3d(output_channels, eps=1e-5, affine=True)
        self.nonlin = nn.LeakyReLU(inplace=True)
        # upstream keeps an aliasing Sequential; it duplicates the keys in state_dict()
        self.all_modules = nn.Sequential(self.conv, self.norm, self.nonlin)

    def forward(self, x):
        return self.all_modules(x)

    def compute_conv_feature_map_size(self, input_size):
        output_size = [i // j for i, j in zip(input_size, self.stride)]
        return int(np.prod([self.output_channels, *output_size], dtype=np.int64))


class StackedConvBlocks(nn.Module):
    def __init__(self, num_convs: int, input_channels: int, output_channels, kernel_size, initial_stride,
                 conv_bias: bool = True):
        super().__init__()
        if not isinstance(output_channels, (tuple, list)):
            output_channels = [output_channels] * num_convs
        self.convs = nn.Sequential(
            ConvDropoutNormReLU(input_channels, output_channels[0], kernel_size, initial_stride, conv_bias),
            *[ConvDropoutNormReLU(output_channels[i - 1], output_channels[i], kernel_size, 1, conv_bias)
              for i in range(1, num_convs)])
        self.output_channels = output_channels[-1]
        self.initial_stride = _triple(initial_stride)

    def forward(self, x):
        return self.convs(x)

    def compute_conv_feature_map_size(self, input_size):
        output = self.convs[0].compute_conv_feature_map_size(input_size)
        size_after_stride = [i // j for i, j in zip(input_size, self.initial_stride)]
        for b in self.convs[1:]:
            output += b.compute_conv_feature_map_size(size_after_stride)
        return output


class PlainConvEncoder(nn.Module):
    """pool='conv': the first conv of each stage carries the stride."""

    def __init__(self, input_channels: int, n_stages: int, features_per_stage, kernel_sizes, strides,
                 n_conv_per_stage, conv_bias: bool = True, return_skips: bool = True):
        super().__init__()
        if isinstance(kernel_sizes, int):
            kernel_sizes = [kernel_sizes] * n_stages
        if isinstance(features_per_stage, int):
            features_per_stage = [features_per_stage] * n_stages
        if isinstance(n_conv_per_stage, int):
            n_conv_per_stage = [n_conv_per_stage] * n_stages
        if isinstance(strides, int):
            strides = [strides] * n_stages
        assert len(kernel_sizes) == len(features_per_stage) == len(n_conv_per_stage) == len(strides) == n_stages
        stages = []
        cin = input_channels
        for s in range(n_stages):
            stages.append(nn.Sequential(
                StackedConvBlocks(n_conv_per_stage[s], cin, features_per_stage[s], kernel_sizes[s], strides[s],
                                  conv_bias)))
            cin = features_per_stage[s]
        self.stages = nn.Sequential(*stages)
        self.output_channels = list(features_per_stage)
        self.strides = [_triple(i) for i in strides]
        self.return_skips = return_skips
        self.conv_op = nn.Conv3d
        self.norm_op = nn.InstanceNorm3d
        self.norm_op_kwargs = {'eps': 1e-5, 'affine': True}
        self.nonlin = nn.LeakyReLU
        self.nonlin_kwargs = {'inplace': True}
        self.dropout_op = None
        self.dropout_op_kwargs = None
        self.conv_bias = conv_bias
        self.kernel_sizes = [_triple(k) for k in kernel_sizes]

    def forward(self, x):
        ret = []
        for s in self.stages:
            x = s(x)
            ret.append(x)
        return ret if self.return_skips else ret[-1]

    def compute_conv_feature_map_size(self, input_size):
        output = 0
        for s in range(len(self.stages)):
            output += self.stages[s][-1].compute_conv_feature_map_size(input_size)
            input_size = [i // j for i, j in zip(input_size, self.strides[s])]
        return output


class UNetDecoder(nn.Module):
    """UNetDecoder.py:35-74 (ctor) and :104-121 (forward), minus the fork's attention lines :75-102."""

    def __init__(self, encoder: PlainConvEncoder, num_classes: int, n_conv_per_stage, deep_supervision: bool):
        super().__init__()
        self.deep_supervision = deep_supervision
        self.encoder = encoder
        self.num_classes = num_classes
        n_stages_encoder = len(encoder.output_channels)
        if isinstance(n_conv_per_stage, int):
            n_conv_per_stage = [n_conv_per_stage] * (n_stages_encoder - 1)
        assert len(n_conv_per_stage) == n_stages_encoder - 1
        stages, transpconvs, seg_layers = [], [], []
        for s in range(1, n_stages_encoder):
            below = encoder.output_channels[-s]
            skip = encoder.output_channels[-(s + 1)]
            st = encoder.strides[-s]
            transpconvs.append(nn.ConvTranspose3d(below, skip, st, st, bias=encoder.conv_bias))
            stages.append(StackedConvBlocks(n_conv_per_stage[s - 1], 2 * skip, skip, encoder.kernel_sizes[-(s + 1)], 1,
                                            encoder.conv_bias))
            seg_layers.append(nn.Conv3d(skip, num_classes, 1, 1, 0, bias=True))
        self.stages = nn.ModuleList(stages)
        self.transpconvs = nn.ModuleList(transpconvs)
        self.seg_layers = nn.ModuleList(seg_layers)

    def forward(self, skips):
        lres_input = skips[-1]
        seg_outputs = []
        for s in range(len(self.stages)):
            x = self.transpconvs[s](lres_input)
            x = torch.cat((x, skips[-(s + 2)]), 1)
            x = self.stages[s](x)
            if self.deep_supervision:
                seg_outputs.append(self.seg_layers[s](x))
            elif s == (len(self.stages) - 1):
                seg_outputs.append(self.seg_layers[-1](x))
            lres_input = x
        seg_outputs = seg_outputs[::-1]
        return seg_outputs if self.deep_supervision else seg_outputs[0]

    def compute_conv_feature_map_size(self, input_size):
        skip_sizes = []
        for s in range(len(self.encoder.strides) - 1):
            skip_sizes.append([i // j for i, j in zip(input_size, self.encoder.strides[s])])
            input_size = skip_sizes[-1]
        output = 0
        for s in range(len(self.stages)):
            output += self.stages[s].compute_conv_feature_map_size(skip_sizes[-(s + 1)])
            output += int(np.prod([self.encoder.output_channels[-(s + 2)], *skip_sizes[-(s + 1)]], dtype=np.int64))
            if self.deep_supervision or (s == (len(self.stages) - 1)):
                output += int(np.prod([self.num_classes, *skip_sizes[-(s + 1)]], dtype=np.int64))
        return output


class PlainConvUNet(nn.Module):
    """kwargs as passed at get_network_from_plans.py:70-83."""

    def __init__(self, input_channels: int, n_stages: int, features_per_stage, kernel_sizes, strides,
                 n_conv_per_stage, num_classes: int, n_conv_per_stage_decoder, conv_bias: bool = True,
                 deep_supervision: bool = True):
        super().__init__()
        if isinstance(n_conv_per_stage, int):
            n_conv_per_stage = [n_conv_per_stage] * n_stages
        if isinstance(n_conv_per_stage_decoder, int):
            n_conv_per_stage_decoder = [n_conv_per_stage_decoder] * (n_stages - 1)
        self.encoder = PlainConvEncoder(input_channels, n_stages, features_per_stage, kernel_sizes, strides,
                                        n_conv_per_stage, conv_bias, return_skips=True)
        self.decoder = UNetDecoder(self.encoder, num_classes, n_conv_per_stage_decoder, deep_supervision)

    def forward(self, x):
        return self.decoder(self.encoder(x))

    def compute_conv_feature_map_size(self, input_size):
        return self.encoder.compute_conv_feature_map_size(input_size) + \
            self.decoder.compute_conv_feature_map_size(input_size)


def get_pool_and_conv_props(spacing, patch_size, min_feature_map_size, max_numpool):
    """Restatement of network_topology.py:30-105 (without the final pad_shape, which callers here never need
    because every configured patch is already divisible).  Returns (num_pool_per_axis, pool_op_kernel_sizes,
    conv_kernel_sizes)."""
    dim = len(spacing)
    current_spacing = deepcopy(list(spacing))
    current_size = deepcopy(list(patch_size))
    pool_op_kernel_sizes = [[1] * dim]
    conv_kernel_sizes = []
    num_pool_per_axis = [0] * dim
    kernel_size = [1] * dim
    while True:
        valid = [i for i in range(dim) if current_size[i] >= 2 * min_feature_map_size]
        if len(valid) < 1:
            break
        spacings_of_axes = [current_spacing[i] for i in valid]
        min_spacing_of_valid = min(spacings_of_axes)
        valid = [i for i in valid if current_spacing[i] / min_spacing_of_valid < 2]
        valid = [i for i in valid if num_pool_per_axis[i] < max_numpool]
        if len(valid) == 1:
            if current_size[valid[0]] >= 3 * min_feature_map_size:
                pass
            else:
                break
        if len(valid) < 1:
            break
        for d in range(dim):
            if kernel_size[d] == 3:
                continue
            # the reference indexes spacings_of_axes (length = #valid at that point) with the axis id
            # (network_topology.py:85); kept verbatim, guarded so anisotropic inputs fail the same way
            if spacings_of_axes[d] / min(current_spacing) < 2:
                kernel_size[d] = 3
        other_axes = [i for i in range(dim) if i not in valid]
        pool_kernel_sizes = [0] * dim
        for v in valid:
            pool_kernel_sizes[v] = 2
            num_pool_per_axis[v] += 1
            current_spacing[v] *= 2
            current_size[v] = np.ceil(current_size[v] / 2)
        for nv in other_axes:
            pool_kernel_sizes[nv] = 1
        pool_op_kernel_sizes.append(pool_kernel_sizes)
        conv_kernel_sizes.append(deepcopy(kernel_size))
    conv_kernel_sizes.append([3] * dim)
    return num_pool_per_axis, pool_op_kernel_sizes, conv_kernel_sizes


def topology_for_patch(patch_size: Sequence[int], spacing: Sequence[float] = (1.0, 1.0, 1.0),
                       base_features: int = 32, max_features: int = 320, min_edge: int = 4):
    """Layer shapes for a 3d_fullres config: topology rule network_topology.py:30-105, bottleneck edge 4
    (default_experiment_planner.py:61), features min(32*2^i, 320) (get_network_from_plans.py:73-74),
    2 convs per stage (default_experiment_planner.py:62-63)."""
    _, pool_ks, conv_ks = get_pool_and_conv_props(spacing, patch_size, min_edge, 999999)
    n_stages = len(pool_ks)
    return dict(n_stages=n_stages,
                features_per_stage=[min(base_features * 2 ** i, max_features) for i in range(n_stages)],
                kernel_sizes=[list(k) for k in conv_ks],
                strides=[list(int(j) for j in k) for k in pool_ks],
                n_conv_per_stage=[2] * n_stages,
                n_conv_per_stage_decoder=[2] * (n_stages - 1))


def build_plain_conv_unet(input_channels: int, num_classes: int, patch_size: Sequence[int],
                          deep_supervision: bool = True, seed: Union[int, None] = 0) -> PlainConvUNet:
    """get_network_from_plans.py:15-92 for the PlainConvUNet branch, with the topology derived from the patch."""
    topo = topology_for_patch(patch_size)
    if seed is not None:
        torch.manual_seed(seed)
    net = PlainConvUNet(input_channels=input_channels, num_classes=num_classes, conv_bias=True,
                        deep_supervision=deep_supervision, **topo)
    net.apply(InitWeights_He(1e-2))
    return net
