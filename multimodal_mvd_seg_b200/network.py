"""PlainConvUNet on libmvdseg: the nn.Module interface of the network nnunetv2 builds for 3d_fullres.

Drop-in for the object returned by ``get_network_from_plans`` (reference:
nnunetv2/utilities/get_network_from_plans.py:15-92; class home: dynamic_network_architectures, un-vendored):
  * same module tree and state_dict keys (encoder.stages.S.0.convs.I.{conv,norm,all_modules.N}.*,
    decoder.{encoder.*,stages,transpconvs,seg_layers}.*) so checkpoints and ``load_pretrained_weights``
    (run/load_pretrained_weights.py:6-64) interchange;
  * ``forward(x[B,Cin,D,H,W]) -> list of logits, hi-res first`` when ``decoder.deep_supervision`` else one tensor
    (training/my_network/UNetDecoder.py:104-121); logits carry no final nonlinearity;
  * ``.encoder``, ``.decoder``, ``.decoder.deep_supervision``, ``.decoder.seg_layers``,
    ``compute_conv_feature_map_size`` (experiment_planning/.../default_experiment_planner.py:112).
The nn.Conv3d / nn.InstanceNorm3d / nn.ConvTranspose3d children are parameter containers only: their ATen/cuDNN
forwards are never called.  Every device computation goes through the hand-written kernels (ops.py); activations
are bf16 NDHWC between layers, fp32 accumulation, and the decoder's torch.cat (UNetDecoder.py:107) is replaced by
producers writing straight into the two channel halves of one buffer.
"""
from typing import List, Optional, Sequence, Union

import numpy as np
import torch
from torch import nn

from . import ops
from .ops import ConvGeom, Slot


class InitWeights_He(object):
    """utilities/network_initialization.py:4-12."""

    def __init__(self, neg_slope: float = 1e-2):
        self.neg_slope = neg_slope

    def __call__(self, module):
        if isinstance(module, (nn.Conv3d, nn.Conv2d, nn.ConvTranspose2d, nn.ConvTranspose3d)):
            module.weight = nn.init.kaiming_normal_(module.weight, a=self.neg_slope)
            if module.bias is not None:
                module.bias = nn.init.constant_(module.bias, 0)


def _t3(v):
    if isinstance(v, (int, np.integer)):
        return (int(v),) * 3
    v = tuple(int(i) for i in v)
    assert len(v) == 3, 'this build covers 3-D configurations (3d_fullres) only'
    return v


class ConvDropoutNormReLU(nn.Module):
    """Conv3d(k, stride, pad=(k-1)//2, bias) -> InstanceNorm3d(eps=1e-5, affine) -> LeakyReLU(0.01)
    (kwargs of get_network_from_plans.py:39-45; dropout_op is None there)."""

    def __init__(self, input_channels, output_channels, kernel_size, stride, conv_bias=True):
        super().__init__()
        k, s = _t3(kernel_size), _t3(stride)
        self.input_channels, self.output_channels, self.stride = input_channels, output_channels, s
        self.conv = nn.Conv3d(input_channels, output_channels, k, s, padding=[(i - 1) // 2 for i in k], bias=conv_bias)
        self.norm = nn.InstanceNorm3d(output_channels, eps=1e-5, affine=True)
        self.nonlin = nn.LeakyReLU(inplace=True)
        self.all_modules = nn.Sequential(self.conv, self.norm, self.nonlin)
        self._geom = ConvGeom(k, s, [(i - 1) // 2 for i in k])

    def forward_cl(self, x_cl: torch.Tensor, out: Optional[torch.Tensor] = None,
                   head: Optional[nn.Conv3d] = None, private_input: bool = False) -> torch.Tensor:
        """``head``: the 1x1x1 segmentation layer that is the only consumer of this block's output -- the call then
        returns the head's logits (InstanceNorm + LeakyReLU folded into the head, ops.ConvNormActFn).
        ``private_input``: x_cl is another block's output that nothing else reads (ops.ConvNormActFn)."""
        params = [p for p in (self.conv.weight, self.conv.bias, self.norm.weight, self.norm.bias) if p is not None]
        if head is not None:
            params = params + [p for p in (head.weight, head.bias) if p is not None]
            return ops.ConvNormActFn.apply(x_cl, self.conv.weight, self.conv.bias, self.norm.weight, self.norm.bias,
                                           self._geom, float(self.norm.eps), float(self.nonlin.negative_slope), None,
                                           params, head.weight, head.bias, private_input)
        return ops.ConvNormActFn.apply(x_cl, self.conv.weight, self.conv.bias, self.norm.weight, self.norm.bias,
                                       self._geom, float(self.norm.eps), float(self.nonlin.negative_slope),
                                       Slot(out) if out is not None else None, params, None, None, private_input)

    def forward(self, x):
        return ops.ncdhw_view(self.forward_cl(_to_cl(x)))

    def compute_conv_feature_map_size(self, input_size):
        output_size = [i // j for i, j in zip(input_size, self.stride)]
        return int(np.prod([self.output_channels, *output_size], dtype=np.int64))


def _to_cl(x: torch.Tensor) -> torch.Tensor:
    """public-API edge: logical [B,C,D,H,W] in any dtype/layout -> bf16 NDHWC."""
    if x.dtype == torch.bfloat16:
        return ops.to_cl_view(x)
    return ops.input_to_cl(x)


class StackedConvBlocks(nn.Module):
    def __init__(self, num_convs, input_channels, output_channels, kernel_size, initial_stride, conv_bias=True):
        super().__init__()
        if not isinstance(output_channels, (tuple, list)):
            output_channels = [output_channels] * num_convs
        self.convs = nn.Sequential(
            ConvDropoutNormReLU(input_channels, output_channels[0], kernel_size, initial_stride, conv_bias),
            *[ConvDropoutNormReLU(output_channels[i - 1], output_channels[i], kernel_size, 1, conv_bias)
              for i in range(1, num_convs)])
        self.output_channels = output_channels[-1]
        self.initial_stride = _t3(initial_stride)

    def forward_cl(self, x_cl, out: Optional[torch.Tensor] = None, head: Optional[nn.Conv3d] = None):
        n = len(self.convs)
        for i, blk in enumerate(self.convs):
            x_cl = blk.forward_cl(x_cl, out if i == n - 1 else None, head if i == n - 1 else None, private_input=i > 0)
        return x_cl

    def forward(self, x):
        return ops.ncdhw_view(self.forward_cl(_to_cl(x)))

    def compute_conv_feature_map_size(self, input_size):
        output = self.convs[0].compute_conv_feature_map_size(input_size)
        size_after_stride = [i // j for i, j in zip(input_size, self.initial_stride)]
        for b in self.convs[1:]:
            output += b.compute_conv_feature_map_size(size_after_stride)
        return output


class PlainConvEncoder(nn.Module):
    def __init__(self, input_channels, n_stages, features_per_stage, kernel_sizes, strides, n_conv_per_stage,
                 conv_bias=True, return_skips=True):
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
        stages, cin = [], input_channels
        for s in range(n_stages):
            stages.append(nn.Sequential(StackedConvBlocks(n_conv_per_stage[s], cin, features_per_stage[s],
                                                          kernel_sizes[s], strides[s], conv_bias)))
            cin = features_per_stage[s]
        self.stages = nn.Sequential(*stages)
        self.output_channels = list(features_per_stage)
        self.strides = [_t3(i) for i in strides]
        self.return_skips = return_skips
        self.conv_op = nn.Conv3d
        self.norm_op = nn.InstanceNorm3d
        self.norm_op_kwargs = {'eps': 1e-5, 'affine': True}
        self.nonlin = nn.LeakyReLU
        self.nonlin_kwargs = {'inplace': True}
        self.dropout_op = None
        self.dropout_op_kwargs = None
        self.conv_bias = conv_bias
        self.kernel_sizes = [_t3(k) for k in kernel_sizes]

    def forward_cl(self, x_cl, skip_outs: Optional[List[Optional[torch.Tensor]]] = None):
        ret = []
        for s, stage in enumerate(self.stages):
            out = skip_outs[s] if skip_outs is not None else None
            x_cl = stage[0].forward_cl(x_cl, out)
            ret.append(x_cl)
        return ret

    def forward(self, x):
        skips = [ops.ncdhw_view(t) for t in self.forward_cl(_to_cl(x))]
        return skips if self.return_skips else skips[-1]

    def compute_conv_feature_map_size(self, input_size):
        output = 0
        for s in range(len(self.stages)):
            output += self.stages[s][-1].compute_conv_feature_map_size(input_size)
            input_size = [i // j for i, j in zip(input_size, self.strides[s])]
        return output


class UNetDecoder(nn.Module):
    """ctor UNetDecoder.py:35-74; forward :104-121 (the fork's attention lines :75-102 are not part of
    PlainConvUNet)."""

    def __init__(self, encoder: PlainConvEncoder, num_classes: int, n_conv_per_stage, deep_supervision: bool):
        super().__init__()
        self.deep_supervision = deep_supervision
        self.encoder = encoder
        self.num_classes = num_classes
        n_enc = len(encoder.output_channels)
        if isinstance(n_conv_per_stage, int):
            n_conv_per_stage = [n_conv_per_stage] * (n_enc - 1)
        assert len(n_conv_per_stage) == n_enc - 1
        stages, transpconvs, seg_layers = [], [], []
        for s in range(1, n_enc):
            below, skip, st = encoder.output_channels[-s], encoder.output_channels[-(s + 1)], encoder.strides[-s]
            transpconvs.append(nn.ConvTranspose3d(below, skip, st, st, bias=encoder.conv_bias))
            stages.append(StackedConvBlocks(n_conv_per_stage[s - 1], 2 * skip, skip, encoder.kernel_sizes[-(s + 1)], 1,
                                            encoder.conv_bias))
            seg_layers.append(nn.Conv3d(skip, num_classes, 1, 1, 0, bias=True))
        self.stages = nn.ModuleList(stages)
        self.transpconvs = nn.ModuleList(transpconvs)
        self.seg_layers = nn.ModuleList(seg_layers)

    def concat_buffers(self, skip_shapes, device):
        """one [B,D,H,W,2C] buffer per decoder stage; returns (buffers, encoder skip destinations)."""
        n = len(self.stages)
        bufs, skip_outs = [], [None] * (n + 1)
        for s in range(n):
            B, D, H, W, C = skip_shapes[-(s + 2)]
            buf = torch.empty((B, D, H, W, 2 * C), dtype=torch.bfloat16, device=device)
            bufs.append(buf)
            skip_outs[n - 1 - s] = buf[..., C:]
        return bufs, skip_outs

    def forward_cl(self, skips_cl, bufs=None):
        lres = skips_cl[-1]
        seg = []
        for s in range(len(self.stages)):
            tc = self.transpconvs[s]
            skip = skips_cl[-(s + 2)]
            C = skip.shape[-1]
            tparams = [p for p in (tc.weight, tc.bias) if p is not None]
            if bufs is not None:
                buf = bufs[s]
                up = ops.ConvTransposeFn.apply(lres, tc.weight, tc.bias, tuple(tc.stride), Slot(buf[..., :C]), tparams)
                x = ops.ConcatViewFn.apply(up, skip, Slot(buf))
            else:
                up = ops.ConvTransposeFn.apply(lres, tc.weight, tc.bias, tuple(tc.stride), None, tparams)
                x = torch.cat((up, skip), -1)
            last = s == len(self.stages) - 1
            head = (self.seg_layers[s] if self.deep_supervision else self.seg_layers[-1]) \
                if (self.deep_supervision or last) else None
            if last and ops.head_fusion_ok(self.stages[s].output_channels, head):
                # nothing but the head reads the full-resolution stage output: InstanceNorm + LeakyReLU of its last
                # block run inside the head kernels, the activation itself is never written
                seg.append(self.stages[s].forward_cl(x, head=head))
                break
            x = self.stages[s].forward_cl(x)
            if head is not None:
                seg.append(ops.HeadFn.apply(x, head.weight, head.bias, [head.weight, head.bias]))
            lres = x
        seg = seg[::-1]
        return seg if self.deep_supervision else seg[0]

    def forward(self, skips):
        r = self.forward_cl([_to_cl(s) for s in skips])
        return [ops.ncdhw_view(t) for t in r] if isinstance(r, list) else ops.ncdhw_view(r)

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
    """kwargs as passed at get_network_from_plans.py:70-83 (conv_op / norm_op / nonlin arguments are accepted and
    checked: only the 3d_fullres instantiation Conv3d + InstanceNorm3d(affine) + LeakyReLU is built)."""

    def __init__(self, input_channels, n_stages, features_per_stage, conv_op=nn.Conv3d, kernel_sizes=3, strides=1,
                 n_conv_per_stage=2, num_classes=2, n_conv_per_stage_decoder=2, conv_bias=True,
                 norm_op=nn.InstanceNorm3d, norm_op_kwargs=None, dropout_op=None, dropout_op_kwargs=None,
                 nonlin=nn.LeakyReLU, nonlin_kwargs=None, deep_supervision=True, nonlin_first=False):
        super().__init__()
        assert conv_op is nn.Conv3d, 'only the 3-D (3d_fullres) path is built'
        assert norm_op is nn.InstanceNorm3d and nonlin is nn.LeakyReLU and dropout_op is None and not nonlin_first
        if norm_op_kwargs is not None:
            assert norm_op_kwargs.get('affine', True) and abs(norm_op_kwargs.get('eps', 1e-5) - 1e-5) < 1e-12
        if isinstance(n_conv_per_stage, int):
            n_conv_per_stage = [n_conv_per_stage] * n_stages
        if isinstance(n_conv_per_stage_decoder, int):
            n_conv_per_stage_decoder = [n_conv_per_stage_decoder] * (n_stages - 1)
        self.encoder = PlainConvEncoder(input_channels, n_stages, features_per_stage, kernel_sizes, strides,
                                        n_conv_per_stage, conv_bias, return_skips=True)
        self.decoder = UNetDecoder(self.encoder, num_classes, n_conv_per_stage_decoder, deep_supervision)

    def _weight_packer(self):
        """one launch refreshes the bf16 GEMM layouts of every conv / transposed-conv weight (ops.WeightPacker);
        rebuilt when a weight's storage moved (``.to(device)``)."""
        ws = []
        for m in self.modules():
            if isinstance(m, ConvDropoutNormReLU):
                kpad = ops.stem_kpad_for(m.conv.weight, m.stride)
                if kpad or m.conv.weight.shape[1] > 4:
                    ws.append((m.conv.weight, kpad))
        ws += [(t.weight, 0) for t in self.decoder.transpconvs]
        seen, uniq = set(), []
        for w, kpad in ws:   # decoder.encoder aliases the encoder modules
            if id(w) not in seen:
                seen.add(id(w))
                uniq.append((w, kpad))
        sig = tuple(w.data_ptr() for w, _ in uniq)
        if getattr(self, '_packer_sig', None) != sig:
            self._packer = ops.WeightPacker([(w, True, not kpad, kpad) for w, kpad in uniq])
            self._packer_sig = sig
        return self._packer

    def forward(self, x: torch.Tensor):
        x_cl = _to_cl(x)
        ops.reset_skip_registry()
        if x_cl.is_cuda and ops._default_algo != 1:
            pk = self._weight_packer()
            if not pk.is_fresh():      # the fused optimiser kernel keeps the layouts current between steps
                pk.run()
            ops.set_active_packer(pk)
        try:
            return self._forward_impl(x_cl)
        finally:
            ops.set_active_packer(None)

    def _forward_impl(self, x_cl: torch.Tensor):
        B, D, H, W, _ = x_cl.shape
        shapes, cur = [], [D, H, W]
        for s, st in enumerate(self.encoder.strides):
            k = self.encoder.kernel_sizes[s]
            cur = [(c + 2 * ((kk - 1) // 2) - kk) // ss + 1 for c, kk, ss in zip(cur, k, st)]
            shapes.append((B, *cur, self.encoder.output_channels[s]))
        bufs, skip_outs = self.decoder.concat_buffers(shapes, x_cl.device)
        skips = self.encoder.forward_cl(x_cl, skip_outs)
        r = self.decoder.forward_cl(skips, bufs)
        return [ops.ncdhw_view(t) for t in r] if isinstance(r, list) else ops.ncdhw_view(r)

    def compute_conv_feature_map_size(self, input_size):
        return self.encoder.compute_conv_feature_map_size(input_size) + \
            self.decoder.compute_conv_feature_map_size(input_size)


def get_network_from_plans(plans_manager, dataset_json, configuration_manager, num_input_channels: int,
                           deep_supervision: bool = True):
    """utilities/get_network_from_plans.py:15-92 for the PlainConvUNet class name, duck-typed on the reference's
    PlansManager / ConfigurationManager (plans_handling/plans_handler.py:55-122)."""
    cm = configuration_manager
    num_stages = len(cm.conv_kernel_sizes)
    assert len(cm.conv_kernel_sizes[0]) == 3, 'only 3-D configurations are built'
    name = getattr(cm, 'UNet_class_name', 'PlainConvUNet')
    if name != 'PlainConvUNet':
        raise NotImplementedError(f'{name}: only PlainConvUNet is on the built hot path')
    label_manager = plans_manager.get_label_manager(dataset_json)
    model = PlainConvUNet(
        input_channels=num_input_channels, n_stages=num_stages,
        features_per_stage=[min(cm.UNet_base_num_features * 2 ** i, cm.unet_max_num_features)
                            for i in range(num_stages)],
        conv_op=nn.Conv3d, kernel_sizes=cm.conv_kernel_sizes, strides=cm.pool_op_kernel_sizes,
        num_classes=label_manager.num_segmentation_heads, deep_supervision=deep_supervision,
        n_conv_per_stage=cm.n_conv_per_stage_encoder, n_conv_per_stage_decoder=cm.n_conv_per_stage_decoder,
        conv_bias=True, norm_op=nn.InstanceNorm3d, norm_op_kwargs={'eps': 1e-5, 'affine': True},
        dropout_op=None, dropout_op_kwargs=None, nonlin=nn.LeakyReLU, nonlin_kwargs={'inplace': True})
    model.apply(InitWeights_He(1e-2))
    return model


def load_pretrained_weights(network: nn.Module, fname, verbose: bool = False) -> None:
    """contract of run/load_pretrained_weights.py:6-64: everything except the segmentation heads (keys containing
    '.seg_layers.') is taken from the checkpoint's 'network_weights'; each of those keys must exist there with the same
    shape, otherwise the checkpoint does not belong to this architecture (AssertionError).  ``fname``: a checkpoint
    path or an already loaded checkpoint dict."""
    checkpoint = fname if isinstance(fname, dict) else torch.load(fname, map_location='cpu', weights_only=False)
    source = checkpoint['network_weights']
    target = network.module if isinstance(network, nn.parallel.DistributedDataParallel) else network
    state = target.state_dict()
    wanted = [k for k in state if '.seg_layers.' not in k]
    absent = [k for k in wanted if k not in source]
    assert not absent, (f'Key {absent[0]} is missing in the pretrained model weights. The pretrained weights do not '
                        f'seem to be compatible with your network.')
    for k in wanted:
        assert state[k].shape == source[k].shape, \
            (f'The shape of the parameters of key {k} is not the same. Pretrained model: {source[k].shape}; your '
             f'network: {state[k].shape}.')
        if verbose:
            print(k, 'shape', tuple(source[k].shape))
        state[k] = source[k]
    target.load_state_dict(state)
