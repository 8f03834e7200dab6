"""Deep-supervision targets produced on the GPU.

Host-side mirror of ``DownsampleSegForDSTransform2`` (nnunetv2/training/data_augmentation/custom_transforms/
deep_supervision_donwsampling.py:8-55): same constructor, same ``__call__(**data_dict)`` contract, same output list
(``ds_scales`` order, an all-ones scale returns the input object itself, :43-44), but the segmentation is a CUDA tensor
and every scale comes out of ONE launch of ``mvd_downsample_seg_nearest``.  The reference resamples per (b, c) with
batchgenerators' ``resize_segmentation(seg, new_shape, order)`` (:52); only ``order = 0`` (what nnU-Net uses for
segmentations, MVDTrainer.py:757-760) is built -- anything else raises.  No CPU fallback."""
import ctypes
from typing import List, Tuple, Union

import numpy as np
import torch

from ._lib import lib

__all__ = ['DownsampleSegForDSTransform2', 'downsample_seg_for_ds']


def _new_shape(shape, axes, s) -> Tuple[int, ...]:
    # deep_supervision_donwsampling.py:46-49: float shape, scaled per axis, np.round (half to even), int
    new_shape = np.array(shape).astype(float)
    for i, a in enumerate(axes):
        new_shape[a] *= s[i]
    return tuple(int(v) for v in np.round(new_shape).astype(int))


def downsample_seg_for_ds(seg: torch.Tensor, ds_scales, axes=None) -> List[torch.Tensor]:
    if not (isinstance(seg, torch.Tensor) and seg.is_cuda):
        raise RuntimeError('downsample_seg_for_ds: the segmentation must be a CUDA tensor (no CPU path in this package)')
    if seg.dim() != 5:
        raise ValueError('downsample_seg_for_ds: expected a (B, C, D, H, W) segmentation')
    if axes is None:
        axes = list(range(2, seg.dim()))
    if any(a < 2 or a > 4 for a in axes):
        raise ValueError('downsample_seg_for_ds: axes must be spatial axes (2, 3, 4)')
    src = seg if (seg.dtype == torch.float32 and seg.is_contiguous()) else seg.float().contiguous()
    out: List[torch.Tensor] = []
    todo = []
    for s in ds_scales:
        if not isinstance(s, (tuple, list)):
            s = [s] * len(axes)
        else:
            assert len(s) == len(axes), 'one downsampling factor per axis'
        if all(i == 1 for i in s):
            out.append(seg)
            continue
        t = torch.empty(_new_shape(seg.shape, axes, s), dtype=torch.float32, device=seg.device)
        todo.append(t)
        out.append(t)
    B, C, D, H, W = src.shape
    for k in range(0, len(todo), 8):
        part = todo[k:k + 8]
        ptrs = (ctypes.c_void_p * len(part))(*[t.data_ptr() for t in part])
        dhw = (ctypes.c_int * (3 * len(part)))(*[d for t in part for d in t.shape[2:]])
        lib.downsample_seg_nearest(src.data_ptr(), B * C, D, H, W, len(part), ptrs, dhw,
                                   torch.cuda.current_stream(seg.device).cuda_stream)
    if seg.dtype != torch.float32:
        out = [o if o is seg else o.to(seg.dtype) for o in out]
    return out


class DownsampleSegForDSTransform2:
    """data_dict[output_key] becomes the list of segmentations scaled according to ds_scales."""

    def __init__(self, ds_scales: Union[List, Tuple], order: int = 0, input_key: str = 'seg', output_key: str = 'seg',
                 axes: Tuple[int] = None):
        if order != 0:
            raise NotImplementedError('only order = 0 (nearest neighbour, what nnU-Net uses for segmentations) is built')
        self.axes = axes
        self.output_key = output_key
        self.input_key = input_key
        self.order = order
        self.ds_scales = ds_scales

    def __call__(self, **data_dict):
        data_dict[self.output_key] = downsample_seg_for_ds(data_dict[self.input_key], self.ds_scales, self.axes)
        return data_dict
