"""Sliding-window inference of the hot path's network.

Mirrors the pieces of the reference that run at prediction time:
  compute_gaussian, compute_steps_for_sliding_window        inference/sliding_window_prediction.py:10-58
  nnUNetPredictor._internal_get_sliding_window_slicers       inference/predict_from_raw_data.py:528-560
  nnUNetPredictor._internal_maybe_mirror_and_predict         inference/predict_from_raw_data.py:562-589
  nnUNetPredictor.predict_sliding_window_return_logits       inference/predict_from_raw_data.py:643-714
The network forward is the tcgen05 path of this package; the Gaussian-weighted accumulation of every tile into the
full-volume logits runs in libmvdseg (``mvd_sw_accumulate`` / ``mvd_sw_finalize``).  Accumulators are fp32 (the
reference keeps them in fp16) and every mirrored pass is added un-averaged with scale 1/2^n, so nothing between the
network's bf16 logits and the final division is rounded to 16 bits.  There is no CPU fallback.
"""
from typing import List, Optional, Sequence, Tuple, Union

import numpy as np
import torch
from scipy.ndimage import gaussian_filter

from . import ops
from ._lib import MvdError, lib


def compute_gaussian(tile_size: Union[Tuple[int, ...], List[int]], sigma_scale: float = 1. / 8,
                     value_scaling_factor: float = 1, dtype=torch.float32, device=torch.device('cuda', 0)) -> torch.Tensor:
    """sliding_window_prediction.py:10-30 (same scipy call; zeros replaced by the smallest non-zero weight)."""
    tmp = np.zeros(tile_size)
    center_coords = [i // 2 for i in tile_size]
    sigmas = [i * sigma_scale for i in tile_size]
    tmp[tuple(center_coords)] = 1
    g = gaussian_filter(tmp, sigmas, 0, mode='constant', cval=0)
    g = torch.from_numpy(g).type(dtype).to(device)
    g = g / torch.max(g) * value_scaling_factor
    g = g.type(dtype)
    g[g == 0] = torch.min(g[g != 0])
    return g


def compute_steps_for_sliding_window(image_size: Sequence[int], tile_size: Sequence[int], tile_step_size: float) \
        -> List[List[int]]:
    """sliding_window_prediction.py:33-58."""
    assert all(i >= j for i, j in zip(image_size, tile_size)), 'image size must be as large or larger than patch_size'
    assert 0 < tile_step_size <= 1, 'step_size must be larger than 0 and smaller or equal to 1'
    target = [i * tile_step_size for i in tile_size]
    num_steps = [int(np.ceil((i - k) / j)) + 1 for i, j, k in zip(image_size, target, tile_size)]
    steps = []
    for dim in range(len(tile_size)):
        max_step_value = image_size[dim] - tile_size[dim]
        actual = max_step_value / (num_steps[dim] - 1) if num_steps[dim] > 1 else 99999999999
        steps.append([int(np.round(actual * i)) for i in range(num_steps[dim])])
    return steps


class SlidingWindowPredictor:
    """the prediction-time subset of nnUNetPredictor (predict_from_raw_data.py) for 3-D configurations.

    ``network``: a PlainConvUNet of this package (deep supervision is switched off for the call, as
    nnUNetTrainer.set_deep_supervision_enabled(False) does before validation)."""

    def __init__(self, network, patch_size: Sequence[int], num_segmentation_heads: int, tile_step_size: float = 0.5,
                 use_gaussian: bool = True, use_mirroring: bool = True,
                 allowed_mirroring_axes: Optional[Sequence[int]] = (0, 1, 2), device=torch.device('cuda', 0)):
        self.network = network
        self.patch_size = tuple(int(i) for i in patch_size)
        assert len(self.patch_size) == 3, 'only 3-D configurations are built'
        self.num_segmentation_heads = int(num_segmentation_heads)
        self.tile_step_size = tile_step_size
        self.use_gaussian = use_gaussian
        self.use_mirroring = use_mirroring
        self.allowed_mirroring_axes = tuple(allowed_mirroring_axes) if allowed_mirroring_axes is not None else None
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise MvdError('SlidingWindowPredictor: libmvdseg has no CPU path')

    # predict_from_raw_data.py:528-560 (3-D branch)
    def _internal_get_sliding_window_slicers(self, image_size: Sequence[int]):
        steps = compute_steps_for_sliding_window(image_size, self.patch_size, self.tile_step_size)
        return [tuple([slice(None), *[slice(si, si + ti) for si, ti in zip((sx, sy, sz), self.patch_size)]])
                for sx in steps[0] for sy in steps[1] for sz in steps[2]]

    # predict_from_raw_data.py:562-589
    def _mirror_passes(self) -> List[Tuple[Tuple[int, ...], int]]:
        """(flip dims of the [1,c,x,y,z] input, flip_mask of mvd_sw_accumulate) for every test-time-augmentation pass:
        the identity first, then the reference's order (2,), (3,), (4,), (2,3), (2,4), (3,4), (2,3,4) restricted to the
        allowed axes."""
        passes = [((), 0)]
        mirror_axes = self.allowed_mirroring_axes if self.use_mirroring else None
        if mirror_axes is not None:
            assert max(mirror_axes) <= 2, 'mirror_axes does not match the dimension of the input!'
            for c in ((2,), (3,), (4,), (2, 3), (2, 4), (3, 4), (2, 3, 4)):
                if all((a - 2) in mirror_axes for a in c):
                    passes.append((c, sum(1 << (a - 2) for a in c)))
        return passes

    def _internal_maybe_mirror_and_predict(self, x: torch.Tensor) -> torch.Tensor:
        """mean over the mirrored passes as an fp32 [1,K,x,y,z] tensor (the reference's return value; the sliding
        window below does not go through it: it hands every pass to mvd_sw_accumulate un-averaged)."""
        passes = self._mirror_passes()
        prediction = None
        for dims, _ in passes:
            p = self.network(torch.flip(x, dims) if dims else x).float()
            p = torch.flip(p, dims) if dims else p
            prediction = p if prediction is None else prediction + p
        return prediction / len(passes)

    @torch.no_grad()
    def predict_sliding_window_return_logits(self, input_image: torch.Tensor) -> torch.Tensor:
        """input_image (c, x, y, z) -> logits (num_segmentation_heads, x, y, z), fp32 on the device."""
        assert isinstance(input_image, torch.Tensor) and input_image.dim() == 4, \
            'input_image must be a 4D torch.Tensor (c, x, y, z)'
        net = self.network
        was_training, ds = net.training, net.decoder.deep_supervision
        net.eval()
        net.decoder.deep_supervision = False
        try:
            data = input_image.to(self.device, torch.float32)
            # pad_nd_image(..., patch_size, 'constant', value 0): centred zero padding up to the patch size (:666-668)
            pads, revert = [], [slice(None)]
            for s, p in zip(data.shape[1:], self.patch_size):
                total = max(p - s, 0)
                lo = total // 2
                pads.append((lo, total - lo))
                revert.append(slice(lo, lo + s))
            if any(a or b for a, b in pads):
                data = torch.nn.functional.pad(data, [v for ab in reversed(pads) for v in ab])
            D, H, W = data.shape[1:]
            K = self.num_segmentation_heads
            slicers = self._internal_get_sliding_window_slicers((D, H, W))
            acc = torch.zeros((K, D, H, W), dtype=torch.float32, device=self.device)
            npred = torch.zeros((D, H, W), dtype=torch.float32, device=self.device)
            gaussian = compute_gaussian(self.patch_size, sigma_scale=1. / 8, value_scaling_factor=1000,
                                        dtype=torch.float32, device=self.device) if self.use_gaussian else None
            st = torch.cuda.current_stream().cuda_stream
            d, h, w = self.patch_size
            passes = self._mirror_passes()
            scale = 1.0 / len(passes)     # = 1 / 2^len(mirror_axes), predict_from_raw_data.py:588
            for sl in slicers:
                workon = data[sl][None]
                for pi, (dims, mask) in enumerate(passes):
                    # every pass's bf16 logits go straight into the fp32 accumulator (flipped back by the kernel's
                    # addressing): the averaged prediction is never rounded to 16 bits
                    pred_cl = ops.to_cl_view(self.network(torch.flip(workon, dims) if dims else workon))[0]
                    lib.sw_accumulate(pred_cl.data_ptr(), ops.cl_pitch(pred_cl[None]), ops._ptr(gaussian), scale,
                                      acc.data_ptr(), npred.data_ptr() if pi == 0 else None, K, d, h, w, D, H, W,
                                      sl[1].start, sl[2].start, sl[3].start, mask, st)
            lib.sw_finalize(acc.data_ptr(), npred.data_ptr(), K, D * H * W, st)
            return acc[tuple(revert)]
        finally:
            net.decoder.deep_supervision = ds
            net.train(was_training)

    def predict_segmentation(self, input_image: torch.Tensor) -> torch.Tensor:
        """argmax over the class axis of the sliding-window logits (export_prediction.py: label_manager argmax)."""
        return self.predict_sliding_window_return_logits(input_image).argmax(0)
