"""ORACLE (test infrastructure only; never imported by the product path): CPU restatement of the deep-supervision
target transform.

Follows ``DownsampleSegForDSTransform2.__call__`` (nnunetv2/training/data_augmentation/custom_transforms/
deep_supervision_donwsampling.py:27-55) line by line.  The resampling itself lives in a third-party dependency that is
NOT in the tree and not installable here: ``batchgenerators>=0.25`` (setup.py:20) ``augmentations.utils.
resize_segmentation(segmentation, new_shape, order)``, which for order 0 returns ``skimage.transform.resize(
segmentation.astype(float), new_shape, order=0, mode="edge", clip=True, anti_aliasing=False).astype(dtype)``;
scikit-image >= 0.19 implements that call as ``scipy.ndimage.zoom(image, new/old, order=0, mode='nearest',
grid_mode=True)``.  scipy IS installed, so ``resize_segmentation`` below calls exactly that, and
``nearest_index`` states the closed form the CUDA kernel uses (src = floor((o + 0.5) * I / O)); the CPU tests pin one
against the other.  PARITY UNPINNED by the reference's own tests (it has none for this transform)."""
import numpy as np
from scipy import ndimage


def nearest_index(out_size: int, in_size: int) -> np.ndarray:
    o = np.arange(out_size, dtype=np.int64)
    return np.minimum(((2 * o + 1) * in_size) // (2 * out_size), in_size - 1)


def resize_segmentation(segmentation: np.ndarray, new_shape, order: int = 0) -> np.ndarray:
    assert order == 0, 'only the nearest-neighbour branch is restated'
    assert len(segmentation.shape) == len(new_shape)
    zoom = [n / o for n, o in zip(new_shape, segmentation.shape)]
    out = ndimage.zoom(segmentation.astype(float), zoom, order=0, mode='nearest', grid_mode=True)
    assert out.shape == tuple(new_shape)
    return out.astype(segmentation.dtype)


def resize_segmentation_closed_form(segmentation: np.ndarray, new_shape) -> np.ndarray:
    idx = [nearest_index(n, o) for n, o in zip(new_shape, segmentation.shape)]
    return segmentation[np.ix_(*idx)]


class DownsampleSegForDSTransform2:
    """same constructor and call contract as the reference transform (:13-25, :27-55)."""

    def __init__(self, ds_scales, order: int = 0, input_key: str = 'seg', output_key: str = 'seg', axes=None):
        self.ds_scales, self.order, self.input_key, self.output_key, self.axes = ds_scales, order, input_key, output_key, axes

    def _scaled_shape(self, shape, axes, scale):
        # :46-49 -- float shape, multiplied per listed axis, np.round (half to even), int
        target = np.asarray(shape, dtype=float)
        target[list(axes)] *= np.asarray(scale, dtype=float)
        return tuple(int(v) for v in np.round(target))

    def __call__(self, **data_dict):
        seg = data_dict[self.input_key]
        axes = tuple(range(2, seg.ndim)) if self.axes is None else tuple(self.axes)            # :28-31
        pyramid = []
        for scale in self.ds_scales:                                                            # :34
            per_axis = list(scale) if isinstance(scale, (tuple, list)) else [scale] * len(axes)
            assert len(per_axis) == len(axes), 'one downsampling factor per axis'
            if all(f == 1 for f in per_axis):                                                   # :43-44: the input itself
                pyramid.append(seg)
                continue
            shape = self._scaled_shape(seg.shape, axes, per_axis)
            level = np.zeros(shape, dtype=seg.dtype)                                            # :50
            for b, c in np.ndindex(*seg.shape[:2]):                                             # :51-52: per (sample, channel)
                level[b, c] = resize_segmentation(seg[b, c], shape[2:], self.order)
            pyramid.append(level)
        data_dict[self.output_key] = pyramid
        return data_dict
