"""oracle/synthetic.py -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Deterministic synthetic batches.
  * 'rand'       : the reference's own benchmark convention -- data = torch.rand, target = round(rand * max_label)
                   independently per deep-supervision scale
                   (variants/benchmarking/nnUNetTrainerBenchmark_5epochs_noDataLoading.py:16-22)
  * 'structured' : data ~ N(0,1) (z-scored MRI, preprocessing/normalization/default_normalization_schemes.py:27) and a
                   label volume with blobs for classes 1,3 and thin tubes for class 2 (the vessel class,
                   MVDTrainer.py:897,907), nearest-neighbour downsampled per scale like
                   training/data_augmentation/custom_transforms/deep_supervision_donwsampling.py:45-52
"""
from typing import List, Sequence

import numpy as np
import torch
import torch.nn.functional as F


def ds_shapes(patch: Sequence[int], strides: Sequence[Sequence[int]]) -> List[List[int]]:
    """Spatial shapes of the deep-supervision outputs, hi-res first: cumulative strides, all stages but the
    bottleneck (MVDTrainer._get_deep_supervision_scales, MVDTrainer.py:311-314)."""
    shapes, cur = [], list(patch)
    for s in strides[:-1]:
        cur = [c // int(k) for c, k in zip(cur, s)]
        shapes.append(list(cur))
    return shapes


def structured_labels(B: int, patch: Sequence[int], seed: int = 4321, n_tubes: int = 6) -> torch.Tensor:
    """(B,1,D,H,W) float32 holding {0,1,2,3}; class 2 = thin random-walk tubes, 1 and 3 = smooth blobs."""
    g = torch.Generator().manual_seed(seed)
    D, H, W = patch
    noise = torch.randn((B, 2, D, H, W), generator=g)
    k = max(3, (min(patch) // 8) | 1)
    smooth = F.avg_pool3d(F.avg_pool3d(noise, k, 1, k // 2), k, 1, k // 2)
    smooth = smooth / smooth.flatten(2).std(-1)[:, :, None, None, None]
    lab = torch.zeros((B, D, H, W), dtype=torch.float32)
    lab[smooth[:, 0] > 0.9] = 1
    lab[smooth[:, 1] > 1.1] = 3
    rng = np.random.RandomState(seed)
    for b in range(B):
        for _ in range(n_tubes):
            p = np.array([rng.randint(0, D), rng.randint(0, H), rng.randint(0, W)], dtype=np.float64)
            d = rng.randn(3)
            d /= np.linalg.norm(d) + 1e-9
            for _ in range(2 * max(patch)):
                d = d + 0.25 * rng.randn(3)
                d /= np.linalg.norm(d) + 1e-9
                p = p + d
                q = np.round(p).astype(int)
                if (q < 0).any() or q[0] >= D or q[1] >= H or q[2] >= W:
                    break
                lab[b, max(q[0] - 1, 0):q[0] + 1, max(q[1] - 1, 0):q[1] + 1, q[2]] = 2
                lab[b, q[0], q[1], q[2]] = 2
    return lab[:, None]


def make_batch(B: int, Cin: int, patch: Sequence[int], strides: Sequence[Sequence[int]], max_label: int = 3,
               seed: int = 1234, kind: str = 'rand') -> dict:
    """{'data': (B,Cin,*patch) f32, 'target': [ (B,1,*patch/scale_i) f32 ]} -- the dict train_step consumes
    (MVDTrainer.py:762-765)."""
    g = torch.Generator().manual_seed(seed)
    shapes = ds_shapes(patch, strides)
    if kind == 'rand':
        data = torch.rand((B, Cin, *patch), generator=g)
        target = [torch.round(torch.rand((B, 1, *s), generator=g) * max_label) for s in shapes]
    elif kind == 'structured':
        data = torch.randn((B, Cin, *patch), generator=g)
        full = structured_labels(B, patch, seed=seed + 3087)
        target = []
        for s in shapes:
            if list(s) == list(patch):
                target.append(full.clone())
            else:
                target.append(F.interpolate(full, size=tuple(s), mode='nearest'))
    else:
        raise ValueError(kind)
    return {'data': data, 'target': target}
