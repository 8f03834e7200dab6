"""Training-batch augmentation on the GPU.

Host-side mirror of the transform chain ``nnUNetTrainer.get_training_transforms`` builds (MVDTrainer.py:700-765) and the
reference runs in 12+ CPU worker processes per GPU (batchgenerators ``MultiThreadedAugmenter``): at ~200 patches/s per B200
those workers, not the training step, would set the pace.  Here the loader hands over the raw (oversized) patch and the
chain runs as a handful of HBM-bound kernels (csrc/augment.cu) on the batch that is already on the device:

  SpatialTransform(rotation, isotropic scale; order 3 images / order 1 per-label segmentation; constant border)  :700-711
  GaussianNoiseTransform(p 0.1)   GaussianBlurTransform((0.5, 1), per channel, p 0.2 / 0.5)                       :716-718
  BrightnessMultiplicativeTransform((0.75, 1.25), p 0.15)   ContrastAugmentationTransform(p 0.15)                  :719-720
  SimulateLowResolutionTransform(zoom (0.5, 1), per channel p 0.5, nearest down / cubic up, p 0.25)               :721-725
  GammaTransform((0.7, 1.5), invert, retain_stats, p 0.1)   GammaTransform((0.7, 1.5), retain_stats, p 0.3)        :726-727
  MirrorTransform(mirror_axes)   RemoveLabelTransform(-1, 0)   DownsampleSegForDSTransform2 (ds_targets.py)        :729-760

The random draws are made on the host with the reference's distributions (``sample_parameters``) and handed to the kernels
as small parameter arrays, so every transform is a deterministic function of (batch, parameters) -- which is what the
tests compare with the numpy / scipy restatement in ``oracle/augment.py``.
NOT built: elastic deformation (disabled in the reference call, p_el_per_sample = 0) and the cascade / region / mask
transforms (:732-752, not used by 3d_fullres without cascade).
No CPU path: the inputs must be CUDA tensors."""
import ctypes
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from ._lib import lib
from .ds_targets import downsample_seg_for_ds

__all__ = ['sample_parameters', 'rotation_scale_matrix', 'GpuAugmenter', 'augmented_batches']

_DEG30 = 30.0 / 360.0 * 2.0 * np.pi


def rotation_scale_matrix(ax: float, ay: float, az: float, scale: float) -> np.ndarray:
    """source offset = M @ centred output index, for batchgenerators' rotate_coords_3d (coordinate rows times
    Rx @ Ry @ Rz) followed by scale_coords: M = scale * (Rx Ry Rz)^T."""
    cx, sx, cy, sy, cz, sz = np.cos(ax), np.sin(ax), np.cos(ay), np.sin(ay), np.cos(az), np.sin(az)
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    return float(scale) * (rx @ ry @ rz).T


def _range_draw(rng, lo, hi):
    """the 'factor' draw of augment_contrast / augment_gamma / augment_spatial's scale: below 1 with probability 1/2"""
    if rng.random() < 0.5 and lo < 1:
        return rng.uniform(lo, 1.0)
    return rng.uniform(max(lo, 1.0), hi)


def sample_parameters(rng: np.random.Generator, B: int, C: int, rotation_for_DA: Optional[dict] = None,
                      mirror_axes: Sequence[int] = (0, 1, 2)) -> Dict[str, np.ndarray]:
    """One batch worth of random draws with the probabilities / ranges of MVDTrainer.py:700-730."""
    rot = rotation_for_DA or {'x': (-_DEG30, _DEG30), 'y': (-_DEG30, _DEG30), 'z': (-_DEG30, _DEG30)}
    P = dict(mat=np.zeros((B, 3, 3), np.float32), mode=np.zeros((B,), np.int32),
             noise_sigma=np.zeros((B, C), np.float32), blur_sigma=np.zeros((B, C), np.float32),
             brightness=np.ones((B, C), np.float32), contrast=np.zeros((B, C), np.float32),
             lowres_zoom=np.zeros((B, C), np.float32),
             gamma_inv=np.zeros((B, C), np.float32), gamma=np.zeros((B, C), np.float32),
             flips=np.zeros((B, 3), np.uint8))
    for b in range(B):
        ax = ay = az = 0.0
        sc = 1.0
        if rng.random() < 0.2:                                   # p_rot_per_sample, p_rot_per_axis = 1
            ax, ay, az = rng.uniform(*rot['x']), rng.uniform(*rot['y']), rng.uniform(*rot['z'])
            P['mode'][b] = 1
        if rng.random() < 0.2:                                   # p_scale_per_sample, scale = (0.7, 1.4), isotropic
            sc = _range_draw(rng, 0.7, 1.4)
            P['mode'][b] = 1
        P['mat'][b] = rotation_scale_matrix(ax, ay, az, sc)
        if rng.random() < 0.1:                                   # GaussianNoise: one draw from (0, 0.1) for the sample
            P['noise_sigma'][b, :] = rng.uniform(0.0, 0.1)
        if rng.random() < 0.2:                                   # GaussianBlur: per channel with p 0.5
            for c in range(C):
                if rng.random() <= 0.5:
                    P['blur_sigma'][b, c] = rng.uniform(0.5, 1.0)
        if rng.random() < 0.15:                                  # BrightnessMultiplicative, per channel
            P['brightness'][b, :] = rng.uniform(0.75, 1.25, size=C)
        if rng.random() < 0.15:                                  # Contrast, per channel
            for c in range(C):
                P['contrast'][b, c] = _range_draw(rng, 0.75, 1.25)
        if rng.random() < 0.25:                                  # SimulateLowResolution: per channel with p 0.5
            for c in range(C):
                if rng.random() < 0.5:
                    P['lowres_zoom'][b, c] = rng.uniform(0.5, 1.0)
        if rng.random() < 0.1:                                   # Gamma on the inverted image
            for c in range(C):
                P['gamma_inv'][b, c] = _range_draw(rng, 0.7, 1.5)
        if rng.random() < 0.3:                                   # Gamma
            for c in range(C):
                P['gamma'][b, c] = _range_draw(rng, 0.7, 1.5)
        for a in mirror_axes:
            if rng.random() < 0.5:
                P['flips'][b, a] = 1
    return P


def _stream(dev):
    return torch.cuda.current_stream(dev).cuda_stream


class GpuAugmenter:
    """``aug(data, seg, params=None)`` -> ``{'data': [B,C,*patch] fp32, 'target': [list of DS targets] | seg}``.

    data: CUDA fp32 [B, C, Di, Hi, Wi] (the loader's oversized patch, ``initial_patch_size`` of
    configure_rotation_dummyDA_mirroring_and_inital_patch_size, MVDTrainer.py:434-436; at least ``patch_size``);
    seg: CUDA [B, 1, Di, Hi, Wi] holding class ids (-1 allowed: removed as RemoveLabelTransform(-1, 0) does)."""

    def __init__(self, patch_size: Sequence[int], n_seg_labels: int, deep_supervision_scales=None,
                 rotation_for_DA: Optional[dict] = None, mirror_axes: Sequence[int] = (0, 1, 2), order_data: int = 3,
                 seed: int = 0):
        if len(patch_size) != 3:
            raise NotImplementedError('GPU augmentation is built for 3-D patches (3d_fullres)')
        if order_data not in (1, 3):
            raise NotImplementedError('order_resampling_data must be 1 or 3')
        self.patch_size = tuple(int(p) for p in patch_size)
        self.n_seg_labels = int(n_seg_labels)
        self.ds_scales = deep_supervision_scales
        self.rotation_for_DA, self.mirror_axes, self.order_data = rotation_for_DA, tuple(mirror_axes), order_data
        self.rng = np.random.default_rng(seed)
        self._calls = 0

    def sample(self, B: int, C: int) -> Dict[str, np.ndarray]:
        return sample_parameters(self.rng, B, C, self.rotation_for_DA, self.mirror_axes)

    def _upload(self, pieces: Dict[str, np.ndarray], dev) -> Dict[str, torch.Tensor]:
        """pack the arrays (16-byte aligned) into a pinned staging buffer, copy once, return typed device views"""
        offs, total = {}, 0
        for k, a in pieces.items():
            offs[k] = total
            total += (a.nbytes + 15) // 16 * 16
        # two staging buffers in turn: the asynchronous copy of the previous call may still be reading the other one
        ring = self.__dict__.setdefault('_staging', [None, None])
        turn = self.__dict__.get('_staging_turn', 0)
        self._staging_turn = turn ^ 1
        if ring[turn] is None or ring[turn][0].numel() < total:
            ring[turn] = (torch.empty((max(total, 4096),), dtype=torch.uint8).pin_memory(), torch.cuda.Event())
        else:
            ring[turn][1].synchronize()
        host, ev = ring[turn]
        hv = host.numpy()
        for k, a in pieces.items():
            hv[offs[k]:offs[k] + a.nbytes] = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
        d = host[:total].to(dev, non_blocking=True)
        ev.record(torch.cuda.current_stream(dev))
        tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.int32): torch.int32, np.dtype(np.uint8): torch.uint8,
               np.dtype(np.float64): torch.float64}
        return {k: d[offs[k]:offs[k] + a.nbytes].view(tdt[a.dtype]) for k, a in pieces.items()}

    def __call__(self, data: torch.Tensor, seg: torch.Tensor, params: Optional[Dict[str, np.ndarray]] = None,
                 noise_seed: Optional[int] = None):
        if not (data.is_cuda and seg.is_cuda):
            raise RuntimeError('GpuAugmenter: data and seg must be CUDA tensors (no CPU path in this package)')
        if data.dim() != 5 or seg.dim() != 5:
            raise ValueError('GpuAugmenter: expected (B, C, D, H, W) tensors')
        dev = data.device
        data = data.float().contiguous()
        segf = seg.float().contiguous()
        B, C, Di, Hi, Wi = data.shape
        D, H, W = self.patch_size
        if params is None:
            params = self.sample(B, C)
        st = _stream(dev)
        N, V = B * C, D * H * W
        # every parameter array of the call travels in ONE pinned buffer / one asynchronous copy (the dozen tiny
        # synchronous uploads this replaces cost more than the kernels of an average batch)
        tshape = np.zeros((N, 3), np.int32)
        if 'lowres_zoom' in params and params['lowres_zoom'].any():
            zf = params['lowres_zoom'].reshape(-1).astype(np.float64)
            sel = zf != 0
            if (zf[sel] > 1).any() or (zf[sel] <= 0).any():
                raise ValueError('lowres_zoom must lie in (0, 1] (SimulateLowResolutionTransform zoom_range, MVDTrainer.py:721)')
            tshape[sel] = np.maximum(np.round(np.array([D, H, W])[None, :] * zf[sel, None]), 1).astype(np.int32)   # np.round: half to even
        pieces = {'mat': params['mat'].astype(np.float32).reshape(-1), 'mode': params['mode'].astype(np.int32),
                  'tshape': tshape.reshape(-1), 'apply': np.repeat(params['mode'].astype(np.uint8), C),
                  'flips': params['flips'].astype(np.uint8).reshape(-1),
                  'stats_init': np.tile(np.array([0.0, 0.0, np.inf, -np.inf]), N),
                  'mm_init': np.tile(np.array([np.inf, -np.inf]), N)}
        for k in ('noise_sigma', 'blur_sigma', 'brightness', 'contrast', 'gamma_inv', 'gamma'):
            pieces[k] = params[k].astype(np.float32).reshape(-1)
        pd = self._upload(pieces, dev)
        mode, mat = pd['mode'], pd['mat']
        # ---- SpatialTransform ---------------------------------------------------------------------------------------
        src = data
        if self.order_data == 3 and int(params['mode'].max(initial=0)) > 0:
            src = data.clone()        # spline coefficients of the samples that are interpolated (the input stays intact)
            lib.aug_spline_prefilter(src.data_ptr(), B * C, Di, Hi, Wi, pd['apply'].data_ptr(), st)
        x = torch.empty((B, C, D, H, W), dtype=torch.float32, device=dev)
        lib.aug_spatial(src.data_ptr(), B, C, Di, Hi, Wi, x.data_ptr(), D, H, W, mat.data_ptr(), mode.data_ptr(),
                        self.order_data, 0.0, 0, st)
        s = torch.empty((B, segf.shape[1], D, H, W), dtype=torch.float32, device=dev)
        lib.aug_spatial(segf.data_ptr(), B, segf.shape[1], Di, Hi, Wi, s.data_ptr(), D, H, W, mat.data_ptr(),
                        mode.data_ptr(), 1, -1.0, self.n_seg_labels, st)
        # ---- intensity chain (in place, per-plane parameters) ------------------------------------------------------------
        if params['noise_sigma'].any():
            seed = noise_seed if noise_seed is not None else int(self.rng.integers(0, 2 ** 63 - 1))
            lib.aug_gaussian_noise(x.data_ptr(), V, N, pd['noise_sigma'].data_ptr(),
                                   ctypes.c_ulonglong(seed), st)
        tmp = None
        if params['blur_sigma'].any():
            tmp = torch.empty_like(x)
            lib.aug_gaussian_blur(x.data_ptr(), tmp.data_ptr(), N, D, H, W,
                                  pd['blur_sigma'].data_ptr(), st)
        if (params['brightness'] != 1).any():
            lib.aug_intensity(x.data_ptr(), V, N, 0, pd['brightness'].data_ptr(), None, None,
                              0, st)
        def stats():
            out = pd['stats_init'].clone()
            lib.aug_plane_stats(x.data_ptr(), V, N, out.data_ptr(), st)
            return out

        if params['contrast'].any():
            s0 = stats()
            lib.aug_intensity(x.data_ptr(), V, N, 1, pd['contrast'].data_ptr(),
                              s0.data_ptr(), None, 0, st)
        if tshape.any():
            stride = (D + 24) * (H + 24) * (W + 24)
            scratch = torch.empty((N, stride), dtype=torch.float32, device=dev)
            mm = pd['mm_init'].clone()
            lib.aug_simulate_lowres(x.data_ptr(), N, D, H, W, pd['tshape'].data_ptr(), scratch.data_ptr(), stride,
                                    mm.data_ptr(), st)
        for key, invert in (('gamma_inv', 1), ('gamma', 0)):
            if params[key].any():
                g = pd[key]
                s0 = stats()
                lib.aug_intensity(x.data_ptr(), V, N, 2, g.data_ptr(), s0.data_ptr(), None, invert, st)
                s1 = stats()
                lib.aug_intensity(x.data_ptr(), V, N, 3, g.data_ptr(), s0.data_ptr(), s1.data_ptr(), invert, st)
        # ---- MirrorTransform ---------------------------------------------------------------------------------------
        if params['flips'].any():
            flips = pd['flips']
            x2 = tmp if tmp is not None else torch.empty_like(x)
            lib.aug_mirror(x.data_ptr(), x2.data_ptr(), B, C, D, H, W, flips.data_ptr(), st)
            s2 = torch.empty_like(s)
            lib.aug_mirror(s.data_ptr(), s2.data_ptr(), B, s.shape[1], D, H, W, flips.data_ptr(), st)
            x, s = x2, s2
        self._calls += 1
        target = downsample_seg_for_ds(s, self.ds_scales) if self.ds_scales is not None else s
        return {'data': x, 'target': target}


def augmented_batches(raw_batches, augmenter: GpuAugmenter, device, data_key: str = 'data', seg_key: str = 'seg'):
    """Iterate the loader's RAW batches (``{'data': [B,C,Di,Hi,Wi], 'seg': [B,1,Di,Hi,Wi]}`` host tensors / arrays, what
    nnUNetDataLoader3D.generate_train_batch returns before any transform, data_loader_3d.py:7-48) and yield augmented
    device batches ``{'data', 'target'}`` ready for ``trainer.train_step`` -- the role of the reference's
    LimitedLenWrapper(MultiThreadedAugmenter) at MVDTrainer.py:652-663, without the worker processes.

    Batch i + 1 is uploaded and augmented on a side stream while the caller steps batch i (the chain is 1.6 ms for a cfg-2
    batch, HBM-bound, and hides under the step's tensor-pipe-bound kernels)."""
    dev = torch.device(device)
    side = torch.cuda.Stream(device=dev)

    def launch(raw):
        with torch.cuda.stream(side):
            d = torch.as_tensor(raw[data_key]).to(dev, non_blocking=True)
            s = torch.as_tensor(raw[seg_key]).to(dev, non_blocking=True)
            out = augmenter(d, s)
            ev = torch.cuda.Event()
            ev.record(side)
        return out, ev

    it = iter(raw_batches)
    try:
        nxt = launch(next(it))
    except StopIteration:
        return
    while nxt is not None:
        out, ev = nxt
        try:
            nxt = launch(next(it))
        except StopIteration:
            nxt = None
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(ev)
        tensors = [out['data']] + (list(out['target']) if isinstance(out['target'], (list, tuple)) else [out['target']])
        for t in tensors:
            t.record_stream(cur)
        yield out
