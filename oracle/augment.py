"""CPU restatement (numpy / scipy) of the training augmentations the reference applies to every batch
(nnUNetTrainer.get_training_transforms, MVDTrainer.py:700-765).  TEST INFRASTRUCTURE ONLY: imported by tests/ and nothing
else; the product path (multimodal_mvd_seg_b200/augment.py -> csrc/augment.cu) never touches it.

PARITY UNPINNED: the transforms are classes of the third-party package ``batchgenerators`` (a pip dependency of nnunetv2,
absent from /root/reference and from this image), so there is neither reference source nor a reference test vector to
pin against.  What is restated here is the published algorithm of batchgenerators 0.25 (the version nnunetv2 2.1 pins):
  augment_spatial          batchgenerators/augmentations/spatial_transformations.py  (zero-centred mesh, rotate_coords_3d,
                           scale_coords, interpolate_img -> scipy.ndimage.map_coordinates)
  augment_gaussian_noise   .../noise_augmentations.py          augment_gaussian_blur   .../noise_augmentations.py
  augment_brightness_multiplicative / augment_contrast / augment_gamma   .../color_augmentations.py
  augment_mirroring        .../spatial_transformations.py
  augment_linear_downsampling_scipy   .../resample_augmentations.py  (skimage.transform.resize, itself scipy.ndimage.zoom(
                           grid_mode=True) + clipping in skimage >= 0.19; skimage is not in this image either)
Every function takes the random draws as explicit arguments (the call sites in MVDTrainer.py fix the distributions, see
``multimodal_mvd_seg_b200.augment.sample_parameters``), which turns each transform into a deterministic function the CUDA
kernels can be compared with."""
import numpy as np
from scipy.ndimage import gaussian_filter, map_coordinates, zoom as _zoom


def rotation_scale_matrix(angle_x: float, angle_y: float, angle_z: float, scale: float) -> np.ndarray:
    """coords_new = M @ coords for rotate_coords_3d followed by scale_coords: batchgenerators multiplies the (n, 3) coordinate
    rows with R = Rx @ Ry @ Rz from the right, i.e. M = scale * R^T."""
    cx, sx, cy, sy, cz, sz = np.cos(angle_x), np.sin(angle_x), np.cos(angle_y), np.sin(angle_y), np.cos(angle_z), np.sin(angle_z)
    rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]], dtype=np.float64)
    ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]], dtype=np.float64)
    rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]], dtype=np.float64)
    return float(scale) * (rx @ ry @ rz).T


def _mesh(patch, in_shape, mat):
    """create_zero_centered_coordinate_mesh -> M -> + centre of the source volume (random_crop = False: ctr = in / 2 - 0.5)"""
    axes = [np.arange(p, dtype=np.float64) - (p - 1) / 2.0 for p in patch]
    grid = np.stack(np.meshgrid(*axes, indexing='ij'))
    coords = (mat @ grid.reshape(3, -1)).reshape(grid.shape)
    for d in range(3):
        coords[d] += in_shape[d] / 2.0 - 0.5
    return coords


def spatial_transform(data, seg, patch, mats, modes, order_data=3, order_seg=1, cval_data=0.0, n_labels=None):
    """SpatialTransform as configured at MVDTrainer.py:700-711 (no elastic deformation, constant border, border_cval_seg = -1)
    followed by RemoveLabelTransform(-1, 0) (:738).  data [B,C,Di,Hi,Wi], seg [B,1,Di,Hi,Wi]; mats [B,3,3]; modes [B]: 0 = the
    sample drew neither rotation nor scaling -> centre crop, 1 = interpolate."""
    B, C = data.shape[:2]
    in_shape = data.shape[2:]
    out = np.zeros((B, C) + tuple(patch), dtype=np.float32)
    out_seg = np.zeros((B, seg.shape[1]) + tuple(patch), dtype=np.float32)
    for b in range(B):
        if modes[b] == 0:
            lb = [(in_shape[d] - patch[d]) // 2 for d in range(3)]
            sl = tuple(slice(lb[d], lb[d] + patch[d]) for d in range(3))
            out[b] = data[b][(slice(None),) + sl]
            out_seg[b] = seg[b][(slice(None),) + sl]
        else:
            coords = _mesh(patch, in_shape, np.asarray(mats[b], dtype=np.float64))
            for c in range(C):
                out[b, c] = map_coordinates(data[b, c].astype(np.float32), coords, order=order_data, mode='constant',
                                            cval=cval_data).astype(np.float32)
            for c in range(seg.shape[1]):
                s = seg[b, c]
                res = np.zeros(tuple(patch), dtype=np.float32)
                for lab in np.sort(np.unique(s)):          # interpolate_img(is_seg=True): per label, ascending
                    m = map_coordinates((s == lab).astype(np.float32), coords, order=order_seg, mode='constant', cval=-1.0)
                    res[m >= 0.5] = lab
                out_seg[b, c] = res
    out_seg[out_seg == -1] = 0                            # RemoveLabelTransform(-1, 0)
    return out, out_seg


def gaussian_blur(x, sigma):
    """augment_gaussian_blur: per plane scipy gaussian_filter(order=0) with the drawn sigma (0 = plane not selected)"""
    y = x.copy()
    for p in range(x.shape[0]):
        if sigma[p] > 0:
            y[p] = gaussian_filter(x[p], float(sigma[p]), order=0)
    return y


def brightness_multiplicative(x, mult):
    return (x * np.asarray(mult, dtype=np.float32).reshape(-1, 1, 1, 1)).astype(np.float32)


def contrast(x, factor):
    """augment_contrast(preserve_range=True, per_channel=True); factor 0 = plane not selected"""
    y = x.copy()
    for p in range(x.shape[0]):
        if factor[p] != 0:
            mn, lo, hi = x[p].mean(), x[p].min(), x[p].max()
            y[p] = np.clip((x[p] - mn) * np.float32(factor[p]) + mn, lo, hi)
    return y


def gamma(x, g, invert, retain_stats=True, epsilon=1e-7):
    """augment_gamma(per_channel=True): g 0 = plane not selected"""
    y = x.astype(np.float32).copy()
    for p in range(x.shape[0]):
        if g[p] == 0:
            continue
        d = -y[p] if invert else y[p]
        if retain_stats:
            mn, sd = d.mean(), d.std()
        minm = d.min()
        rnge = d.max() - minm
        d = np.power((d - minm) / float(rnge + epsilon), np.float32(g[p])) * rnge + minm
        if retain_stats:
            d = d - d.mean()
            d = d / (d.std() + 1e-8) * sd
            d = d + mn
        y[p] = -d if invert else d
    return y


def mirror(x, flips):
    """augment_mirroring: x [B,C,D,H,W], flips [B,3]"""
    y = x.copy()
    for b in range(x.shape[0]):
        for ax in range(3):
            if flips[b][ax]:
                y[b] = np.flip(y[b], axis=1 + ax)
    return np.ascontiguousarray(y)


def simulate_lowres(x, zoom_factor):
    """augment_linear_downsampling_scipy(order_downsample=0, order_upsample=3, per_channel=True): per plane with
    zoom_factor != 0: resize to round(shape * zoom) with nearest neighbours, resize back with cubic splines; resize =
    skimage.transform.resize(mode='edge', anti_aliasing=False, clip=True) = scipy zoom(grid_mode=True, mode='nearest')
    clipped to the range of ITS input."""
    y = x.astype(np.float32).copy()
    for p in range(x.shape[0]):
        if zoom_factor[p] == 0:
            continue
        shp = np.array(x[p].shape)
        tgt = np.round(shp * float(zoom_factor[p])).astype(int)
        vol = x[p].astype(np.float64)
        down = _zoom(vol, tgt / shp, order=0, mode='nearest', grid_mode=True)
        down = np.clip(down, vol.min(), vol.max())
        up = _zoom(down, shp / tgt, order=3, mode='nearest', grid_mode=True)
        y[p] = np.clip(up, down.min(), down.max()).astype(np.float32)
    return y
