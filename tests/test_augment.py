"""GPU batch augmentation (multimodal_mvd_seg_b200/augment.py, csrc/augment.cu) against the numpy / scipy restatement of the
batchgenerators transforms the reference configures at MVDTrainer.py:700-765 (oracle/augment.py; parity unpinned: the
package is not part of the reference tree).  Every transform is compared given the SAME random draws."""
import numpy as np
import pytest
import torch

import oracle.augment as oa


def _rng(seed=0):
    return np.random.default_rng(seed)


# ---------------------------------------------------------------------------------------------------------------------
# host logic + oracle self-checks (CPU)
# ---------------------------------------------------------------------------------------------------------------------
def test_rotation_scale_matrix_matches_batchgenerators_convention():
    from multimodal_mvd_seg_b200.augment import rotation_scale_matrix
    m = rotation_scale_matrix(0.3, -0.2, 0.5, 1.2)
    np.testing.assert_allclose(m, oa.rotation_scale_matrix(0.3, -0.2, 0.5, 1.2), rtol=1e-12)
    np.testing.assert_allclose(m @ m.T, 1.44 * np.eye(3), atol=1e-12)           # scale * rotation
    # rotate_coords_3d multiplies coordinate ROWS with Rx Ry Rz: a pure x-rotation leaves axis 0 alone
    mx = rotation_scale_matrix(0.4, 0.0, 0.0, 1.0)
    np.testing.assert_allclose(mx[0], [1, 0, 0], atol=1e-12)
    np.testing.assert_allclose(mx[1:, 1:], [[np.cos(0.4), np.sin(0.4)], [-np.sin(0.4), np.cos(0.4)]], atol=1e-12)


def test_sample_parameters_follow_the_reference_probabilities():
    from multimodal_mvd_seg_b200.augment import sample_parameters
    rng = _rng(1)
    n = 4000
    P = sample_parameters(rng, n, 2)
    rate = lambda a: float(np.mean(a))
    assert abs(rate(P['mode'] == 1) - (1 - 0.8 * 0.8)) < 0.03                 # rotation or scaling drawn
    assert abs(rate(P['noise_sigma'][:, 0] > 0) - 0.1) < 0.02 and P['noise_sigma'].max() <= 0.1
    assert (P['noise_sigma'][:, 0] == P['noise_sigma'][:, 1]).all()           # one variance per sample
    assert abs(rate(P['blur_sigma'] > 0) - 0.2 * 0.5) < 0.02
    bs = P['blur_sigma'][P['blur_sigma'] > 0]
    assert bs.min() >= 0.5 and bs.max() <= 1.0
    assert abs(rate(P['brightness'][:, 0] != 1) - 0.15) < 0.02
    assert P['brightness'].min() >= 0.75 and P['brightness'].max() <= 1.25
    c = P['contrast'][P['contrast'] != 0]
    assert abs(rate(P['contrast'][:, 0] != 0) - 0.15) < 0.02 and c.min() >= 0.75 and c.max() <= 1.25
    assert abs(rate(c < 1) - 0.5) < 0.06                                      # below 1 with probability 1/2
    assert abs(rate(P['lowres_zoom'] != 0) - 0.25 * 0.5) < 0.02
    lz = P['lowres_zoom'][P['lowres_zoom'] != 0]
    assert lz.min() >= 0.5 and lz.max() <= 1.0
    assert abs(rate(P['gamma_inv'][:, 0] != 0) - 0.1) < 0.02 and abs(rate(P['gamma'][:, 0] != 0) - 0.3) < 0.03
    g = P['gamma'][P['gamma'] != 0]
    assert g.min() >= 0.7 and g.max() <= 1.5
    assert abs(rate(P['flips']) - 0.5) < 0.03
    # scaling factors: isotropic, in (0.7, 1.4); rotations within +-30 degrees
    det = np.linalg.det(P['mat'].astype(np.float64))
    sc = np.cbrt(det)
    assert sc.min() > 0.69 and sc.max() < 1.41
    none = sample_parameters(_rng(2), 64, 1, mirror_axes=())
    assert none['flips'].sum() == 0


def test_oracle_identity_is_the_centre_crop_and_gamma_keeps_the_statistics():
    rng = _rng(3)
    data = rng.normal(size=(2, 2, 12, 14, 10)).astype(np.float32)
    seg = rng.integers(-1, 4, size=(2, 1, 12, 14, 10)).astype(np.float32)
    patch = (8, 8, 8)
    mats = np.stack([np.eye(3)] * 2)
    crop, crop_seg = oa.spatial_transform(data, seg, patch, mats, [0, 0])
    np.testing.assert_array_equal(crop, data[:, :, 2:10, 3:11, 1:9])
    assert crop_seg.min() == 0                                                 # -1 removed
    same, same_seg = oa.spatial_transform(data, seg, patch, mats, [1, 1], order_data=1)
    np.testing.assert_allclose(same, crop, atol=1e-6)                          # identity transform = crop (even margins)
    np.testing.assert_array_equal(same_seg, crop_seg)
    x = rng.normal(2.0, 3.0, size=(3, 6, 7, 8)).astype(np.float32)
    y = oa.gamma(x, np.array([1.3, 0.0, 0.8], np.float32), invert=True)
    np.testing.assert_allclose(y.mean((1, 2, 3)), x.mean((1, 2, 3)), atol=2e-5)
    np.testing.assert_allclose(y.std((1, 2, 3)), x.std((1, 2, 3)), rtol=1e-5)
    np.testing.assert_array_equal(y[1], x[1])
    m = oa.mirror(oa.mirror(data, [[1, 0, 1], [0, 1, 0]]), [[1, 0, 1], [0, 1, 0]])
    np.testing.assert_array_equal(m, data)


def test_gpu_augmenter_refuses_cpu_tensors():
    from multimodal_mvd_seg_b200.augment import GpuAugmenter
    aug = GpuAugmenter((8, 8, 8), 4)
    with pytest.raises(RuntimeError, match='CUDA'):
        aug(torch.zeros(1, 1, 8, 8, 8), torch.zeros(1, 1, 8, 8, 8))
    with pytest.raises(NotImplementedError):
        GpuAugmenter((8, 8), 4)


# ---------------------------------------------------------------------------------------------------------------------
# kernels against the oracle (GPU)
# ---------------------------------------------------------------------------------------------------------------------
def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to('cuda:0')


def _st():
    return torch.cuda.current_stream().cuda_stream


@pytest.mark.gpu
def test_spline_prefilter_matches_scipy():
    import multimodal_mvd_seg_b200 as m
    from scipy.ndimage import spline_filter
    rng = _rng(10)
    x = rng.normal(size=(3, 9, 17, 12)).astype(np.float32)
    apply = np.array([1, 0, 1], np.uint8)
    t = _cuda(x)
    apply_d = _cuda(apply)
    m.lib.aug_spline_prefilter(t.data_ptr(), 3, 9, 17, 12, apply_d.data_ptr(), _st())
    got = t.cpu().numpy()
    np.testing.assert_array_equal(got[1], x[1])
    for p in (0, 2):
        want = spline_filter(x[p].astype(np.float64), order=3, mode='mirror')
        np.testing.assert_allclose(got[p], want, rtol=2e-5, atol=2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize('order', [3, 1])
def test_spatial_transform_matches_map_coordinates(order):
    import multimodal_mvd_seg_b200 as m
    rng = _rng(11)
    B, C, shape, patch = 3, 2, (30, 34, 26), (20, 24, 16)
    # smooth images (what cubic interpolation is for) + label blobs incl. the -1 "outside" label
    zz, yy, xx = np.meshgrid(*[np.linspace(-1, 1, s) for s in shape], indexing='ij')
    data = np.stack([np.stack([np.sin(3 * zz + b) * np.cos(2 * yy * (c + 1)) + 0.3 * xx for c in range(C)]) for b in range(B)])
    data = (data + 0.05 * rng.normal(size=data.shape)).astype(np.float32)
    seg = np.zeros((B, 1) + shape, np.float32)
    seg[:, :, 5:20, 8:28, 4:18] = 1
    seg[:, :, 10:16, 12:20, 8:14] = 3
    seg[:, :, 22:, :, :] = 2
    seg[:, :, :, :3, :] = -1
    mats = np.stack([oa.rotation_scale_matrix(0.4, -0.3, 0.2, 0.8), np.eye(3), oa.rotation_scale_matrix(-0.5, 0.1, 0.45, 1.35)])
    modes = np.array([1, 0, 1], np.int32)
    want, want_seg = oa.spatial_transform(data, seg, patch, mats, modes, order_data=order)
    src = _cuda(data)
    if order == 3:
        apply_d = _cuda(np.repeat(modes.astype(np.uint8), C))
        m.lib.aug_spline_prefilter(src.data_ptr(), B * C, *shape, apply_d.data_ptr(), _st())
    out = torch.empty((B, C) + patch, dtype=torch.float32, device='cuda:0')
    mat_d, mode_d = _cuda(mats.astype(np.float32).reshape(B, 9)), _cuda(modes)
    m.lib.aug_spatial(src.data_ptr(), B, C, *shape, out.data_ptr(), *patch, mat_d.data_ptr(), mode_d.data_ptr(), order, 0.0, 0,
                      _st())
    got = out.cpu().numpy()
    np.testing.assert_array_equal(got[1], want[1])                                       # centre crop: exact
    err = np.abs(got - want).max()
    assert err < (3e-4 if order == 3 else 2e-5), err
    assert (want[2] == 0).mean() > 0.01                                                  # the constant border is exercised (scale 1.35)
    sg = torch.empty((B, 1) + patch, dtype=torch.float32, device='cuda:0')
    seg_d = _cuda(seg)
    m.lib.aug_spatial(seg_d.data_ptr(), B, 1, *shape, sg.data_ptr(), *patch, mat_d.data_ptr(), mode_d.data_ptr(), 1, -1.0,
                      4, _st())
    got_seg = sg.cpu().numpy()
    agree = (got_seg == want_seg).mean()
    assert agree > 0.9995, agree                                                         # ties at exactly 0.5 may differ
    assert set(np.unique(got_seg)) <= {0.0, 1.0, 2.0, 3.0}


@pytest.mark.gpu
def test_intensity_transforms_match_oracle():
    import multimodal_mvd_seg_b200 as m
    lib = m.lib
    rng = _rng(12)
    N, D, H, W = 4, 10, 14, 18
    V = D * H * W
    x = rng.normal(1.0, 2.0, size=(N, D, H, W)).astype(np.float32)

    def stats(t):
        out = torch.tensor([0.0, 0.0, float('inf'), float('-inf')], dtype=torch.float64, device='cuda:0').repeat(N, 1).contiguous()
        lib.aug_plane_stats(t.data_ptr(), V, N, out.data_ptr(), _st())
        return out

    t = _cuda(x)
    s = stats(t).cpu().numpy()
    np.testing.assert_allclose(s[:, 0], x.astype(np.float64).sum((1, 2, 3)), rtol=1e-10)
    np.testing.assert_allclose(s[:, 1], (x.astype(np.float64) ** 2).sum((1, 2, 3)), rtol=1e-10)
    np.testing.assert_array_equal(s[:, 2], x.min((1, 2, 3)))
    np.testing.assert_array_equal(s[:, 3], x.max((1, 2, 3)))
    # blur
    sigma = np.array([0.7, 0.0, 1.0, 0.5], np.float32)
    t = _cuda(x)
    tmp, sig_d = torch.empty_like(t), _cuda(sigma)
    lib.aug_gaussian_blur(t.data_ptr(), tmp.data_ptr(), N, D, H, W, sig_d.data_ptr(), _st())
    np.testing.assert_allclose(t.cpu().numpy(), oa.gaussian_blur(x, sigma), rtol=1e-4, atol=2e-5)
    # brightness
    mult = np.array([0.8, 1.0, 1.2, 1.1], np.float32)
    t = _cuda(x)
    mult_d = _cuda(mult)
    lib.aug_intensity(t.data_ptr(), V, N, 0, mult_d.data_ptr(), None, None, 0, _st())
    np.testing.assert_allclose(t.cpu().numpy(), oa.brightness_multiplicative(x, mult), rtol=1e-6)
    # contrast
    fac = np.array([0.8, 0.0, 1.2, 1.0], np.float32)
    t = _cuda(x)
    fac_d, st_d = _cuda(fac), stats(t)
    lib.aug_intensity(t.data_ptr(), V, N, 1, fac_d.data_ptr(), st_d.data_ptr(), None, 0, _st())
    np.testing.assert_allclose(t.cpu().numpy(), oa.contrast(x, fac), rtol=1e-5, atol=1e-5)
    # gamma, both variants
    for invert in (1, 0):
        g = np.array([0.75, 1.4, 0.0, 1.1], np.float32)
        t = _cuda(x)
        gd = _cuda(g)
        s0 = stats(t)
        lib.aug_intensity(t.data_ptr(), V, N, 2, gd.data_ptr(), s0.data_ptr(), None, invert, _st())
        s1 = stats(t)
        lib.aug_intensity(t.data_ptr(), V, N, 3, gd.data_ptr(), s0.data_ptr(), s1.data_ptr(), invert, _st())
        want = oa.gamma(x, g, invert=bool(invert))
        np.testing.assert_allclose(t.cpu().numpy(), want, rtol=2e-4, atol=2e-4)
        np.testing.assert_array_equal(t.cpu().numpy()[2], x[2])
    # mirror
    xb = x.reshape(2, 2, D, H, W)
    flips = np.array([[1, 0, 1], [0, 1, 1]], np.uint8)
    out = torch.empty((2, 2, D, H, W), dtype=torch.float32, device='cuda:0')
    xb_d, flips_d = _cuda(xb), _cuda(flips)
    lib.aug_mirror(xb_d.data_ptr(), out.data_ptr(), 2, 2, D, H, W, flips_d.data_ptr(), _st())
    np.testing.assert_array_equal(out.cpu().numpy(), oa.mirror(xb, flips))
    # noise: N(0, sigma) on the selected planes only, different every voxel, reproducible per seed
    sig = np.array([0.05, 0.0, 0.1, 0.0], np.float32)
    t = torch.zeros((N, 40, 40, 40), dtype=torch.float32, device='cuda:0')
    sig_d = _cuda(sig)
    lib.aug_gaussian_noise(t.data_ptr(), 64000, N, sig_d.data_ptr(), 1234, _st())
    n = t.cpu().numpy()
    assert n[1].any() == False and n[3].any() == False
    for p in (0, 2):
        assert abs(n[p].mean()) < 4 * sig[p] / np.sqrt(64000) and abs(n[p].std() / sig[p] - 1) < 0.02
        assert abs(float(np.mean(np.abs(n[p]) > 2 * sig[p])) - 0.0455) < 0.005          # Gaussian tails
    t2 = torch.zeros_like(t)
    lib.aug_gaussian_noise(t2.data_ptr(), 64000, N, sig_d.data_ptr(), 1234, _st())
    assert torch.equal(t, t2)


@pytest.mark.gpu
def test_simulate_lowres_matches_scipy_zoom():
    import multimodal_mvd_seg_b200 as m
    rng = _rng(14)
    N, D, H, W = 4, 20, 26, 18
    zz, yy, xx = np.meshgrid(*[np.linspace(-1, 1, s) for s in (D, H, W)], indexing='ij')
    x = np.stack([np.sin(4 * zz + p) * np.cos(3 * yy) + xx * (p - 1) for p in range(N)]).astype(np.float32)
    x += 0.1 * rng.normal(size=x.shape).astype(np.float32)
    zoom = np.array([0.5, 0.0, 0.77, 1.0], np.float32)
    want = oa.simulate_lowres(x, zoom)
    tshape = np.zeros((N, 3), np.int32)
    for p in range(N):
        if zoom[p]:
            tshape[p] = np.round(np.array([D, H, W]) * float(zoom[p])).astype(np.int32)
    t = _cuda(x)
    stride = (D + 24) * (H + 24) * (W + 24)
    buf = torch.empty((N, stride), dtype=torch.float32, device='cuda:0')
    mm = torch.tensor([float('inf'), float('-inf')], dtype=torch.float64, device='cuda:0').repeat(N, 1).contiguous()
    ts = _cuda(tshape)
    m.lib.aug_simulate_lowres(t.data_ptr(), N, D, H, W, ts.data_ptr(), buf.data_ptr(), stride, mm.data_ptr(), _st())
    got = t.cpu().numpy()
    np.testing.assert_array_equal(got[1], x[1])
    assert np.abs(got - want).max() < 3e-4
    np.testing.assert_allclose(got[3], x[3], atol=3e-4)          # zoom 1: cubic interpolation at the samples = identity
    assert np.abs(want[0] - x[0]).max() > 0.05                   # ... and zoom 0.5 is not


@pytest.mark.gpu
def test_gpu_augmenter_pipeline_matches_the_chained_oracle():
    from multimodal_mvd_seg_b200.augment import GpuAugmenter, sample_parameters
    rng = _rng(13)
    B, C, shape, patch = 4, 2, (40, 44, 36), (32, 32, 24)
    zz, yy, xx = np.meshgrid(*[np.linspace(-1, 1, s) for s in shape], indexing='ij')
    data = np.stack([np.stack([np.cos(2 * zz * (b + 1)) * np.sin(3 * yy + c) + 0.5 * xx * zz for c in range(C)]) for b in range(B)])
    data = (data + 0.02 * rng.normal(size=data.shape)).astype(np.float32)
    seg = (rng.random((B, 1) + shape) < 0.02).astype(np.float32)
    seg[:, :, 10:30, 12:30, 8:28] = 2
    seg[:, :, 14:22, 16:24, 12:20] = 1
    P = sample_parameters(_rng(5), B, C)
    # force every branch at least once (noise stays off: the generators differ by construction)
    P['mode'][:] = [1, 0, 1, 0]
    P['mat'][0] = oa.rotation_scale_matrix(0.3, 0.2, -0.4, 0.9)
    P['mat'][2] = oa.rotation_scale_matrix(-0.2, 0.5, 0.1, 1.3)
    P['noise_sigma'][:] = 0
    P['blur_sigma'][:] = [[0.6, 0.0], [0.0, 0.9], [0.0, 0.0], [0.8, 0.7]]
    P['brightness'][1] = [0.8, 1.2]
    P['contrast'][:] = [[1.2, 0.8], [0, 0], [0.9, 1.1], [0, 0]]
    P['lowres_zoom'][:] = [[0.0, 0.6], [0.83, 0.0], [0.0, 0.0], [0.5, 1.0]]
    P['gamma_inv'][:] = [[0, 0], [1.3, 0.8], [0, 0], [0, 0]]
    P['gamma'][:] = [[0.75, 1.4], [0, 0], [0, 0], [1.2, 0.9]]
    P['flips'][:] = [[1, 0, 0], [0, 1, 1], [0, 0, 0], [1, 1, 1]]
    scales = [[1, 1, 1], [0.5, 0.5, 0.5], [0.25, 0.25, 0.25]]
    aug = GpuAugmenter(patch, 4, deep_supervision_scales=scales)
    out = aug(_cuda(data), _cuda(seg), params=P)
    # the same chain on the CPU
    x, s = oa.spatial_transform(data, seg, patch, P['mat'], P['mode'])
    x = x.reshape((B * C,) + patch)
    x = oa.gaussian_blur(x, P['blur_sigma'].reshape(-1))
    x = oa.brightness_multiplicative(x, P['brightness'].reshape(-1))
    x = oa.contrast(x, P['contrast'].reshape(-1))
    x = oa.simulate_lowres(x, P['lowres_zoom'].reshape(-1))
    x = oa.gamma(x, P['gamma_inv'].reshape(-1), invert=True)
    x = oa.gamma(x, P['gamma'].reshape(-1), invert=False)
    x = oa.mirror(x.reshape((B, C) + patch), P['flips'])
    s = oa.mirror(s, P['flips'])
    got = out['data'].cpu().numpy()
    assert np.abs(got - x).max() < 2e-3 * max(1.0, float(np.abs(x).max()))
    tg = out['target']
    assert [tuple(t.shape[2:]) for t in tg] == [patch, (16, 16, 12), (8, 8, 6)]
    assert (tg[0].cpu().numpy() == s).mean() > 0.9995
    # random parameters end to end: runs, shapes, finite values, labels in range
    aug2 = GpuAugmenter(patch, 4, seed=7)
    for _ in range(3):
        o = aug2(_cuda(data), _cuda(seg))
        assert o['data'].shape == (B, C) + patch and bool(torch.isfinite(o['data']).all())
        assert float(o['target'].min()) >= 0 and float(o['target'].max()) <= 3
    # throughput of the chain with every transform switched on, BASELINE cfg-2 sized batch (2 x 2 x 128^3 from 2 x 2 x 160^3)
    big = torch.randn((2, 2, 160, 160, 160), device='cuda:0')
    bseg = (torch.rand((2, 1, 160, 160, 160), device='cuda:0') < 0.1).float()
    Pb = {k: v[:2].copy() for k, v in P.items()}
    Pb['noise_sigma'][:] = 0.05
    Pb['lowres_zoom'][:] = [[0.7, 0.9], [0.55, 0.8]]
    augb = GpuAugmenter((128, 128, 128), 4, deep_supervision_scales=scales)
    augb(big, bseg, params=Pb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        augb(big, bseg, params=Pb)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print('GPU augmentation, all transforms on, 2 x 2 x 160^3 -> 128^3: %.2f ms per batch (%.0f patches/s)' % (ms, 2000.0 / ms))
    assert ms < 50


@pytest.mark.gpu
def test_augmented_batches_runs_one_batch_ahead_and_feeds_train_step():
    """raw loader batches -> side-stream augmentation -> train_step (device batches): same tensors as the direct call with
    an identically seeded augmenter, and a trainer steps on them."""
    import multimodal_mvd_seg_b200 as m
    from multimodal_mvd_seg_b200.augment import GpuAugmenter, augmented_batches
    rng = _rng(21)
    patch, shape = (32, 32, 32), (40, 40, 40)
    plans, dj = m.make_plans(patch, batch_size=2, n_modalities=2, n_classes=4)
    tr = m.nnUNetTrainer(plans, '3d_fullres', 0, dj, device=torch.device('cuda:0'))
    tr.initialize()
    scales = tr._get_deep_supervision_scales() if hasattr(tr, '_get_deep_supervision_scales') else None
    assert scales is not None
    raws = [{'data': rng.normal(size=(2, 2) + shape).astype(np.float32),
             'seg': rng.integers(-1, 4, size=(2, 1) + shape).astype(np.float32)} for _ in range(3)]
    direct = GpuAugmenter(patch, 4, deep_supervision_scales=scales, seed=3)
    want = [direct(_cuda(r['data']), _cuda(r['seg'])) for r in raws]
    torch.cuda.synchronize()
    piped = GpuAugmenter(patch, 4, deep_supervision_scales=scales, seed=3)
    tr.on_train_epoch_start()
    n = 0
    for got, ref in zip(augmented_batches(raws, piped, 'cuda:0'), want):
        assert torch.equal(got['data'], ref['data'])
        assert all(torch.equal(a, b) for a, b in zip(got['target'], ref['target']))
        out = tr.train_step(got)
        assert np.isfinite(out['loss'])
        n += 1
    assert n == 3
