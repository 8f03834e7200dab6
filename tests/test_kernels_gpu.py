"""GPU parity tests, kernel by kernel, through the C ABI (ctypes) against the oracle / PyTorch fp32 math on the same
bf16-rounded inputs.  Tolerances (BASELINE.json north_star): fp32 loss reductions 1e-5 relative, bf16 tensors 2e-2
relative (norm-wise: ||got - want||_2 <= 2e-2 * ||want||_2), integer/morphology outputs bit-exact."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF = torch.bfloat16


def rel_err(got, want):
    got, want = got.detach().double().flatten(), want.detach().double().flatten()
    return float((got - want).norm() / want.norm().clamp_min(1e-30))


@pytest.fixture(scope='module')
def m():
    import multimodal_mvd_seg_b200 as mod
    assert torch.cuda.is_available()
    return mod


def dev():
    return torch.device('cuda:0')


def rand_cl(shape, seed, scale=1.0):
    g = torch.Generator(device='cpu').manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(BF).to(dev())


# ------------------------------------------------------------------------------------------------------------------
def test_layout_roundtrip(m):
    x = torch.randn(2, 2, 5, 6, 7, device=dev())
    cl = m.ops.input_to_cl(x)
    assert cl.shape == (2, 5, 6, 7, 2) and cl.dtype == BF
    assert torch.equal(cl, x.permute(0, 2, 3, 4, 1).to(BF))
    back = m.ops.cl_to_ncdhw_f32(cl)
    assert torch.equal(back, x.to(BF).float())


CONV_CASES = [
    # (B, Cin, Cout, (D,H,W), k, s)
    (2, 32, 32, (8, 8, 8), 3, (1, 1, 1)),
    (1, 32, 64, (8, 12, 16), 3, (2, 2, 2)),
    (1, 64, 32, (6, 10, 12), 3, (2, 2, 1)),
    (2, 2, 32, (9, 8, 10), 3, (1, 1, 1)),     # stem: Cin = 2
    (1, 1, 32, (8, 8, 8), 3, (1, 1, 1)),      # MVD stem: Cin = 1
    (1, 96, 40, (5, 5, 6), 3, (1, 1, 1)),     # odd extents, channels not multiples of 32
    (1, 16, 24, (4, 4, 4), 1, (1, 1, 1)),
    # tcgen05 coverage: several K chunks, N tiles of 128 / 160 / 256, tails in every direction, tiny volumes
    (2, 64, 64, (16, 16, 16), 3, (1, 1, 1)),
    (1, 128, 128, (9, 17, 11), 3, (1, 1, 1)),
    (1, 256, 320, (4, 4, 4), 3, (1, 1, 1)),
    (1, 640, 320, (5, 5, 6), 3, (1, 1, 1)),
    (1, 512, 256, (8, 8, 8), 3, (1, 1, 1)),
    (1, 320, 320, (8, 8, 8), 3, (2, 2, 2)),
    (1, 320, 320, (10, 10, 6), 3, (2, 2, 1)),
    # depth-folded halo kernel (resident weights): K/N in {32, 64}, depth tails (D % MT != 0), D < MT
    (1, 64, 32, (9, 17, 11), 3, (1, 1, 1)),
    (1, 32, 64, (5, 6, 7), 3, (1, 1, 1)),
    (2, 32, 32, (19, 8, 16), 3, (1, 1, 1)),
    (1, 32, 32, (3, 20, 9), 3, (1, 1, 1)),
    # sub-pixel strided dgrad (Cin = 32, Cout = 64, s = 2): odd extents, several bricks, depth runs with tails
    (2, 32, 64, (9, 14, 18), 3, (2, 2, 2)),
    (1, 32, 64, (22, 34, 16), 3, (2, 2, 2)),
]


def _conv_ref(x_cl, w, b, k, s):
    x = x_cl.float().permute(0, 4, 1, 2, 3)
    y = F.conv3d(x, w.to(BF).float(), None if b is None else b.to(BF).float(), stride=s, padding=(k - 1) // 2)
    return y


@pytest.mark.parametrize('case', CONV_CASES)
@pytest.mark.parametrize('algo', ['generic', 'auto'])
def test_conv_fprop_dgrad_wgrad(m, case, algo):
    B, Cin, Cout, (D, H, W), k, s = case
    ops = m.ops
    ops.set_conv_algo(algo)
    try:
        torch.manual_seed(1)
        x = rand_cl((B, D, H, W, Cin), 3)
        w = (torch.randn(Cout, Cin, k, k, k) * (1.0 / np.sqrt(Cin * k ** 3))).to(dev())
        b = torch.randn(Cout).to(dev()) * 0.1
        geom = ops.ConvGeom((k,) * 3, s, ((k - 1) // 2,) * 3)
        Do, Ho, Wo = geom.out_size((D, H, W))
        wf, wd = ops.pack_weights(w)
        # fprop (pitched output: channel slice of a wider buffer)
        ybuf = torch.zeros((B, Do, Ho, Wo, Cout + 8), dtype=BF, device=dev())
        y = ybuf[..., 8:]
        stats = torch.zeros((B, Cout, 2), dtype=torch.float64, device=dev()) if Cout % 8 == 0 else None
        ops.conv_fprop(geom, x, y, wf, bias=b, stats=stats)
        if stats is not None:   # InstanceNorm sums of the bf16-rounded output (fused into the tcgen05 epilogue)
            yd = y.double().reshape(B, -1, Cout)
            np.testing.assert_allclose(stats[..., 0].cpu(), yd.sum(1).cpu(), rtol=1e-4, atol=1e-3)
            np.testing.assert_allclose(stats[..., 1].cpu(), (yd * yd).sum(1).cpu(), rtol=1e-4, atol=1e-3)
        xr = x.float().permute(0, 4, 1, 2, 3).requires_grad_(True)
        wr = w.to(BF).float().requires_grad_(True)
        yr = F.conv3d(xr, wr, b.to(BF).float(), stride=s, padding=(k - 1) // 2)
        assert rel_err(y.float().permute(0, 4, 1, 2, 3), yr) < 1e-2
        assert float(ybuf[..., :8].abs().max()) == 0.0
        # backward operands
        dy = rand_cl((B, Do, Ho, Wo, Cout), 5)
        yr.backward(dy.float().permute(0, 4, 1, 2, 3))
        dx = torch.empty_like(x)
        ops.conv_dgrad(geom, dx, dy, wd)
        assert rel_err(dx.float().permute(0, 4, 1, 2, 3), xr.grad) < 1e-2
        # accumulate mode
        dx2 = dx.clone()
        ops.conv_dgrad(geom, dx2, dy, wd, accumulate=True)
        assert rel_err(dx2.float(), 2 * dx.float()) < 1e-2
        dw = torch.empty_like(w)
        db = torch.empty_like(b)
        ops.conv_wgrad(geom, x, dy, dw, db)
        assert rel_err(dw, wr.grad) < 1e-2
        assert rel_err(db, dy.float().sum((0, 1, 2, 3))) < 1e-3
    finally:
        ops.set_conv_algo('auto')


@pytest.mark.parametrize('cin,cout,E,s,k', [(32, 32, 40, 1, 3), (64, 64, 24, 1, 3), (32, 64, 24, 2, 3), (320, 320, 8, 1, 3),
                                            (64, 32, 16, 2, 2)])
def test_wgrad_deterministic_two_stage_reduction(m, cin, cout, E, s, k):
    """deterministic reduction mode (mvd_set_deterministic(1)): every split of the voxel range stores its partial into its
    own workspace slice and the slices are added in a fixed order -> dw is bit-reproducible from run to run (and equals
    the default atomics mode up to fp32 summation order)."""
    ops = m.ops
    g = torch.Generator().manual_seed(5)
    p = (k - 1) // 2 if k == 3 else 0
    geom = ops.ConvGeom((k,) * 3, (s,) * 3, (p,) * 3)
    Eo = geom.out_size((E, E, E))[0]
    x = torch.randn((2, E, E, E, cin), generator=g).to(BF).to(dev())
    dy = torch.randn((2, Eo, Eo, Eo, cout), generator=g).to(BF).to(dev())
    runs = []
    m.lib.set_deterministic(1)
    try:
        assert m.lib.get_deterministic() == 1
        for _ in range(3):
            dw = torch.empty((cout, cin, k, k, k), dtype=torch.float32, device=dev())
            ops.conv_wgrad(geom, x, dy, dw)
            runs.append(dw)
            torch.randn((1 << 22,), device=dev()).sum()     # perturb the timing between the runs
    finally:
        m.lib.set_deterministic(0)
    assert torch.equal(runs[0], runs[1]) and torch.equal(runs[0], runs[2])
    dwa = torch.empty_like(runs[0])
    ops.conv_wgrad(geom, x, dy, dwa)                        # default mode: vector atomics
    assert rel_err(dwa, runs[0]) < 1e-5
    ref = torch.nn.grad.conv3d_weight(x.float().permute(0, 4, 1, 2, 3), (cout, cin, k, k, k),
                                      dy.float().permute(0, 4, 1, 2, 3), stride=s, padding=p)
    assert rel_err(runs[0], ref) < 1e-3


@pytest.mark.parametrize('cin,cout,shape,s', [(320, 320, (2, 8, 8, 8), 1), (320, 320, (2, 4, 4, 4), 1), (256, 320, (2, 16, 16, 16), 2),
                                             (640, 320, (1, 5, 5, 6), 1), (320, 320, (2, 10, 10, 6), 1)])
def test_small_lattice_split_k_is_reproducible_and_matches_torch(m, cin, cout, shape, s):
    """layers whose produced lattice is small run split-K (K-sliced work units, sliced fp32 partials, ordered finishing
    sum): fprop / dgrad bit-reproducible, equal to the unsplit kernel up to fp32 summation order, and to fp32 F.conv3d."""
    import torch.nn.functional as F
    ops = m.ops
    g = torch.Generator().manual_seed(6)
    geom = ops.ConvGeom((3,) * 3, (s,) * 3, (1,) * 3)
    B, D, H, W = shape
    Do, Ho, Wo = geom.out_size((D, H, W))
    x = torch.randn((B, D, H, W, cin), generator=g).to(BF).to(dev())
    dy = torch.randn((B, Do, Ho, Wo, cout), generator=g).to(BF).to(dev())
    w = (torch.randn((cout, cin, 3, 3, 3), generator=g) / np.sqrt(cin * 27)).to(dev())
    bias = torch.randn((cout,), generator=g).to(dev())
    wf, wd = ops.pack_weights(w)
    ys = [torch.empty_like(dy) for _ in range(2)]
    for y in ys:
        ops.conv_fprop(geom, x, y, wf, bias=bias)
    assert torch.equal(ys[0], ys[1])
    ref = F.conv3d(x.float().permute(0, 4, 1, 2, 3), w.to(BF).float(), bias.to(BF).float(), stride=s, padding=1)
    assert rel_err(ys[0].float().permute(0, 4, 1, 2, 3), ref) < 1e-2
    dxs = [torch.empty_like(x) for _ in range(2)]
    for dx in dxs:
        ops.conv_dgrad(geom, dx, dy, wd)
    assert torch.equal(dxs[0], dxs[1])
    refdx = torch.nn.grad.conv3d_input(x.float().permute(0, 4, 1, 2, 3).shape, w.to(BF).float(),
                                       dy.float().permute(0, 4, 1, 2, 3), stride=s, padding=1)
    assert rel_err(dxs[0].float().permute(0, 4, 1, 2, 3), refdx) < 1e-2
    # accumulate into an existing gradient (the skip-connection fold)
    base = torch.randn(x.shape, generator=g).to(BF).to(dev())
    acc = base.clone()
    ops.conv_dgrad(geom, acc, dy, wd, accumulate=True)
    assert rel_err(acc.float(), (dxs[0].float() + base.float())) < 1e-2


@pytest.mark.parametrize('cout', [64, 32])
@pytest.mark.parametrize('shape', [(2, 7, 9, 11), (1, 18, 34, 20), (1, 2, 2, 2), (1, 9, 40, 16)])
def test_conv_s2_halo_fprop_pitched(m, shape, cout):
    """Stride-2 halo-plane forward (conv_halo_s2_kernel): the input is the second channel half of a wider buffer (the
    encoder skip lives inside the decoder's concat buffer), odd extents in every direction, depth tails, bit-for-bit
    equal to the tap-by-tap tcgen05 kernel up to accumulation order."""
    ops = m.ops
    B, D, H, W = shape
    xbuf = rand_cl((B, D, H, W, 64), 11)
    x = xbuf[..., 32:]
    w = (torch.randn(cout, 32, 3, 3, 3) * (1.0 / np.sqrt(32 * 27))).to(dev())
    b = torch.randn(cout).to(dev()) * 0.1
    geom = ops.ConvGeom((3,) * 3, (2,) * 3, (1,) * 3)
    Do, Ho, Wo = geom.out_size((D, H, W))
    wf, _ = ops.pack_weights(w)
    ybuf = torch.zeros((B, Do, Ho, Wo, cout + 8), dtype=BF, device=dev())
    y = ybuf[..., 8:]
    stats = torch.zeros((B, cout, 2), dtype=torch.float64, device=dev())
    ops.conv_fprop(geom, x, y, wf, bias=b, stats=stats)
    yr = F.conv3d(x.float().permute(0, 4, 1, 2, 3), w.to(BF).float(), b.to(BF).float(), stride=2, padding=1)
    assert rel_err(y.float().permute(0, 4, 1, 2, 3), yr) < 1e-2
    assert float(ybuf[..., :8].abs().max()) == 0.0
    yd = y.double().reshape(B, -1, cout)
    np.testing.assert_allclose(stats[..., 0].cpu(), yd.sum(1).cpu(), rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(stats[..., 1].cpu(), (yd * yd).sum(1).cpu(), rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize('cin,shape', [(1, (2, 5, 6, 7)), (2, (1, 9, 8, 10)), (2, (1, 3, 4, 70)), (2, (2, 4, 5, 16)), (1, (1, 3, 4, 24))])
def test_stem_im2col_bit_exact(m, cin, shape):
    """X_col[v][tap*Cin + ci] = x[v + tap - 1][ci] (zero outside the volume, zero padding columns): a pure gather."""
    B, D, H, W = shape
    x = rand_cl((B, D, H, W, cin), 21)
    kpad = 32 if cin == 1 else 64
    col = torch.full((B, D, H, W, kpad), 7.0, dtype=BF, device=dev())
    m.lib.im2col_small(x.data_ptr(), cin, B, D, H, W, cin, 3, 3, 3, 1, 1, 1, col.data_ptr(), kpad,
                       torch.cuda.current_stream().cuda_stream)
    xp = F.pad(x.float(), (0, 0, 1, 1, 1, 1, 1, 1))
    want = torch.zeros((B, D, H, W, kpad), device=dev())
    t = 0
    for td in range(3):
        for th in range(3):
            for tw in range(3):
                want[..., t * cin:(t + 1) * cin] = xp[:, td:td + D, th:th + H, tw:tw + W, :]
                t += 1
    assert torch.equal(col.float(), want)


@pytest.mark.parametrize('mode', ['fused', 'im2col'])
@pytest.mark.parametrize('cin,shape', [(2, (2, 9, 20, 13)), (1, (1, 6, 17, 24)), (2, (1, 3, 16, 8))])
def test_stem_block_matches_torch(m, mode, cin, shape):
    """first block (Conv3d(cin, 32, 3, padding=1) + InstanceNorm + LeakyReLU): tensor-core GEMM over the in-smem im2col
    tile (and the explicit X_col form) against torch, forward and all parameter gradients."""
    ops = m.ops
    ops.set_stem_mode(mode)
    try:
        B, D, H, W = shape
        x = rand_cl((B, D, H, W, cin), 41)
        w = (torch.randn(32, cin, 3, 3, 3) / np.sqrt(27 * cin)).to(dev()).requires_grad_(True)
        b = (torch.randn(32) * 0.1).to(dev()).requires_grad_(True)
        gam = (1 + 0.1 * torch.randn(32)).to(dev()).requires_grad_(True)
        bet = (0.1 * torch.randn(32)).to(dev()).requires_grad_(True)
        geom = ops.ConvGeom((3, 3, 3), (1, 1, 1), (1, 1, 1))
        z = ops.ConvNormActFn.apply(x, w, b, gam, bet, geom, 1e-5, 0.01, None, None)
        xr = x.float().permute(0, 4, 1, 2, 3)
        wr = w.detach().to(BF).float().requires_grad_(True)
        br = b.detach().to(BF).float().requires_grad_(True)
        gr = gam.detach().clone().requires_grad_(True)
        ber = bet.detach().clone().requires_grad_(True)
        yr = F.conv3d(xr, wr, br, padding=1).to(BF).float()
        zr = F.leaky_relu(F.instance_norm(yr, weight=gr, bias=ber, eps=1e-5), 0.01)
        assert rel_err(z.float().permute(0, 4, 1, 2, 3), zr) < 1e-2
        g = rand_cl(tuple(z.shape), 42)
        z.backward(g)
        zr.backward(g.float().permute(0, 4, 1, 2, 3))
        assert rel_err(w.grad, wr.grad) < 2e-2
        assert rel_err(gam.grad, gr.grad) < 2e-2 and rel_err(bet.grad, ber.grad) < 2e-2
    finally:
        ops.set_stem_mode('fused')


@pytest.mark.parametrize('stride', [(2, 2, 2), (2, 2, 1)])
@pytest.mark.parametrize('cin,cout', [(64, 32), (320, 320), (48, 24)])
def test_conv_transpose(m, stride, cin, cout):
    ops = m.ops
    x = rand_cl((2, 3, 4, 5, cin), 11)
    w = (torch.randn(cin, cout, *stride) * (1.0 / np.sqrt(cin))).to(dev()).requires_grad_(True)
    b = (torch.randn(cout) * 0.1).to(dev()).requires_grad_(True)
    xin = x.clone().requires_grad_(True)
    up = ops.ConvTransposeFn.apply(xin, w, b, stride, None, None)
    xr = x.float().permute(0, 4, 1, 2, 3).requires_grad_(True)
    wr = w.detach().to(BF).float().requires_grad_(True)
    br = b.detach().to(BF).float().requires_grad_(True)
    ur = F.conv_transpose3d(xr, wr, br, stride=stride)
    assert rel_err(up.float().permute(0, 4, 1, 2, 3), ur) < 1e-2
    g = rand_cl(tuple(up.shape), 12)
    up.backward(g)
    ur.backward(g.float().permute(0, 4, 1, 2, 3))
    assert rel_err(xin.grad.float().permute(0, 4, 1, 2, 3), xr.grad) < 1e-2
    assert rel_err(w.grad, wr.grad) < 1e-2
    assert rel_err(b.grad, br.grad) < 1e-3


@pytest.mark.parametrize('C,shape', [(32, (2, 8, 9, 10)), (320, (2, 4, 4, 4)), (64, (1, 16, 16, 8))])
def test_instance_norm_lrelu_fwd_bwd(m, C, shape):
    lib, ops = m.lib, m.ops
    B, D, H, W = shape
    V = D * H * W
    y = rand_cl((B, D, H, W, C), 21, scale=2.0) + 0.5
    gamma = (torch.rand(C) + 0.5).to(dev())
    beta = (torch.randn(C) * 0.2).to(dev())
    st = torch.cuda.current_stream().cuda_stream
    stats = torch.zeros((B, C, 2), dtype=torch.float64, device=dev())
    lib.inorm_stats(y.data_ptr(), C, B, V, C, stats.data_ptr(), st)
    yf = y.double().reshape(B, V, C)
    np.testing.assert_allclose(stats[..., 0].cpu(), yf.sum(1).cpu(), rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(stats[..., 1].cpu(), (yf * yf).sum(1).cpu(), rtol=1e-6)
    zbuf = torch.zeros((B, D, H, W, 2 * C), dtype=BF, device=dev())
    z = zbuf[..., C:]
    lib.inorm_lrelu_fwd(y.data_ptr(), C, z.data_ptr(), 2 * C, stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                        B, V, C, 1e-5, 0.01, st)
    # reference: fp32 instance norm -> bf16 -> leaky relu -> bf16, with autograd for the backward
    yr = y.float().permute(0, 4, 1, 2, 3).requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    t = F.instance_norm(yr, weight=gr, bias=br, eps=1e-5)
    zr = F.leaky_relu(t, 0.01)
    assert rel_err(z.float().permute(0, 4, 1, 2, 3), zr) < 6e-3   # two bf16 roundings
    assert float(zbuf[..., :C].abs().max()) == 0.0
    dz = rand_cl((B, D, H, W, C), 22)
    zr.backward(dz.float().permute(0, 4, 1, 2, 3))
    bstats = torch.zeros((B, C, 2), dtype=torch.float64, device=dev())
    lib.inorm_lrelu_bwd_stats(dz.data_ptr(), C, y.data_ptr(), C, stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(),
                              B, V, C, 1e-5, 0.01, bstats.data_ptr(), st)
    dy = torch.empty_like(y)
    dg, db = torch.empty_like(gamma), torch.empty_like(beta)
    dsum = torch.zeros_like(gamma)
    lib.inorm_lrelu_bwd_apply(dz.data_ptr(), C, y.data_ptr(), C, dy.data_ptr(), C, stats.data_ptr(),
                              bstats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), B, V, C, 1e-5, 0.01,
                              dg.data_ptr(), db.data_ptr(), dsum.data_ptr(), st)
    np.testing.assert_allclose(dsum.cpu(), dy.float().sum((0, 1, 2, 3)).cpu(), rtol=1e-3, atol=2e-2)
    assert rel_err(dy.float().permute(0, 4, 1, 2, 3), yr.grad) < 1e-2
    assert rel_err(dg, gr.grad) < 5e-3
    assert rel_err(db, br.grad) < 5e-3


@pytest.mark.parametrize('cin,cout,shape,accumulate,fused', [
    (32, 32, (2, 20, 24, 40), False, True),      # depth-folded kernel, two samples
    (32, 64, (2, 9, 17, 21), False, True),       # ragged bricks in h and w, MT tail in d
    (32, 32, (1, 16, 16, 24), True, True),       # accumulate into an existing gradient (6144 voxels: no split-K)
    (32, 32, (2, 3, 48, 32), False, True),       # short volume: MT = 2
    (64, 64, (1, 16, 32, 24), False, False),     # wider layers: the statistics pass runs inside the call
    (128, 128, (2, 12, 16, 16), True, False),
])
def test_dgrad_epilogue_forms_norm_backward_statistics(m, cin, cout, shape, accumulate, fused):
    """mvd_conv3d_args.norm_bwd: the data gradient call also leaves the InstanceNorm + LeakyReLU backward sums of the block
    in front of the produced tensor -- for 32 produced channels out of the conv epilogue (one launch) -- equal to
    mvd_inorm_lrelu_bwd_stats run over the gradient it wrote; the produced gradient is bit-identical to the plain call."""
    lib, ops = m.lib, m.ops
    g = torch.Generator().manual_seed(31)
    B, D, H, W = shape
    V = D * H * W
    geom = ops.ConvGeom((3,) * 3, (1,) * 3, (1,) * 3)
    dy = torch.randn((B, D, H, W, cout), generator=g).to(BF).to(dev())
    w = (torch.randn((cout, cin, 3, 3, 3), generator=g) / np.sqrt(cout * 27)).to(dev())
    _, wd = ops.pack_weights(w)
    # the block in front: raw conv output y_prev, its forward sums, affine parameters
    y_prev = (torch.randn((B, D, H, W, cin), generator=g) * 1.5 + 0.3).to(BF).to(dev())
    gamma = (torch.rand(cin, generator=g) - 0.3)                     # both signs ...
    gamma[3] = 0.0                                                   # ... and constant masks: beta > 0 / beta <= 0
    gamma[5] = 0.0
    beta = torch.randn(cin, generator=g) * 0.3
    beta[3], beta[5] = 0.25, -0.25
    gamma, beta = gamma.to(dev()), beta.to(dev())
    st = torch.cuda.current_stream().cuda_stream
    stats = torch.zeros((B, cin, 2), dtype=torch.float64, device=dev())
    lib.inorm_stats(y_prev.data_ptr(), cin, B, V, cin, stats.data_ptr(), st)
    nb = (y_prev, stats, gamma, beta, 1e-5, 0.01)
    base = torch.randn((B, D, H, W, cin), generator=g).to(BF).to(dev()) if accumulate else None
    plain = base.clone() if accumulate else torch.empty((B, D, H, W, cin), dtype=BF, device=dev())
    ops.conv_dgrad(geom, plain, dy, wd, accumulate=accumulate)
    assert ops.conv_dgrad_fuses_norm_bwd(geom, plain, dy, wd, nb) == fused
    out = base.clone() if accumulate else torch.empty_like(plain)
    bstats = torch.zeros((B, cin, 2), dtype=torch.float64, device=dev())
    launches = lib.launch_count()
    ops.conv_dgrad(geom, out, dy, wd, accumulate=accumulate, norm_bwd=nb, bstats=bstats)
    assert lib.launch_count() - launches == (1 if fused else 2)
    assert torch.equal(out, plain)
    want = torch.zeros_like(bstats)
    lib.inorm_lrelu_bwd_stats(plain.data_ptr(), cin, y_prev.data_ptr(), cin, stats.data_ptr(), gamma.data_ptr(),
                              beta.data_ptr(), B, V, cin, 1e-5, 0.01, want.data_ptr(), st)
    if not fused:
        assert torch.equal(bstats, want)
        return
    # fp32 partial sums in a different order, and sum g'y - mean sum g' instead of sum g'(y - mean): judge both kernels
    # against an fp64 evaluation, in units of the root-sum-of-squares of the summands (the sums themselves cancel heavily)
    m_ = stats[..., 0] / V
    var = (stats[..., 1] / V - m_ * m_).clamp_min(0)
    rstd = (1.0 / torch.sqrt(var + 1e-5)).float()
    mean = m_.float()
    sc = gamma[None] * rstd
    shf = beta[None] - mean * gamma[None] * rstd
    yf = y_prev.float().reshape(B, V, cin)
    t = yf * sc[:, None] + shf[:, None]
    gp = plain.float().reshape(B, V, cin).double() * torch.where(t > 0, 1.0, 0.01).double()
    xh = (yf - mean[:, None]).double() * rstd[:, None].double()
    ref = torch.stack([gp.sum(1), (gp * xh).sum(1)], -1)
    rss = torch.stack([gp.pow(2).sum(1).sqrt(), (gp * xh).pow(2).sum(1).sqrt()], -1) + 1e-12
    err_fused = float(((bstats - ref).abs() / rss).max())
    err_pass = float(((want - ref).abs() / rss).max())
    print('norm-backward sums vs fp64: epilogue %.2e, streaming pass %.2e (units of rss)' % (err_fused, err_pass))
    assert err_fused < 2e-6


def test_dgrad_norm_backward_statistics_fallback_pass(m):
    """shapes the halo epilogue does not cover (stride 2 here) still honour norm_bwd through a pass over the gradient"""
    lib, ops = m.lib, m.ops
    g = torch.Generator().manual_seed(32)
    B, E, cin, cout = 2, 12, 32, 64
    geom = ops.ConvGeom((3,) * 3, (2,) * 3, (1,) * 3)
    dy = torch.randn((B, E // 2, E // 2, E // 2, cout), generator=g).to(BF).to(dev())
    w = (torch.randn((cout, cin, 3, 3, 3), generator=g) / 30).to(dev())
    _, wd = ops.pack_weights(w)
    y_prev = torch.randn((B, E, E, E, cin), generator=g).to(BF).to(dev())
    gamma, beta = torch.ones(cin, device=dev()), torch.zeros(cin, device=dev())
    st = torch.cuda.current_stream().cuda_stream
    stats = torch.zeros((B, cin, 2), dtype=torch.float64, device=dev())
    lib.inorm_stats(y_prev.data_ptr(), cin, B, E ** 3, cin, stats.data_ptr(), st)
    nb = (y_prev, stats, gamma, beta, 1e-5, 0.01)
    dx = torch.empty_like(y_prev)
    assert not ops.conv_dgrad_fuses_norm_bwd(geom, dx, dy, wd, nb)
    bstats = torch.zeros((B, cin, 2), dtype=torch.float64, device=dev())
    ops.conv_dgrad(geom, dx, dy, wd, norm_bwd=nb, bstats=bstats)
    want = torch.zeros_like(bstats)
    lib.inorm_lrelu_bwd_stats(dx.data_ptr(), cin, y_prev.data_ptr(), cin, stats.data_ptr(), gamma.data_ptr(),
                              beta.data_ptr(), B, E ** 3, cin, 1e-5, 0.01, want.data_ptr(), st)
    assert torch.equal(bstats, want)


@pytest.mark.parametrize('C,K', [(32, 4), (320, 4), (64, 3), (128, 4), (256, 4), (64, 4)])
@pytest.mark.parametrize('vol', [(2, 5, 6, 7), (2, 48, 64, 40)])   # tails only / several steps of the staged sweep
def test_head(m, C, K, vol):
    ops = m.ops
    if C == 320 and vol[1] > 8:
        pytest.skip('the 320-channel head only exists at 8^3')
    z = rand_cl(vol + (C,), 31)
    w = (torch.randn(K, C, 1, 1, 1) / np.sqrt(C)).to(dev()).requires_grad_(True)
    b = (torch.randn(K) * 0.1).to(dev()).requires_grad_(True)
    zin = z.clone().requires_grad_(True)
    out = ops.HeadFn.apply(zin, w, b, None)
    zr = z.float().permute(0, 4, 1, 2, 3).requires_grad_(True)
    wr = w.detach().to(BF).float().requires_grad_(True)
    br = b.detach().to(BF).float().requires_grad_(True)
    outr = F.conv3d(zr, wr, br)
    assert rel_err(out.float().permute(0, 4, 1, 2, 3), outr) < 5e-3
    g = rand_cl(tuple(out.shape), 32)
    out.backward(g)
    outr.backward(g.float().permute(0, 4, 1, 2, 3))
    assert rel_err(zin.grad.float().permute(0, 4, 1, 2, 3), zr.grad) < 5e-3
    assert rel_err(w.grad, wr.grad) < 1e-3
    assert rel_err(b.grad, br.grad) < 1e-3


# ------------------------------------------------------------------------------------------------------------------
def _logits_targets(B, C, shape, seed, scales=1):
    g = torch.Generator().manual_seed(seed)
    outs, tgts = [], []
    cur = list(shape)
    for _ in range(scales):
        lg = (torch.randn((B, *cur, C), generator=g) * 2).to(BF).to(dev())
        tg = torch.randint(0, C, (B, 1, *cur), generator=g).float().to(dev())
        outs.append(lg)
        tgts.append(tg)
        cur = [max(1, c // 2) for c in cur]
    return outs, tgts


@pytest.mark.parametrize('batch_dice', [False, True])
@pytest.mark.parametrize('C', [4, 3])
@pytest.mark.parametrize('shape', [(9, 10, 11), (8, 10, 12), (40, 36, 44)])   # V % 4 != 0 (per-voxel kernels) / == 0 (quad-staged)
def test_dc_and_ce_matches_oracle(m, batch_dice, C, shape):
    import oracle
    (lg,), (tg,) = _logits_targets(2, C, shape, 41)
    mine = m.DC_and_CE_loss({'batch_dice': batch_dice, 'smooth': 1e-5, 'do_bg': False, 'ddp': False}, {},
                            weight_ce=1, weight_dice=1, ignore_label=None, dice_class=m.MemoryEfficientSoftDiceLoss)
    ref = oracle.DC_and_CE_loss({'batch_dice': batch_dice, 'smooth': 1e-5, 'do_bg': False, 'ddp': False}, {},
                                weight_ce=1, weight_dice=1, ignore_label=None,
                                dice_class=oracle.MemoryEfficientSoftDiceLoss)
    x = m.ops.ncdhw_view(lg).requires_grad_(True)
    l = mine(x, tg)
    xr = lg.float().permute(0, 4, 1, 2, 3).requires_grad_(True)
    lr = ref(xr, tg)
    assert abs(float(l) - float(lr)) <= 1e-5 * max(1.0, abs(float(lr)))
    (l * 3.0).backward()
    (lr * 3.0).backward()
    assert rel_err(x.grad.float(), xr.grad) < 6e-3     # bf16 rounding of the gradient
    # individual members
    dc = m.MemoryEfficientSoftDiceLoss(apply_nonlin=m.softmax_helper_dim1, batch_dice=batch_dice, do_bg=False,
                                       smooth=1e-5, ddp=False)(m.ops.ncdhw_view(lg), tg)
    dcr = oracle.MemoryEfficientSoftDiceLoss(apply_nonlin=oracle.softmax_helper_dim1, batch_dice=batch_dice,
                                             do_bg=False, smooth=1e-5, ddp=False)(lg.float().permute(0, 4, 1, 2, 3), tg)
    assert abs(float(dc) - float(dcr)) <= 1e-5
    ce = m.RobustCrossEntropyLoss()(m.ops.ncdhw_view(lg), tg)
    cer = oracle.RobustCrossEntropyLoss()(lg.float().permute(0, 4, 1, 2, 3), tg)
    assert abs(float(ce) - float(cer)) <= 1e-5 * float(cer)


def test_deep_supervision_wrapper_matches_oracle(m):
    import oracle
    outs, tgts = _logits_targets(2, 4, (16, 16, 12), 43, scales=4)
    w = m.deep_supervision_weights(4)
    mk = lambda mod: mod.DeepSupervisionWrapper(
        mod.DC_and_CE_loss({'batch_dice': False, 'smooth': 1e-5, 'do_bg': False, 'ddp': False}, {}, weight_ce=1,
                           weight_dice=1, ignore_label=None, dice_class=mod.MemoryEfficientSoftDiceLoss), w)
    xs = [m.ops.ncdhw_view(o).requires_grad_(True) for o in outs]
    l = mk(m)(xs, tgts)
    xr = [o.float().permute(0, 4, 1, 2, 3).requires_grad_(True) for o in outs]
    lr = mk(oracle)(xr, tgts)
    assert abs(float(l) - float(lr)) <= 1e-5 * max(1.0, abs(float(lr)))
    l.backward()
    lr.backward()
    for i in range(3):
        assert rel_err(xs[i].grad.float(), xr[i].grad) < 6e-3
    assert xs[3].grad is None and float(xr[3].grad.abs().max()) == 0.0   # zero-weighted scale


@pytest.mark.parametrize('batch_dice', [False, True])
def test_deep_supervision_two_networks_one_launch(m, batch_dice):
    """loss(out1, tgt) + loss(out2, tgt) of the mutual-distillation step through DeepSupervisionWrapper.forward_networks:
    one fused forward launch (all scales of both networks, last-block-done scalar algebra) and one backward launch;
    value 1e-5 and gradients against the oracle, and identical to the two separate calls."""
    import oracle
    outs1, tgts = _logits_targets(2, 4, (16, 20, 12), 43, scales=4)
    outs2, _ = _logits_targets(2, 4, (16, 20, 12), 47, scales=4)
    w = m.deep_supervision_weights(4)
    mk = lambda mod: mod.DeepSupervisionWrapper(
        mod.DC_and_CE_loss({'batch_dice': batch_dice, 'smooth': 1e-5, 'do_bg': False, 'ddp': False}, {}, weight_ce=1,
                           weight_dice=1, ignore_label=None, dice_class=mod.MemoryEfficientSoftDiceLoss), w)
    x1 = [m.ops.ncdhw_view(o).requires_grad_(True) for o in outs1]
    x2 = [m.ops.ncdhw_view(o).requires_grad_(True) for o in outs2]
    before = m.lib.launch_count()
    l = mk(m).forward_networks([x1, x2], tgts)
    fwd_launches = m.lib.launch_count() - before
    r1 = [o.float().permute(0, 4, 1, 2, 3).requires_grad_(True) for o in outs1]
    r2 = [o.float().permute(0, 4, 1, 2, 3).requires_grad_(True) for o in outs2]
    lr = mk(oracle)(r1, tgts) + mk(oracle)(r2, tgts)
    assert abs(float(l) - float(lr)) <= 1e-5 * max(1.0, abs(float(lr)))
    before = m.lib.launch_count()
    (l * 2.0).backward()
    bwd_launches = m.lib.launch_count() - before
    (lr * 2.0).backward()
    assert fwd_launches == 1 and bwd_launches == 1, (fwd_launches, bwd_launches)
    for xs, rs in ((x1, r1), (x2, r2)):
        for i in range(3):
            assert rel_err(xs[i].grad.float(), rs[i].grad) < 6e-3
        assert xs[3].grad is None
    # same numbers as two separate wrapper calls
    y1 = [m.ops.ncdhw_view(o).requires_grad_(True) for o in outs1]
    l1 = mk(m)(y1, tgts)
    l2 = mk(m)([m.ops.ncdhw_view(o) for o in outs2], tgts)
    assert abs(float(l1) + float(l2) - float(l)) <= 1e-6 * max(1.0, abs(float(l)))
    (l1 * 2.0).backward()
    assert rel_err(y1[0].grad.float(), x1[0].grad.float()) < 1e-3
    # repeated calls: the ticket counter is left at zero
    l_again = mk(m).forward_networks([x1, x2], tgts)
    assert abs(float(l_again) - float(l)) <= 1e-6 * max(1.0, abs(float(l)))


@pytest.mark.parametrize('batch_dice', [False, True])
def test_dc_and_ce_ignore_label_matches_oracle(m, batch_dice):
    """DC_and_CE_loss(ignore_label = the id behind the last class, nnUNetTrainer.py:353-361): ignored voxels are masked out
    of the Dice sums (loss_mask) and of the cross-entropy mean (ignore_index) and receive no gradient -- value and
    gradients against the oracle, through the deep-supervision wrapper; plus the all-ignored corner (CE term = 0)."""
    import oracle
    outs, tgts = _logits_targets(2, 4, (16, 20, 12), 51, scales=3)
    g = torch.Generator().manual_seed(52)
    tgts = [t.clone() for t in tgts]
    for t in tgts:                                   # ~20 % ignored voxels, in blobs and scattered
        t[:, :, : t.shape[2] // 3] = 4.0
        mask = torch.rand(t.shape, generator=g).to(t.device) < 0.05
        t[mask] = 4.0
    w = m.deep_supervision_weights(3)
    mk = lambda mod: mod.DeepSupervisionWrapper(
        mod.DC_and_CE_loss({'batch_dice': batch_dice, 'smooth': 1e-5, 'do_bg': False, 'ddp': False}, {}, weight_ce=1,
                           weight_dice=1, ignore_label=4, dice_class=mod.MemoryEfficientSoftDiceLoss), w)
    x = [m.ops.ncdhw_view(o).requires_grad_(True) for o in outs]
    r = [o.float().permute(0, 4, 1, 2, 3).requires_grad_(True) for o in outs]
    l, lr = mk(m)(x, tgts), mk(oracle)(r, tgts)
    assert abs(float(l) - float(lr)) <= 1e-5 * max(1.0, abs(float(lr))), (float(l), float(lr))
    l.backward()
    lr.backward()
    for i in range(2):
        assert rel_err(x[i].grad.float(), r[i].grad) < 6e-3
        ign = (tgts[i] == 4).expand(-1, 4, -1, -1, -1)
        assert float(x[i].grad.float()[ign].abs().max()) == 0.0          # no gradient on ignored voxels
    # without ignored voxels nothing changes with respect to ignore_label=None
    outs2, tg2 = _logits_targets(2, 4, (16, 20, 12), 53, scales=3)
    plain = m.DeepSupervisionWrapper(
        m.DC_and_CE_loss({'batch_dice': batch_dice, 'smooth': 1e-5, 'do_bg': False, 'ddp': False}, {}, weight_ce=1,
                         weight_dice=1, ignore_label=None, dice_class=m.MemoryEfficientSoftDiceLoss), w)
    xs = [m.ops.ncdhw_view(o) for o in outs2]
    assert float(plain(xs, tg2)) == float(mk(m)(xs, tg2))
    # every voxel ignored: the cross-entropy term is 0 (the reference's num_fg > 0 guard), Dice = -smooth/smooth
    allign = [torch.full_like(t, 4.0) for t in tg2]
    la, lo = mk(m)(xs, allign), mk(oracle)([o.float().permute(0, 4, 1, 2, 3) for o in outs2], allign)
    assert abs(float(la) - float(lo)) <= 1e-5 * max(1.0, abs(float(lo))), (float(la), float(lo))
    # an ignore label inside the class range is refused
    bad = m.DC_and_CE_loss({'batch_dice': False, 'smooth': 1e-5, 'do_bg': False, 'ddp': False}, {}, ignore_label=2)
    with pytest.raises(NotImplementedError):
        bad(xs[0], tg2[0])


@pytest.mark.parametrize('T', [1.0, 2.0])
def test_distill_kl_single_pass_with_upstream_factor(m, T):
    """distill_kl(..., upstream_grad=lambda1): loss and both gradients from ONE kernel; exact for the announced upstream
    factor (backward = a rescale launch that exits on the device) and for any other one (rescaled)."""
    import oracle
    (a,), _ = _logits_targets(2, 4, (36, 40, 44), 51)
    (b,), _ = _logits_targets(2, 4, (36, 40, 44), 52)
    ar, br = (t.float().permute(0, 4, 1, 2, 3).detach().clone().requires_grad_(True) for t in (a, b))
    lr = oracle.distill_kl(ar, br, T)
    (0.5 * lr).backward()
    for actual in (0.5, 1.25):
        av, bv = (m.ops.ncdhw_view(t).detach().requires_grad_(True) for t in (a, b))
        before = m.lib.launch_count()
        l = m.distill_kl(av, bv, T, upstream_grad=0.5)
        assert m.lib.launch_count() - before == 2          # the fused pass + the scalar scaling
        assert abs(float(l) - float(lr)) <= 1e-5 * max(abs(float(lr)), 1e-3)
        (actual * l).backward()
        assert rel_err(av.grad.float(), ar.grad * (actual / 0.5)) < 8e-3
        assert rel_err(bv.grad.float(), br.grad * (actual / 0.5)) < 8e-3
    # only one side needs a gradient (a detached teacher)
    av = m.ops.ncdhw_view(a).detach().requires_grad_(True)
    l = m.distill_kl(av, m.ops.ncdhw_view(b), T, upstream_grad=0.5)
    (0.5 * l).backward()
    assert rel_err(av.grad.float(), ar.grad) < 8e-3


def test_argmax_tp_fp_fn(m):
    import oracle
    (lg,), (tg,) = _logits_targets(2, 4, (7, 8, 9), 44)
    tp, fp, fn = m.ops.argmax_tp_fp_fn(m.ops.ncdhw_view(lg), tg)
    x = lg.float().permute(0, 4, 1, 2, 3)
    onehot = torch.zeros_like(x).scatter_(1, x.argmax(1)[:, None], 1)
    tpr, fpr, fnr, _ = oracle.get_tp_fp_fn_tn(onehot, tg, axes=[0, 2, 3, 4])
    assert torch.equal(tp.float(), tpr) and torch.equal(fp.float(), fpr) and torch.equal(fn.float(), fnr)


@pytest.mark.parametrize('C,T', [(4, 1.0), (4, 2.0), (1, 1.0)])
@pytest.mark.parametrize('shape', [(8, 9, 10), (7, 9, 5), (36, 40, 44)])   # NV % 4 == 0 (quad-staged) / != 0 (per-voxel)
def test_distill_kl_matches_oracle(m, C, T, shape):
    import oracle
    (a,), _ = _logits_targets(2, 4, shape, 51)
    (b,), _ = _logits_targets(2, 4, shape, 52)
    av, bv = m.ops.ncdhw_view(a), m.ops.ncdhw_view(b)
    ar, br = a.float().permute(0, 4, 1, 2, 3), b.float().permute(0, 4, 1, 2, 3)
    if C == 1:   # the vessel-channel call site (MVDTrainer.py:897-899): a strided one-channel view
        av, bv, ar, br = av[:, 2:3], bv[:, 2:3], ar[:, 2:3], br[:, 2:3]
    av, bv = av.detach().requires_grad_(True), bv.detach().requires_grad_(True)
    ar, br = ar.detach().clone().requires_grad_(True), br.detach().clone().requires_grad_(True)
    l = m.distill_kl(av, bv, T)
    lr = oracle.distill_kl(ar, br, T)
    assert abs(float(l) - float(lr)) <= 1e-5 * max(abs(float(lr)), 1e-3)
    l.backward()
    lr.backward()
    assert rel_err(av.grad.float(), ar.grad) < 6e-3
    assert rel_err(bv.grad.float(), br.grad) < 6e-3


def test_distill_kl_golden(m, golden_misc):
    g = golden_misc
    for tag in ('c4_T1', 'c4_T2', 'c1_T1'):
        ys = torch.from_numpy(g[f'kl.{tag}.ys']).to(BF)
        yt = torch.from_numpy(g[f'kl.{tag}.yt']).to(BF)
        # the fixture inputs are fp32; compare on their bf16 roundings via the oracle (pinned to the same fixture)
        import oracle
        want = oracle.distill_kl(ys.float(), yt.float(), float(g[f'kl.{tag}.T']))
        got = m.distill_kl(ys.to(dev()), yt.to(dev()), float(g[f'kl.{tag}.T']))
        assert abs(float(got) - float(want)) <= 1e-5 * max(abs(float(want)), 1e-3)


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('vol', ['smooth', 'ties'])
@pytest.mark.parametrize('fn', ['soft_erode', 'soft_dilate', 'soft_open'])
def test_morphology_golden_bit_exact(m, golden_skel, vol, fn):
    g = golden_skel
    x = torch.from_numpy(g[f'{vol}.{fn}.in']).to(dev()).requires_grad_(True)
    y = getattr(m, fn)(x)
    assert np.array_equal(y.detach().cpu().numpy(), g[f'{vol}.{fn}.out'])
    (y * torch.from_numpy(g[f'{vol}.{fn}.w']).to(dev())).sum().backward()
    np.testing.assert_allclose(x.grad.cpu().numpy(), g[f'{vol}.{fn}.grad'], rtol=1e-5, atol=1e-6)  # fp32 atomics: order-dependent last bits


@pytest.mark.parametrize('vol', ['smooth', 'ties'])
@pytest.mark.parametrize('it', [0, 1, 3])
def test_soft_skel_golden(m, golden_skel, vol, it):
    g = golden_skel
    x = torch.from_numpy(g[f'{vol}.soft_erode.in']).to(dev()).requires_grad_(True)
    y = m.soft_skel(x, it)
    assert np.array_equal(y.detach().cpu().numpy(), g[f'{vol}.soft_skel{it}.out'])
    (y * torch.from_numpy(g[f'{vol}.soft_skel{it}.w']).to(dev())).sum().backward()
    np.testing.assert_allclose(x.grad.cpu().numpy(), g[f'{vol}.soft_skel{it}.grad'], rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize('shape,it', [((2, 37, 21, 45), 3), ((1, 9, 40, 70), 10), ((1, 3, 5, 4), 6)])
def test_soft_skel_fused_equals_level_by_level(m, shape, it):
    """the on-chip multi-level kernel (<= 4 levels per launch, tiles with halos, volume borders) against the
    level-by-level kernels (mvd_soft_erode + mvd_skel_update): bit-identical skeleton, E and skeleton stacks; and the
    fused shared-memory backward (<= 2 levels per launch, delta recomputed) against the level-by-level backward kernels
    with global atomics (chain + dilate / erosion routing per level)."""
    g = torch.Generator().manual_seed(17)
    x = torch.rand(shape, generator=g).to(dev())
    x[x < 0.3] = 0.0            # plateaus -> ties
    B, D, H, W = shape
    N, L, st = x.numel(), it + 1, torch.cuda.current_stream().cuda_stream
    x5 = x.unsqueeze(1)
    sk, E, skel = m.ops._skel_forward(x5, it, True)
    sk2, _, _ = m.ops._skel_forward(x5, it, False)
    Er = torch.empty((L + 1, N), device=dev()); Er[0] = x.reshape(-1)
    dr = torch.empty((L, N), device=dev()); sr = torch.empty((L, N), device=dev())
    for j in range(L):
        m.lib.soft_erode(Er[j].data_ptr(), Er[j + 1].data_ptr(), B, D, H, W, st)
        m.lib.skel_update(Er[j].data_ptr(), Er[j + 1].data_ptr(), sr[j - 1].data_ptr() if j else None,
                          dr[j].data_ptr(), sr[j].data_ptr(), 1 if j == 0 else 0, B, D, H, W, st)
    assert torch.equal(E, Er[1:]) and torch.equal(skel, sr)
    assert torch.equal(sk.reshape(-1), sr[L - 1]) and torch.equal(sk2.reshape(-1), sr[L - 1])
    # backward: fused vs level by level
    g_skel = torch.randn((N,), generator=g).to(dev())
    got = m.ops._skel_backward(g_skel, x.reshape(-1), E, skel, (B, D, H, W))
    g_delta = torch.empty((L, N), device=dev())
    m.lib.skel_chain_bwd(dr.data_ptr(), sr.data_ptr(), g_skel.data_ptr(), g_delta.data_ptr(), L, N, st)
    gE = torch.zeros((L + 1, N), device=dev())
    for j in range(L):
        m.lib.skel_level_bwd(Er[j + 1].data_ptr(), dr[j].data_ptr(), g_delta[j].data_ptr(), gE[j].data_ptr(),
                             gE[j + 1].data_ptr(), B, D, H, W, st)
    for lvl in range(L, 0, -1):
        m.lib.soft_erode_bwd(Er[lvl - 1].data_ptr(), gE[lvl].data_ptr(), gE[lvl - 1].data_ptr(), B, D, H, W, st)
    torch.testing.assert_close(got, gE[0], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize('shape,it', [((2, 37, 21, 45), 3), ((1, 20, 40, 35), 10), ((1, 3, 5, 4), 2), ((1, 12, 18, 33), 0)])
def test_soft_skel_backward_matches_reference_autograd(m, shape, it):
    """gradient of soft_skel through the fused backward against PyTorch autograd of the reference formulation
    (soft_skeleton.py:6-37 as restated in oracle/losses.py; tie rules of max_pool3d / torch.min / relu): several tiles,
    volume borders, plateaus, odd and even level counts (1- and 2-level launches)."""
    import oracle
    g = torch.Generator().manual_seed(23)
    x = torch.rand((shape[0], 1, *shape[1:]), generator=g)
    x[x < 0.25] = 0.0
    x[x > 0.9] = 1.0
    w = torch.randn(x.shape, generator=g).to(dev())
    xa = x.to(dev()).requires_grad_(True)
    (m.soft_skel(xa, it) * w).sum().backward()
    xr = x.to(dev()).requires_grad_(True)
    (oracle.soft_skel(xr, it) * w).sum().backward()
    torch.testing.assert_close(xa.grad, xr.grad, rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize('it', [3, 10])
def test_soft_cldice_matches_oracle(m, it):
    import oracle
    lab = oracle.structured_labels(2, (24, 28, 20), seed=5)
    gt = (lab == 2).float().to(dev())
    g = torch.Generator().manual_seed(9)
    logits = (torch.randn((2, 24, 28, 20, 4), generator=g) * 1.5).to(BF).to(dev())
    logits[..., 2] += (gt[:, 0] * 3).to(BF)
    x = m.ops.ncdhw_view(logits).detach().requires_grad_(True)
    prob = m.softmax_channel(x, 2)
    l = m.soft_cldice(iter_=it, smooth=1.)(gt, prob)
    xr = logits.float().permute(0, 4, 1, 2, 3).detach().clone().requires_grad_(True)
    pr = torch.softmax(xr, 1)[:, 2:3]
    lr = oracle.soft_cldice(iter_=it, smooth=1.)(gt, pr)
    assert torch.equal(prob.detach(), pr.detach()) or rel_err(prob.detach(), pr.detach()) < 1e-6
    assert abs(float(l) - float(lr)) <= 1e-5
    l.backward()
    lr.backward()
    assert rel_err(x.grad.float(), xr.grad) < 1e-2


# ------------------------------------------------------------------------------------------------------------------
def test_sgd_nesterov_clip_matches_torch(m):
    torch.manual_seed(3)
    shapes = [(32, 2, 3, 3, 3), (32,), (64, 32, 3, 3, 3), (5000,), (4, 32, 1, 1, 1)]
    p_ref = [torch.nn.Parameter(torch.randn(s, device=dev())) for s in shapes]
    p_mine = [torch.nn.Parameter(p.detach().clone()) for p in p_ref]
    ref = torch.optim.SGD(p_ref, 1e-2, weight_decay=3e-5, momentum=0.99, nesterov=True)
    mine = m.SGDNesterovClip(p_mine, 1e-2, weight_decay=3e-5, momentum=0.99, nesterov=True, max_norm=12.0)
    for it in range(4):
        scale = 30.0 if it % 2 == 0 else 0.01     # clipped and unclipped steps
        grads = [torch.randn_like(p) * scale for p in p_ref]
        for a, b, g in zip(p_ref, p_mine, grads):
            a.grad, b.grad = g.clone(), g.clone()
        if it == 2:
            p_ref[3].grad = torch.zeros_like(p_ref[3])
            p_mine[3].grad = None                  # missing gradient behaves like a zero gradient
        total = torch.nn.utils.clip_grad_norm_(p_ref, 12)
        ref.step()
        mine.step()
        assert float(mine.last_sqnorm.sqrt()) == pytest.approx(float(total), rel=1e-5)
        for a, b in zip(p_ref, p_mine):
            assert rel_err(b.detach(), a.detach()) < 1e-6
        ref.param_groups[0]['lr'] = mine.param_groups[0]['lr'] = 5e-3


# ------------------------------------------------------------------------------------------------------------------
# deep-supervision targets on the GPU (mvd_downsample_seg_nearest): bit-exact against the oracle transform
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize('shape,scales', [
    ((2, 1, 16, 16, 12), [[1, 1, 1], [0.5, 0.5, 0.5], [0.25, 0.25, 0.25], [0.125, 0.125, 0.25]]),
    ((1, 2, 20, 18, 9), [[0.5, 0.5, 0.5], [1, 1, 1], [0.25, 0.25, 1]]),
    ((2, 1, 40, 40, 24), [[1, 1, 1], [0.5, 0.5, 0.5], [0.25, 0.25, 0.25], [0.125, 0.125, 0.125], [0.0625, 0.0625, 0.125]]),
])
def test_ds_targets_match_oracle(m, shape, scales):
    from oracle import ds_targets
    rng = np.random.default_rng(9)
    seg = rng.integers(-1, 4, size=shape).astype(np.float32)
    want = ds_targets.DownsampleSegForDSTransform2(scales, 0, input_key='target', output_key='target')(target=seg)['target']
    x = torch.from_numpy(seg).to(dev())
    got = m.DownsampleSegForDSTransform2(scales, 0, input_key='target', output_key='target')(target=x)['target']
    assert len(got) == len(want)
    for g, w, s in zip(got, want, scales):
        if all(v == 1 for v in s):
            assert g is x
        assert tuple(g.shape) == w.shape and np.array_equal(g.cpu().numpy(), w)
    with pytest.raises(RuntimeError):
        m.downsample_seg_for_ds(torch.from_numpy(seg), scales)      # no CPU path
    with pytest.raises(NotImplementedError):
        m.DownsampleSegForDSTransform2(scales, order=1)


def test_train_step_builds_ds_targets_on_gpu(m):
    """a batch carrying only the full-resolution target gives the same loss as the list the oracle transform builds."""
    from oracle import ds_targets
    patch = (32, 32, 32)
    plans, dj = m.make_plans(patch, batch_size=2, n_modalities=2, n_classes=4)
    losses = []
    for mode in ('list', 'full'):
        torch.manual_seed(0)
        tr = m.nnUNetTrainer(plans, '3d_fullres', 0, dj, device=dev())
        tr.initialize()
        tr.on_train_epoch_start()
        g = torch.Generator().manual_seed(5)
        data = torch.rand((2, 2) + patch, generator=g)
        full = torch.round(torch.rand((2, 1) + patch, generator=g) * 3)
        scales = tr._get_deep_supervision_scales()
        if mode == 'list':
            tl = ds_targets.DownsampleSegForDSTransform2(scales, 0, input_key='t', output_key='t')(t=full.numpy())['t']
            target = [torch.from_numpy(np.ascontiguousarray(t)) for t in tl]
        else:
            target = full
        losses.append(float(tr.train_step({'data': data, 'target': target})['loss']))
    assert abs(losses[0] - losses[1]) <= 1e-5 * max(1.0, abs(losses[0]))    # same targets -> same loss (fp64-atomic order aside)
