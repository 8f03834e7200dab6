"""Sliding-window inference: golden vectors of the reference's pure helpers on CPU, and the
CUDA path (tcgen05 network forward + mvd_sw_accumulate) against the oracle restatement on the GPU."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, 'tests', 'golden', 'inference.npz')


@pytest.fixture(scope='module')
def gold():
    return np.load(GOLD)


def _n(gold, prefix):
    return len({k.split('.')[0] for k in gold.files if k.startswith(prefix)})


def test_oracle_gaussian_and_steps_match_reference_golden(gold):
    import oracle
    for i in range(_n(gold, 'gauss')):
        tile = tuple(int(v) for v in gold[f'gauss{i}.tile'])
        g = oracle.inference.compute_gaussian(tile, 1. / 8, 1000, torch.float32, 'cpu').numpy()
        np.testing.assert_array_equal(g, gold[f'gauss{i}.out'])
    for i in range(_n(gold, 'steps')):
        st = oracle.inference.compute_steps_for_sliding_window(tuple(gold[f'steps{i}.img']), tuple(gold[f'steps{i}.tile']),
                                                               float(gold[f'steps{i}.step']))
        for ax in range(3):
            assert st[ax] == list(gold[f'steps{i}.ax{ax}'])


def test_host_gaussian_and_steps_match_reference_golden(gold):
    """the product's own host helpers (no device work involved)."""
    import multimodal_mvd_seg_b200 as m
    for i in range(_n(gold, 'gauss')):
        tile = tuple(int(v) for v in gold[f'gauss{i}.tile'])
        g = m.compute_gaussian(tile, 1. / 8, 1000, torch.float32, torch.device('cpu')).numpy()
        np.testing.assert_array_equal(g, gold[f'gauss{i}.out'])
    for i in range(_n(gold, 'steps')):
        st = m.compute_steps_for_sliding_window(tuple(int(v) for v in gold[f'steps{i}.img']),
                                                tuple(int(v) for v in gold[f'steps{i}.tile']), float(gold[f'steps{i}.step']))
        for ax in range(3):
            assert st[ax] == list(gold[f'steps{i}.ax{ax}'])


def test_predictor_refuses_cpu():
    import multimodal_mvd_seg_b200 as m
    with pytest.raises(m.MvdError):
        m.SlidingWindowPredictor(None, (32, 32, 32), 4, device=torch.device('cpu'))


@pytest.mark.gpu
def test_sw_accumulate_kernel_exact():
    import multimodal_mvd_seg_b200 as m
    dev = torch.device('cuda:0')
    g = torch.Generator().manual_seed(3)
    K, (d, h, w), (D, H, W) = 4, (5, 6, 7), (9, 8, 11)
    pred = torch.randn((d, h, w, K), generator=g).to(torch.bfloat16).to(dev)
    gw = torch.rand((d, h, w), generator=g).to(dev) + 0.1
    acc = torch.randn((K, D, H, W), generator=g).to(dev)
    npred = torch.rand((D, H, W), generator=g).to(dev)
    want_acc, want_n = acc.clone(), npred.clone()
    z0, y0, x0 = 3, 1, 4
    want_acc[:, z0:z0 + d, y0:y0 + h, x0:x0 + w] += pred.float().permute(3, 0, 1, 2) * 0.5 * gw
    want_n[z0:z0 + d, y0:y0 + h, x0:x0 + w] += gw
    st = torch.cuda.current_stream().cuda_stream
    m.lib.sw_accumulate(pred.data_ptr(), K, gw.data_ptr(), 0.5, acc.data_ptr(), npred.data_ptr(), K, d, h, w, D, H, W,
                        z0, y0, x0, 0, st)
    torch.testing.assert_close(acc, want_acc, rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(npred, want_n, rtol=1e-6, atol=1e-6)
    # mirrored read-back (flip_mask bit 0: z, 1: y, 2: x), no npred update
    for mask in (1, 2, 4, 3, 5, 6, 7):
        dims = [i for i in range(3) if mask >> i & 1]
        want_acc[:, z0:z0 + d, y0:y0 + h, x0:x0 + w] += pred.float().flip(dims).permute(3, 0, 1, 2) * 0.25 * gw
        m.lib.sw_accumulate(pred.data_ptr(), K, gw.data_ptr(), 0.25, acc.data_ptr(), None, K, d, h, w, D, H, W,
                            z0, y0, x0, mask, st)
        torch.testing.assert_close(acc, want_acc, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(npred, want_n, rtol=1e-6, atol=1e-6)
    m.lib.sw_finalize(acc.data_ptr(), npred.data_ptr(), K, D * H * W, st)
    torch.testing.assert_close(acc, want_acc / want_n, rtol=1e-5, atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize('mirror', [None, (0, 1, 2)])
def test_sliding_window_matches_oracle(mirror):
    import multimodal_mvd_seg_b200 as m
    import oracle
    dev = torch.device('cuda:0')
    patch = (32, 32, 32)
    ref = oracle.build_plain_conv_unet(2, 4, patch, seed=0).to(dev).eval()
    plans, dj = m.make_plans(patch, batch_size=1, n_modalities=2, n_classes=4)
    tr = m.nnUNetTrainer(plans, '3d_fullres', 0, dj, device=dev)
    tr.initialize()
    tr.network.load_state_dict(ref.state_dict())
    g = torch.Generator().manual_seed(11)
    img = torch.randn((2, 40, 30, 37), generator=g).to(dev)      # one axis smaller than the patch -> padded
    pred = m.SlidingWindowPredictor(tr.network, patch, 4, tile_step_size=0.5, use_gaussian=True,
                                    use_mirroring=mirror is not None, allowed_mirroring_axes=mirror, device=dev)
    got = pred.predict_sliding_window_return_logits(img)
    ref.decoder.deep_supervision = False
    want = oracle.inference.predict_sliding_window_return_logits(ref, img, patch, 4, 0.5, True, mirror, autocast_bf16=True)
    assert got.shape == want.shape == (4, 40, 30, 37)
    rel = float((got - want).norm() / want.norm())
    assert rel < 2e-2, rel
    # segmentation masks: identical wherever the reference's own top-2 margin exceeds the bf16 tolerance
    top2 = want.topk(2, dim=0).values
    decisive = (top2[0] - top2[1]) > 2e-2 * want.abs().max()
    agree = (got.argmax(0) == want.argmax(0))[decisive].float().mean()
    assert float(agree) >= 0.999, float(agree)
    assert tr.network.decoder.deep_supervision and tr.network.training == tr.network.training
