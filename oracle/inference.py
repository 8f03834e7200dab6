"""Oracle (test infrastructure only): pure-PyTorch restatement of the reference's sliding-window prediction,
inference/predict_from_raw_data.py:528-714 and inference/sliding_window_prediction.py:10-58, with fp32 accumulators.
Parity of compute_gaussian / compute_steps_for_sliding_window is PINNED by tests/golden/inference.npz (the reference's
function texts exec'd by tests/golden/make_golden.py); the predictor loop is restated from the cited lines (the class
itself imports batchgenerators / acvl_utils and cannot be imported here)."""
import numpy as np
import torch
from scipy.ndimage import gaussian_filter


def compute_gaussian(tile_size, sigma_scale=1. / 8, value_scaling_factor=1, dtype=torch.float32, device='cpu'):
    tmp = np.zeros(tile_size)
    tmp[tuple(i // 2 for i in tile_size)] = 1
    g = gaussian_filter(tmp, [i * sigma_scale for i in tile_size], 0, mode='constant', cval=0)
    g = torch.from_numpy(g).type(dtype).to(device)
    g = (g / torch.max(g) * value_scaling_factor).type(dtype)
    g[g == 0] = torch.min(g[g != 0])
    return g


def compute_steps_for_sliding_window(image_size, tile_size, tile_step_size):
    target = [i * tile_step_size for i in tile_size]
    num_steps = [int(np.ceil((i - k) / j)) + 1 for i, j, k in zip(image_size, target, tile_size)]
    steps = []
    for dim in range(len(tile_size)):
        max_step_value = image_size[dim] - tile_size[dim]
        actual = max_step_value / (num_steps[dim] - 1) if num_steps[dim] > 1 else 99999999999
        steps.append([int(np.round(actual * i)) for i in range(num_steps[dim])])
    return steps


@torch.no_grad()
def predict_sliding_window_return_logits(network, input_image, patch_size, num_heads, tile_step_size=0.5,
                                         use_gaussian=True, mirror_axes=(0, 1, 2), autocast_bf16=False):
    dev = input_image.device
    data = input_image.float()
    pads, revert = [], [slice(None)]
    for s, p in zip(data.shape[1:], patch_size):
        total = max(p - s, 0)
        lo = total // 2
        pads.append((lo, total - lo))
        revert.append(slice(lo, lo + s))
    if any(a or b for a, b in pads):
        data = torch.nn.functional.pad(data, [v for ab in reversed(pads) for v in ab])
    steps = compute_steps_for_sliding_window(data.shape[1:], patch_size, tile_step_size)
    slicers = [tuple([slice(None), *[slice(si, si + ti) for si, ti in zip((sx, sy, sz), patch_size)]])
               for sx in steps[0] for sy in steps[1] for sz in steps[2]]
    logits = torch.zeros((num_heads, *data.shape[1:]), dtype=torch.float32, device=dev)
    n_pred = torch.zeros(data.shape[1:], dtype=torch.float32, device=dev)
    g = compute_gaussian(tuple(patch_size), 1. / 8, 1000, torch.float32, dev) if use_gaussian else 1.0

    def net(x):
        if autocast_bf16:
            with torch.autocast(dev.type, dtype=torch.bfloat16):
                return network(x).float()
        return network(x).float()

    for sl in slicers:
        x = data[sl][None]
        pred = net(x)
        if mirror_axes is not None:
            combos = [c for c in ((2,), (3,), (4,), (2, 3), (2, 4), (3, 4), (2, 3, 4))
                      if all((a - 2) in mirror_axes for a in c)]
            for c in combos:
                pred = pred + torch.flip(net(torch.flip(x, c)), c)
            pred = pred / (2 ** len(mirror_axes))
        logits[sl] += pred[0] * g
        n_pred[sl[1:]] += g
    logits /= n_pred
    return logits[tuple(revert)]
