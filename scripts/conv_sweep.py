"""Per-layer sweep of the cfg-2 conv shapes through the C ABI: every pass with algo 'auto' (tcgen05 where covered) and
'generic' (CUDA-core tiles), CUDA-event timed with an L2 flush between iterations.
usage: conv_sweep.py [small|all]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_mvd_seg_b200 as m
from multimodal_mvd_seg_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else 'small'
dev = torch.device('cuda:0')
B = 2
# (Cin, Cout, input edge, k, s)
SMALL = [(320, 320, 8, 3, 1), (640, 320, 8, 3, 1), (256, 320, 16, 3, 2), (320, 320, 8, 3, 2), (320, 320, 4, 3, 1),
         (320, 320, 8, 2, 2), (256, 320, 16, 2, 2), (128, 256, 32, 2, 2), (64, 128, 64, 2, 2), (32, 64, 128, 2, 2),
         (256, 256, 16, 3, 1), (512, 256, 16, 3, 1), (128, 256, 32, 3, 2)]
BIG = [(32, 32, 128, 3, 1), (64, 32, 128, 3, 1), (32, 64, 128, 3, 2), (64, 64, 64, 3, 1), (128, 64, 64, 3, 1),
       (64, 128, 64, 3, 2), (128, 128, 32, 3, 1), (256, 128, 32, 3, 1)]
shapes = SMALL if which == 'small' else SMALL + BIG
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def bench(fn, iters=7):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


for cin, cout, E, k, s in shapes:
    geom = ops.ConvGeom((k,) * 3, (s,) * 3, ((k - 1) // 2,) * 3)
    Do, Ho, Wo = geom.out_size((E, E, E))
    x = torch.randn((B, E, E, E, cin), device=dev).to(torch.bfloat16)
    y = torch.randn((B, Do, Ho, Wo, cout), device=dev).to(torch.bfloat16)
    w = torch.randn((cout, cin, k, k, k), device=dev) * 0.05
    wf, wd = ops.pack_weights(w)
    dw = torch.empty_like(w)
    flops = 2.0 * B * Do * Ho * Wo * cout * cin * k ** 3
    row = f'{cin:4d}->{cout:4d} in{E:3d}^3 k{k} s{s}:'
    for p in ('fprop', 'dgrad', 'wgrad'):
        for algo in ('auto', 'generic'):
            ops.set_conv_algo(algo)
            fn = {'fprop': lambda: ops.conv_fprop(geom, x, y, wf), 'dgrad': lambda: ops.conv_dgrad(geom, x, y, wd),
                  'wgrad': lambda: ops.conv_wgrad(geom, x, y, dw)}[p]
            try:
                t = bench(fn)
                row += f'  {p[0]}{algo[0]} {t * 1e3:7.1f}us'
            except Exception as e:   # noqa
                row += f'  {p[0]}{algo[0]}   n/a  '
    ops.set_conv_algo('auto')
    print(row + f'   ({flops / 1e9:.2f} GF)', flush=True)
