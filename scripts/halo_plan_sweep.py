"""Sweep of the tap-major halo kernel's plan (N-tile width, output planes per item, weight-ring depth) per layer shape,
through the debug hook mvd_debug_set_halo_plan.  The GPU is kept busy ahead of every timed launch (torch.cuda._sleep)
so host launch overhead does not leak into the event timing of 10-50 us kernels.
usage: halo_plan_sweep.py [B=2]"""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_mvd_seg_b200 as m
from multimodal_mvd_seg_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device('cuda:0')
cd = ctypes.CDLL(m.LIB_PATH)
cd.mvd_debug_set_halo_plan.argtypes = [ctypes.c_int] * 3
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
LAYERS = [(128, 128, 32), (256, 128, 32), (128, 64, 64), (256, 256, 16), (512, 256, 16), (320, 320, 8), (640, 320, 8),
          (320, 320, 4)]
geom = ops.ConvGeom((3,) * 3, (1,) * 3, (1,) * 3)


def bench(fn, iters=7):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        torch.cuda._sleep(400000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


for cin, cout, E in LAYERS:
    x = torch.randn((B, E, E, E, cin), device=dev).to(torch.bfloat16)
    y = torch.randn((B, E, E, E, cout), device=dev).to(torch.bfloat16)
    w = torch.randn((cout, cin, 3, 3, 3), device=dev) * 0.05
    wf, wd = ops.pack_weights(w)
    for p in ('fprop', 'dgrad'):
        N = cout if p == 'fprop' else cin
        if N <= 64:
            continue
        fn = (lambda: ops.conv_fprop(geom, x, y, wf)) if p == 'fprop' else (lambda: ops.conv_dgrad(geom, x, y, wd))
        cd.mvd_debug_set_halo_plan(0, 0, 8)
        t_old = bench(fn)
        cd.mvd_debug_set_halo_plan(0, 0, 0)
        t_auto = bench(fn)
        row = f'{p} {cin:3d}->{cout:3d} {E:3d}^3: auto(wst<=8) {t_old * 1e3:6.1f}us auto {t_auto * 1e3:6.1f}us |'
        best = (t_auto, 'auto')
        for nt in range(32, 257, 32):
            if N % nt:
                continue
            for mt in (1, 2, 4):
                if 2 * mt * nt > 512 or mt > E:
                    continue
                cd.mvd_debug_set_halo_plan(nt, mt, 0)
                try:
                    t = bench(fn)
                except Exception:
                    continue
                row += f' {nt}x{mt} {t * 1e3:5.1f}'
                if t < best[0]:
                    best = (t, f'{nt}x{mt}')
        cd.mvd_debug_set_halo_plan(0, 0, 0)
        print(row + f' | best {best[1]} {best[0] * 1e3:.1f}us', flush=True)
