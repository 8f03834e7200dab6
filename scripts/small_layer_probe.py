"""fprop / dgrad of the small-lattice layers, a few launches each (for ncu's launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_mvd_seg_b200 as m
from multimodal_mvd_seg_b200 import ops
dev = torch.device('cuda:0')
geom = ops.ConvGeom((3,) * 3, (1,) * 3, (1,) * 3)
for cin, cout, E in ((320, 320, 4), (320, 320, 8), (640, 320, 8)):
    x = torch.randn((2, E, E, E, cin), device=dev).to(torch.bfloat16)
    y = torch.randn((2, E, E, E, cout), device=dev).to(torch.bfloat16)
    w = torch.randn((cout, cin, 3, 3, 3), device=dev) * 0.05
    wf, wd = ops.pack_weights(w)
    for which in ('fprop', 'dgrad'):
        ts = []
        for _ in range(6):
            m.lib.spin(300000, torch.cuda.current_stream().cuda_stream)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if which == 'fprop':
                ops.conv_fprop(geom, x, y, wf)
            else:
                ops.conv_dgrad(geom, x, y, wd)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(which, cin, cout, E, 'median %.1f us' % (sorted(ts)[3] * 1e3))
