"""two weight-gradient launches for ncu: 32->32 @ 2x128^3 (the step's dominant kernel) and 256->256 @ 2x16^3."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_mvd_seg_b200 as m
from multimodal_mvd_seg_b200 import ops
dev = torch.device('cuda:0')
geom = ops.ConvGeom((3,) * 3, (1,) * 3, (1,) * 3)
for cin, cout, E in ((32, 32, 128), (256, 256, 16)):
    x = torch.randn((2, E, E, E, cin), device=dev).to(torch.bfloat16)
    y = torch.randn((2, E, E, E, cout), device=dev).to(torch.bfloat16)
    dw = torch.empty((cout, cin, 3, 3, 3), device=dev)
    for _ in range(2):
        ops.conv_wgrad(geom, x, y, dw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.conv_wgrad(geom, x, y, dw); e1.record(); torch.cuda.synchronize()
    print(cin, cout, E, e0.elapsed_time(e1), 'ms')
