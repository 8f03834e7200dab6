"""Do an HBM-bound InstanceNorm kernel and a tensor-core conv kernel overlap when launched on two streams?
Prints t(conv), t(norm), t(both concurrently).  usage: overlap_probe.py [fprop|dgrad|wgrad]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_mvd_seg_b200 as m
from multimodal_mvd_seg_b200 import ops
lib = m.lib
which = sys.argv[1] if len(sys.argv) > 1 else 'fprop'
dev = torch.device('cuda:0')
B, D, C = 2, 128, 32
geom = ops.ConvGeom((3,) * 3, (1,) * 3, (1,) * 3)
x = torch.randn((B, D, D, D, C), device=dev).to(torch.bfloat16)
y = torch.randn((B, D, D, D, C), device=dev).to(torch.bfloat16)
w = torch.randn((C, C, 3, 3, 3), device=dev) * 0.05
wf, wd = ops.pack_weights(w)
dw = torch.empty_like(w)
yn = torch.randn((B, D, D, D, C), device=dev).to(torch.bfloat16)
zn = torch.empty_like(yn)
stats = torch.zeros((B, C, 2), dtype=torch.float64, device=dev)
gamma = torch.ones(C, device=dev); beta = torch.zeros(C, device=dev)
V = D * D * D
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
lib.inorm_stats(yn.data_ptr(), C, B, V, C, stats.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()

def conv():
    if which == 'fprop': ops.conv_fprop(geom, x, y, wf)
    elif which == 'dgrad': ops.conv_dgrad(geom, x, y, wd)
    else: ops.conv_wgrad(geom, x, y, dw)

def norm():
    lib.inorm_lrelu_fwd(yn.data_ptr(), C, zn.data_ptr(), C, stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), B, V, C,
                        1e-5, 0.01, torch.cuda.current_stream().cuda_stream)

def timed(fa, fb, n=10):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ev = torch.cuda.Event(); ev.record()
        if fa:
            with torch.cuda.stream(s1):
                s1.wait_event(ev); fa()
                ea = torch.cuda.Event(); ea.record(s1)
            torch.cuda.current_stream().wait_event(ea)
        if fb:
            with torch.cuda.stream(s2):
                s2.wait_event(ev); fb()
                eb = torch.cuda.Event(); eb.record(s2)
            torch.cuda.current_stream().wait_event(eb)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]

for _ in range(3):
    conv(); norm()
tc, tn, tb = timed(conv, None), timed(None, norm), timed(conv, norm)
print(f'{which}: conv {tc:.3f} ms, inorm_lrelu_fwd {tn:.3f} ms, both on two streams {tb:.3f} ms (sum {tc + tn:.3f}, max {max(tc, tn):.3f})')
