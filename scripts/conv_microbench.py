"""GPU microbenchmark of one conv layer through the C ABI (CUDA-event timing, inputs larger than L2 or L2 flushed).
usage: conv_microbench.py CIN COUT D H W [pass=fprop|dgrad|wgrad] [stride=1] [iters=10] [B=2] [k=3]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_mvd_seg_b200 as m
from multimodal_mvd_seg_b200 import ops

a = sys.argv[1:]
cin, cout, D, H, W = (int(v) for v in a[:5])
which = a[5] if len(a) > 5 else 'fprop'
s = int(a[6]) if len(a) > 6 else 1
iters = int(a[7]) if len(a) > 7 else 10
B = int(a[8]) if len(a) > 8 else 2
k = int(a[9]) if len(a) > 9 else 3
dev = torch.device('cuda:0')
geom = ops.ConvGeom((k,) * 3, (s,) * 3, ((k - 1) // 2,) * 3)
Do, Ho, Wo = geom.out_size((D, H, W))
x = torch.randn((B, D, H, W, cin), device=dev).to(torch.bfloat16)
y = torch.randn((B, Do, Ho, Wo, cout), device=dev).to(torch.bfloat16)
w = torch.randn((cout, cin, k, k, k), device=dev) * 0.05
bias = torch.zeros(cout, device=dev)
wf, wd = ops.pack_weights(w)
dw = torch.empty_like(w)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
nb = bstats = None
if os.environ.get('NB') == '1' and which == 'dgrad':   # + backward sums of the norm in front (mvd_conv3d_args.norm_bwd)
    y_prev = torch.randn((B, D, H, W, cin), device=dev).to(torch.bfloat16)
    fstats = torch.zeros((B, cin, 2), dtype=torch.float64, device=dev)
    m.lib.inorm_stats(y_prev.data_ptr(), cin, B, D * H * W, cin, fstats.data_ptr(), torch.cuda.current_stream().cuda_stream)
    nb = (y_prev, fstats, torch.ones(cin, device=dev), torch.zeros(cin, device=dev), 1e-5, 0.01)
    bstats = torch.zeros((B, cin, 2), dtype=torch.float64, device=dev)
def run():
    if which == 'fprop':
        ops.conv_fprop(geom, x, y, wf, bias=bias)
    elif which == 'dgrad':
        if nb is not None:
            bstats.zero_()
            ops.conv_dgrad(geom, x, y, wd, norm_bwd=nb, bstats=bstats)
        else:
            ops.conv_dgrad(geom, x, y, wd)
    else:
        ops.conv_wgrad(geom, x, y, dw)
for _ in range(3):
    run()
torch.cuda.synchronize()
ts = []
for _ in range(iters):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts.sort()
flops = 2.0 * B * Do * Ho * Wo * cout * cin * k ** 3
med = ts[len(ts) // 2]
print(f'{which} {cin}->{cout} {D}x{H}x{W} s{s} k{k} B{B}: median {med:.3f} ms  best {ts[0]:.3f} ms  {flops / med / 1e9:.1f} TFLOP/s (median)')
if os.environ.get('HALO_PROF') == '1' and which in ('fprop', 'dgrad'):
    import ctypes
    cd = ctypes.CDLL(m.LIB_PATH)
    buf = torch.zeros((148, 8), dtype=torch.int64, device=dev)
    cd.mvd_debug_set_halo_prof(ctypes.c_void_p(buf.data_ptr()))
    run()
    torch.cuda.synchronize()
    cd.mvd_debug_set_halo_prof(ctypes.c_void_p(0))
    b = buf.float().mean(0)
    print(f'  MMA issuer cycles/CTA: total {b[0]:.0f}  wait tmem-empty {b[1]:.0f}  wait weights {b[2]:.0f}  wait planes {b[3]:.0f}  '
          f'issue+other {b[0] - b[1] - b[2] - b[3]:.0f}')
