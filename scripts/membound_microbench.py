"""CUDA-event microbenchmark of the HBM-bound kernels (achieved GB/s on ALGORITHMIC bytes, inputs larger than L2)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import multimodal_mvd_seg_b200 as m
lib = m.lib
dev = torch.device('cuda:0')
st = torch.cuda.current_stream().cuda_stream
BF = torch.bfloat16
peak = 6538.0
try:
    peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def timeit(fn, iters=7):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]

def report(name, ms, nbytes):
    gbs = nbytes / ms / 1e6
    print(f'{name:44s} {ms:8.3f} ms  {nbytes / 1e6:9.1f} MB  {gbs:8.1f} GB/s  {100 * gbs / peak:5.1f}% of measured HBM peak ({peak:.0f} GB/s)')

B, D, H, W, C = 2, 128, 128, 128, 32
V = D * H * W
y = torch.randn((B, D, H, W, C), device=dev).to(BF)
dz = torch.randn((B, D, H, W, C), device=dev).to(BF)
z = torch.empty_like(y); dy = torch.empty_like(y)
gamma = torch.ones(C, device=dev); beta = torch.zeros(C, device=dev)
stats = torch.zeros((B, C, 2), dtype=torch.float64, device=dev)
bstats = torch.zeros((B, C, 2), dtype=torch.float64, device=dev)
dg = torch.empty(C, device=dev); db = torch.empty(C, device=dev); dsum = torch.zeros(C, device=dev)
n = B * V * C
lib.inorm_stats(y.data_ptr(), C, B, V, C, stats.data_ptr(), st)
report('inorm_stats (C=32, 2x128^3)', timeit(lambda: lib.inorm_stats(y.data_ptr(), C, B, V, C, stats.data_ptr(), st)), n * 2)
report('inorm_lrelu_fwd', timeit(lambda: lib.inorm_lrelu_fwd(y.data_ptr(), C, z.data_ptr(), C, stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), B, V, C, 1e-5, 0.01, st)), n * 4)
report('inorm_lrelu_bwd_stats', timeit(lambda: lib.inorm_lrelu_bwd_stats(dz.data_ptr(), C, y.data_ptr(), C, stats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), B, V, C, 1e-5, 0.01, bstats.data_ptr(), st)), n * 4)
report('inorm_lrelu_bwd_apply (+bias-grad sum)', timeit(lambda: lib.inorm_lrelu_bwd_apply(dz.data_ptr(), C, y.data_ptr(), C, dy.data_ptr(), C, stats.data_ptr(), bstats.data_ptr(), gamma.data_ptr(), beta.data_ptr(), B, V, C, 1e-5, 0.01, dg.data_ptr(), db.data_ptr(), dsum.data_ptr(), st)), n * 6)
# losses on the hi-res scale (C = 4 classes)
K = 4
logits = torch.randn((B, D, H, W, K), device=dev).to(BF)
logits2 = torch.randn((B, D, H, W, K), device=dev).to(BF)
target = torch.randint(0, K, (B, V), device=dev).float()
acc = torch.zeros(B * K * 3 + 1, dtype=torch.float64, device=dev)
coef = torch.zeros((B, K, 2), device=dev); gout = torch.ones(1, device=dev); dl = torch.empty_like(logits); dl2 = torch.empty_like(logits)
nv = B * V
report('dice_ce_fwd (C=4, 2x128^3)', timeit(lambda: lib.dice_ce_fwd(logits.data_ptr(), K, target.data_ptr(), B, V, K, acc.data_ptr(), st)), nv * (K * 2 + 4))
report('dice_ce_bwd', timeit(lambda: lib.dice_ce_bwd(logits.data_ptr(), K, target.data_ptr(), B, V, K, coef.data_ptr(), 1.0, 1.0, gout.data_ptr(), dl.data_ptr(), K, st)), nv * (K * 2 + 4 + K * 2))
ls = torch.zeros(1, dtype=torch.float64, device=dev)
report('kl_fwd', timeit(lambda: lib.kl_fwd(logits.data_ptr(), K, logits2.data_ptr(), K, nv, K, 1.0, ls.data_ptr(), st)), nv * K * 4)
report('kl_bwd (both gradients)', timeit(lambda: lib.kl_bwd(logits.data_ptr(), K, logits2.data_ptr(), K, nv, K, 1.0, 1.0, gout.data_ptr(), dl.data_ptr(), K, dl2.data_ptr(), K, st)), nv * K * 8)
# soft-skeleton level (fp32 volume 2x160x160x96)
Bs, Ds, Hs, Ws = 2, 160, 160, 96
a = torch.rand((Bs, Ds, Hs, Ws), device=dev); b2 = torch.empty_like(a); sk = torch.empty_like(a); dlt = torch.empty_like(a)
ns = a.numel()
report('soft_erode (2x160x160x96 fp32)', timeit(lambda: lib.soft_erode(a.data_ptr(), b2.data_ptr(), Bs, Ds, Hs, Ws, st)), ns * 8)
report('skel_update (dilate+delta+skel)', timeit(lambda: lib.skel_update(a.data_ptr(), b2.data_ptr(), sk.data_ptr(), dlt.data_ptr(), sk.data_ptr(), 0, Bs, Ds, Hs, Ws, st)), ns * 20)
