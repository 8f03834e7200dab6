"""tcgen05.mma issue-rate microbenchmark (cycles per M128 x N x K16 MMA) for several operand layouts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import multimodal_mvd_seg_b200 as m
from probes import _probe_lib as probe
dev = torch.device('cuda:0')
st = torch.cuda.current_stream().cuda_stream
out = torch.zeros(4, dtype=torch.int64, device=dev)
def run(n, n_acc, a_sbo, a_step, mod, rb=128, mn=0, b_step=32, grid=1, n_mma=2048, mode=1):
    probe.tc_mma_bench(n, n_mma, n_acc, a_sbo, a_step, mod, rb, mn, b_step, grid, mode, out.data_ptr(), st)
    torch.cuda.synchronize()
    return float(out[0]) / n_mma
print('cycles per MMA (M=128, K=16); ideal = N/2')
for grid in (1, 148):
    print(f'--- grid {grid}')
    for n in (32, 64, 128, 256):
        div = run(n, 1, 1024, 32, 4, grid=grid, mode=0)
        el1 = run(n, 1, 1024, 32, 4, grid=grid, mode=2)
        el4 = run(n, min(4, 512 // n), 1280, 128, 16, grid=grid, mode=2)
        print(f"N={n:3d}: elect-once loop {el1:6.1f} (halo-style 4acc {el4:6.1f}) ||", end=" ")
        base = run(n, 1, 1024, 32, 4, grid=grid)
        acc2 = run(n, min(2, 512 // n), 1024, 32, 4, grid=grid)
        acc4 = run(n, min(4, 512 // n), 1024, 32, 4, grid=grid)
        same = run(n, 1, 1024, 0, 1, b_step=0, grid=grid)
        halo = run(n, min(4, 512 // n), 1280, 128, 16, grid=grid)
        sw64 = run(n, 1, 512, 32, 2, rb=64, grid=grid)
        sw64h = run(n, min(4, 512 // n), 640, 64, 16, rb=64, grid=grid)
        mnm = run(n, 1, 1024, 2048, 8, mn=1, grid=grid)
        print(f"N={n:3d}: divergent-lane issue {div:6.1f} ||", end=" ")
        print(f'N={n:3d}: K-major sw128 1acc {base:6.1f} | 2acc {acc2:6.1f} | 4acc {acc4:6.1f} | same operands {same:6.1f} | '
              f'halo pitch(1280)+shifts 4acc {halo:6.1f} | sw64 {sw64:6.1f} | sw64 halo {sw64h:6.1f} | MN-major {mnm:6.1f}')
