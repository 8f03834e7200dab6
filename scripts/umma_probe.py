"""Runs mvd_tc_probe over descriptor variants and reports which shared-memory rows the tensor core fetched."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import multimodal_mvd_seg_b200 as m
from probes import _probe_lib as probe
lib = m.lib
dev = torch.device('cuda:0')
st = torch.cuda.current_stream().cuda_stream
ident = torch.zeros((16, 64), dtype=torch.bfloat16, device=dev)
for i in range(16):
    ident[i, i] = 1

def run(row_bytes, start, sbo, lbo, bo, mn, kadv):
    C = row_bytes // 2
    res = []
    for mode in ('row', 'ch'):
        if mode == 'row':
            src = torch.arange(256, dtype=torch.float32)[:, None].expand(256, C)
        else:
            src = torch.arange(C, dtype=torch.float32)[None, :].expand(256, C)
        src = src.contiguous().to(torch.bfloat16).to(dev)
        out = torch.full((2, 128, 16), -1.0, dtype=torch.float32, device=dev)
        probe.tc_probe(src.data_ptr(), ident.data_ptr(), row_bytes, start, sbo, lbo, bo, mn, kadv, out.data_ptr(), st)
        torch.cuda.synchronize()
        res.append(out.cpu())
    return res  # [row-probe, ch-probe], each [2,128,16]

def describe(tag, row_bytes, start, sbo, lbo, bo, mn, kadv):
    rows, chs = run(row_bytes, start, sbo, lbo, bo, mn, kadv)
    C = row_bytes // 2
    if not mn:
        # K-major: D[m][n] = A[row(m)][ch n]; expected under the address-based model:
        exp_row = torch.tensor([(start + (mm // 8) * sbo + (mm % 8) * row_bytes) // row_bytes for mm in range(128)], dtype=torch.float32)
        exp_ch0 = torch.arange(16, dtype=torch.float32) + ((start % row_bytes) // 2)
        ok_row = bool((rows[0] == exp_row[:, None]).all())
        ok_ch = bool((chs[0] == exp_ch0[None, :]).all())
        exp_ch1 = exp_ch0 + kadv // 2
        ok_k = bool((rows[1] == exp_row[:, None]).all() and (chs[1] == exp_ch1[None, :]).all())
        print(f'{tag:58s} rows {"OK " if ok_row else "BAD"} chans {"OK " if ok_ch else "BAD"} k-adv {"OK " if ok_k else "BAD"}'
              f'  m0..9 rows={rows[0][:10,0].int().tolist()} ch(m=1)={chs[0][1,:4].int().tolist()} ch(m=9)={chs[0][9,:4].int().tolist()}')
    else:
        # MN-major: D[m][n] = A[krow(n)][ch m]: expected krow(n) = (start + (n//8)*sbo + (n%8)*row_bytes)/row_bytes, ch = m (via lbo blocks)
        exp_k = torch.tensor([(start + (n // 8) * sbo + (n % 8) * row_bytes) // row_bytes for n in range(16)], dtype=torch.float32)
        ok_row = bool((rows[0][:C] == exp_k[None, :]).all())
        exp_ch = torch.arange(C, dtype=torch.float32)
        ok_ch = bool((chs[0][:C] == exp_ch[:, None]).all())
        exp_k1 = exp_k + kadv // row_bytes
        ok_k = bool((rows[1][:C] == exp_k1[None, :]).all())
        print(f'{tag:58s} krows {"OK " if ok_row else "BAD"} chans {"OK " if ok_ch else "BAD"} k-adv {"OK " if ok_k else "BAD"}'
              f'  n0..15 krows(m=0)={rows[0][0].int().tolist()} ch(m=0..3,n=1)={chs[0][:4,1].int().tolist()}')

for rb in (128, 64):
    g = 8 * rb
    print(f'==== K-major, row_bytes {rb} (8-row group = {g} B)')
    describe('baseline start=0 sbo=group', rb, 0, g, 16, 0, 0, 32)
    describe('shift 1 row, bo=0', rb, rb, g, 16, 0, 0, 32)
    describe('shift 1 row, bo=1', rb, rb, g, 16, 1, 0, 32)
    describe('shift 3 rows, bo=0', rb, 3 * rb, g, 16, 0, 0, 32)
    describe('shift 3 rows, bo=(addr>>7)&7', rb, 3 * rb, g, 16, (3 * rb >> 7) & 7, 0, 32)
    describe('pitch 16 rows (sbo=2*group)', rb, 0, 2 * g, 16, 0, 0, 32)
    describe('pitch 16 rows, shift 17 rows, bo=0', rb, 17 * rb, 2 * g, 16, 0, 0, 32)
    describe('pitch 16 rows, shift 17 rows, bo=(addr>>7)&7', rb, 17 * rb, 2 * g, 16, (17 * rb >> 7) & 7, 0, 32)
    describe('pitch 10 rows, start=0, bo=0', rb, 0, 10 * rb, 16, 0, 0, 32)
    describe('pitch 10 rows, shift 11 rows, bo=0', rb, 11 * rb, 10 * rb, 16, 0, 0, 32)
    describe('pitch 10 rows, shift 11 rows, bo=(addr>>7)&7', rb, 11 * rb, 10 * rb, 16, (11 * rb >> 7) & 7, 0, 32)
    print(f'==== MN-major (A rows = K), row_bytes {rb}')
    describe('baseline start=0 sbo=group', rb, 0, g, 0, 0, 1, 2 * g)
    describe('shift 1 k-row, bo=0', rb, rb, g, 0, 0, 1, 2 * g)
    describe('shift 3 k-rows, bo=0', rb, 3 * rb, g, 0, 0, 1, 2 * g)
    describe('pitch 10 rows, shift 11 rows, bo=0', rb, 11 * rb, 10 * rb, 0, 0, 1, 20 * rb)
    describe('pitch 16 rows, shift 17 rows, bo=0', rb, 17 * rb, 16 * rb, 0, 0, 1, 32 * rb)
