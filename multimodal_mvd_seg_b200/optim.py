"""Optimiser tail of train_step on libmvdseg:

    torch.nn.utils.clip_grad_norm_(network.parameters(), 12); optimizer.step()     (MVDTrainer.py:978-979)
    optimizer = torch.optim.SGD(params, lr, weight_decay=3e-5, momentum=0.99, nesterov=True)   (:482-484)
    lr_scheduler = PolyLRScheduler(optimizer, initial_lr, num_epochs)              (lr_scheduler/polylr.py:4-20)

``SGDNesterovClip`` is a torch.optim.Optimizer (so ``param_groups[0]['lr']``, ``state_dict`` and the scheduler work as
in the reference) whose ``step`` is two multi-tensor kernels: gradient square-norm, then clip+weight-decay+nesterov.
No host synchronisation: the clip coefficient is computed on the device.
"""
from typing import Iterable, List, Optional

import torch

from ._lib import MvdError, lib


def _timed_mem(name, nbytes, fn, *a):
    from . import ops      # bench.py's per-kernel event timer (no-op unless armed)
    return ops._timed_mem(name, nbytes, fn, *a)

_CHUNK = 4096


class SGDNesterovClip(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float, weight_decay: float = 3e-5,
                 momentum: float = 0.99, nesterov: bool = True, max_norm: Optional[float] = 12.0):
        if not nesterov:
            raise NotImplementedError('the reference trains with nesterov=True (MVDTrainer.py:483)')
        # the groups carry exactly torch.optim.SGD's keys, so `optimizer_state` of a checkpoint interchanges with the
        # reference's optimiser in both directions (MVDTrainer.py:1138, 1180); the clipping threshold -- an argument
        # of clip_grad_norm_ in the reference, not optimiser state -- is an attribute
        defaults = dict(lr=lr, momentum=momentum, dampening=0, weight_decay=weight_decay, nesterov=True,
                        maximize=False, foreach=None, differentiable=False, fused=None)
        super().__init__(params, defaults)
        self.max_norm = max_norm
        self._tables = {}      # group index -> cached device tables
        self.last_sqnorm = None  # device double[1]: squared global grad norm of the last step (before clipping)
        self._packers = []     # ops.WeightPacker objects whose weights are updated AND re-packed by one fused kernel

    def attach_weight_packers(self, packers):
        """conv weights that live in these packers are updated by mvd_sgd_pack_conv_weights, which also rewrites their
        bf16 GEMM layouts: the next forward pass skips its weight-pack launch."""
        self._packers = [p for p in packers if p is not None]
        self._tables = {}

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._tables = {}      # the momentum buffers were replaced: cached device pointer tables are stale
        for g in self.param_groups:
            if not g.get('nesterov', True) or g.get('dampening', 0) != 0 or g.get('maximize', False):
                raise NotImplementedError('SGDNesterovClip: only nesterov=True, dampening=0, maximize=False is built')
            g.pop('max_norm', None)      # round-1 checkpoints of this package kept it in the groups
            for k, v in self.defaults.items():
                g.setdefault(k, v)

    def ensure_state(self):
        """create the momentum buffers now (first-step laziness must not be captured into a CUDA graph: a captured
        zero-initialisation would reset the momentum on every replay)."""
        for g in self.param_groups:
            for p in g['params']:
                st = self.state[p]
                if st.get('momentum_buffer') is None:
                    st['momentum_buffer'] = torch.zeros_like(p, memory_format=torch.contiguous_format)

    def _group_tables(self, gi: int, group) -> dict:
        params = [p for p in group['params']]
        for p in params:
            if not p.is_cuda:
                raise MvdError('SGDNesterovClip: parameters must live on a CUDA device; libmvdseg has no CPU path')
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise MvdError('SGDNesterovClip: parameters must be contiguous fp32')
            st = self.state[p]
            if st.get('momentum_buffer') is None:
                st['momentum_buffer'] = torch.zeros_like(p, memory_format=torch.contiguous_format)
            elif st['momentum_buffer'].dtype != torch.float32 or not st['momentum_buffer'].is_contiguous():
                st['momentum_buffer'] = st['momentum_buffer'].float().contiguous()
        sig = []
        for p in params:
            g = p.grad
            if g is None:
                # a parameter that received no gradient behaves as if its gradient were zero (what the reference
                # gets from `0 * loss(...)` on the zero-weighted deep-supervision scale): weight decay and
                # momentum still apply
                p.grad = g = torch.zeros_like(p)
            elif g.dtype != torch.float32 or not g.is_contiguous():
                p.grad = g = g.float().contiguous()
            sig.append((p.data_ptr(), g.data_ptr(), self.state[p]['momentum_buffer'].data_ptr()))
        psig = tuple(id(pk) for pk in self._packers)
        cached = self._tables.get(gi)
        if cached is not None and cached['sig'] == sig and cached['psig'] == psig:
            return cached
        dev = params[0].device
        numel = [p.numel() for p in params]
        # conv weights owned by an attached packer: their update is fused with the re-pack (one table row each)
        import numpy as np
        from .ops import WeightPacker
        fused_rows, fused_idx, blocks = [], set(), 0
        index_of = {id(p): i for i, p in enumerate(params)}
        for pk in self._packers:
            for w, row in zip(pk._keep, pk.rows):
                i = index_of.get(id(w))
                if i is None:
                    continue
                r = row.copy()
                r['block_begin'] = blocks
                blocks += lib.pack_blocks(int(r['Cout']), int(r['Cin']))
                fused_rows.append((r, sig[i][1], sig[i][2]))
                fused_idx.add(i)
        if any(not pk_all for pk_all in [all(id(w) in index_of for w in pk._keep) for pk in self._packers]):
            fused_rows, fused_idx, blocks = [], set(), 0      # a packer with foreign weights cannot be kept fresh
        fused_table = None
        if fused_rows:
            dt = np.dtype(WeightPacker._DESC.descr + [('grad', '<u8'), ('mom', '<u8')])
            arr = np.zeros(len(fused_rows), dtype=dt)
            for j, (r, gptr, mptr) in enumerate(fused_rows):
                for name in WeightPacker._DESC.names:
                    arr[j][name] = r[name]
                arr[j]['grad'], arr[j]['mom'] = gptr, mptr
            fused_table = torch.from_numpy(arr.view(np.uint8).copy()).to(dev)

        def chunks(indices):
            ct, co = [], []
            for i in indices:
                for off in range(0, numel[i], _CHUNK):
                    ct.append(i)
                    co.append(off)
            return (torch.tensor(ct, dtype=torch.int32, device=dev), torch.tensor(co, dtype=torch.int64, device=dev), len(ct))
        all_ct, all_co, all_n = chunks(range(len(params)))
        plain = [i for i in range(len(params)) if i not in fused_idx]
        pl_ct, pl_co, pl_n = chunks(plain) if fused_idx else (all_ct, all_co, all_n)
        # uint64 pointers stored as int64 bit patterns
        flat = [v - (1 << 64) if v >= (1 << 63) else v for row in sig for v in row]
        cached = dict(sig=sig, psig=psig,
                      ptrs=torch.tensor(flat, dtype=torch.int64, device=dev),
                      numel=torch.tensor(numel, dtype=torch.int64, device=dev),
                      chunk_t=all_ct, chunk_o=all_co, n_chunks=all_n, n_elems=int(sum(numel)),
                      plain_chunk_t=pl_ct, plain_chunk_o=pl_co, plain_n_chunks=pl_n,
                      plain_elems=int(sum(numel[i] for i in plain)),
                      fused_table=fused_table, fused_n=len(fused_rows), fused_blocks=blocks,
                      fused_elems=int(sum(numel[i] for i in fused_idx)))
        self._tables[gi] = cached
        return cached

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        """grad_scale multiplies every gradient before clipping (the 1/world_size of a summed DDP all-reduce)."""
        assert closure is None
        stream = torch.cuda.current_stream().cuda_stream
        tabs = [self._group_tables(gi, g) for gi, g in enumerate(self.param_groups)]
        dev = self.param_groups[0]['params'][0].device
        from . import ops
        sq = ops.zeros((1,), torch.float64, dev)      # from the per-step zero pool (no fill kernel)
        need_norm = self.max_norm is not None and self.max_norm > 0
        if need_norm:  # the norm is global over all groups, like clip_grad_norm_(network.parameters())
            for t in tabs:
                _timed_mem('grad_sqnorm', 4.0 * t['n_elems'], lib.grad_sqnorm, t['ptrs'].data_ptr(), t['numel'].data_ptr(),
                           t['chunk_t'].data_ptr(), t['chunk_o'].data_ptr(), t['n_chunks'], sq.data_ptr(), stream)
        fused_any = False
        for t, g in zip(tabs, self.param_groups):
            mn = self.max_norm if need_norm else 0.0
            if t['plain_n_chunks']:
                _timed_mem('sgd_nesterov_clip', 20.0 * t['plain_elems'], lib.sgd_nesterov_clip, t['ptrs'].data_ptr(),
                           t['numel'].data_ptr(), t['plain_chunk_t'].data_ptr(), t['plain_chunk_o'].data_ptr(),
                           t['plain_n_chunks'], sq.data_ptr(), float(grad_scale), float(mn), float(g['lr']),
                           float(g['weight_decay']), float(g['momentum']), stream)
            if t['fused_table'] is not None:     # conv weights: update + refresh of the bf16 GEMM layouts in one kernel
                _timed_mem('sgd_pack_conv_weights', 24.0 * t['fused_elems'], lib.sgd_pack_conv_weights,
                           t['fused_table'].data_ptr(), t['fused_n'], t['fused_blocks'], sq.data_ptr(), float(grad_scale),
                           float(mn), float(g['lr']), float(g['weight_decay']), float(g['momentum']), stream)
                fused_any = True
        for pk in self._packers:
            if fused_any:
                pk.mark_fresh()
            else:
                pk.invalidate()
        self.last_sqnorm = sq
        return None


class PolyLRScheduler:
    """lr_scheduler/polylr.py:4-20; only ``step(current_step)`` is used (MVDTrainer.py:869-877)."""

    def __init__(self, optimizer, initial_lr: float, max_steps: int, exponent: float = 0.9, current_step: int = None):
        self.optimizer = optimizer
        self.initial_lr = initial_lr
        self.max_steps = max_steps
        self.exponent = exponent
        self.ctr = 0

    def step(self, current_step=None):
        if current_step is None or current_step == -1:
            current_step = self.ctr
            self.ctr += 1
        new_lr = self.initial_lr * (1 - current_step / self.max_steps) ** self.exponent
        for param_group in self.optimizer.param_groups:
            param_group['lr'] = new_lr
