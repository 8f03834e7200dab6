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
        cached = self._tables.get(gi)
        if cached is not None and cached['sig'] == sig:
            return cached
        dev = params[0].device
        numel = [p.numel() for p in params]
        chunk_t, chunk_o = [], []
        for i, n in enumerate(numel):
            for off in range(0, n, _CHUNK):
                chunk_t.append(i)
                chunk_o.append(off)
        # uint64 pointers stored as int64 bit patterns
        flat = [v - (1 << 64) if v >= (1 << 63) else v for row in sig for v in row]
        cached = dict(sig=sig,
                      ptrs=torch.tensor(flat, dtype=torch.int64, device=dev),
                      numel=torch.tensor(numel, dtype=torch.int64, device=dev),
                      chunk_t=torch.tensor(chunk_t, dtype=torch.int32, device=dev),
                      chunk_o=torch.tensor(chunk_o, dtype=torch.int64, device=dev),
                      n_chunks=len(chunk_t), n_elems=int(sum(numel)))
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
        for t, g in zip(tabs, self.param_groups):
            mn = self.max_norm if need_norm else 0.0
            _timed_mem('sgd_nesterov_clip', 20.0 * t['n_elems'], lib.sgd_nesterov_clip, t['ptrs'].data_ptr(),
                       t['numel'].data_ptr(), t['chunk_t'].data_ptr(), t['chunk_o'].data_ptr(), t['n_chunks'],
                       sq.data_ptr(), float(grad_scale), float(mn), float(g['lr']), float(g['weight_decay']),
                       float(g['momentum']), stream)
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
