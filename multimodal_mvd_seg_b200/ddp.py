"""Data-parallel gradient exchange for the batch-sharded training step.

The reference wraps the network in torch DistributedDataParallel (MVDTrainer.py:236-238): a bucketed all-reduce(mean)
of all gradients over NCCL, overlapped with backward.  Here the same exchange is driven directly:

  * every parameter's gradient is written by the wgrad kernels straight into a flat fp32 arena (no copy-in);
    the arena is laid out in REVERSE forward order so that the regions that finish first in backward (hi-res decoder)
    come first and buckets are contiguous ranges;
  * as soon as backward reports that all parameters of a bucket are written (ops.add_param_grad_ready_hook), an
    ``all_reduce(SUM)`` of that range is enqueued on a side stream behind an event, overlapping the remaining backward;
  * the 1/world_size mean is folded into the optimiser kernel (SGDNesterovClip.step(grad_scale=1/world)).

Works with any torch.distributed backend (NCCL on the GPUs; gloo in the CPU tests of the bucketing logic).
Per-rank batch split and oversampling follow MVDTrainer._set_batch_size_and_oversample (MVDTrainer.py:316-361).
"""
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def split_batch_for_rank(global_batch_size: int, world_size: int, rank: int,
                         oversample_foreground_percent: float = 0.33) -> Tuple[int, float]:
    """(batch_size, oversample_percent) of one rank -- MVDTrainer.py:316-361."""
    assert global_batch_size >= world_size, \
        'Cannot run DDP if the batch size is smaller than the number of GPUs... Duh.'
    batch_size_per_GPU = int(np.ceil(global_batch_size / world_size))
    batch_sizes, oversample_percents = [], []
    for r in range(world_size):
        if (r + 1) * batch_size_per_GPU > global_batch_size:
            batch_size = batch_size_per_GPU - ((r + 1) * batch_size_per_GPU - global_batch_size)
        else:
            batch_size = batch_size_per_GPU
        batch_sizes.append(batch_size)
        sample_id_low = 0 if len(batch_sizes) == 0 else np.sum(batch_sizes[:-1])
        sample_id_high = np.sum(batch_sizes)
        if sample_id_high / global_batch_size < (1 - oversample_foreground_percent):
            oversample_percents.append(0.0)
        elif sample_id_low / global_batch_size > (1 - oversample_foreground_percent):
            oversample_percents.append(1.0)
        else:
            covered = sample_id_high / global_batch_size - sample_id_low / global_batch_size
            oversample_percents.append(
                1 - (((1 - oversample_foreground_percent) - sample_id_low / global_batch_size) / covered))
    return int(batch_sizes[rank]), float(oversample_percents[rank])


class GradArena:
    """flat fp32 gradient storage + bucketed, backward-overlapped all-reduce."""

    def __init__(self, params: Sequence[torch.nn.Parameter], bucket_bytes: int = 32 << 20,
                 process_group=None, world_size: Optional[int] = None):
        self.params = [p for p in params if p.requires_grad]
        self.group = process_group
        self.distributed = dist.is_available() and dist.is_initialized() and \
            (world_size if world_size is not None else dist.get_world_size(process_group)) > 1
        self.world_size = dist.get_world_size(process_group) if self.distributed else 1
        dev = self.params[0].device
        order = list(reversed(self.params))  # backward order
        self.offset: Dict[int, Tuple[int, int]] = {}
        off = 0
        for p in order:
            n = p.numel()
            self.offset[id(p)] = (off, n)
            off += (n + 3) // 4 * 4  # keep every view 16-byte aligned
        self.flat = torch.zeros((off,), dtype=torch.float32, device=dev)
        # buckets: contiguous ranges of ~bucket_bytes in backward order
        self.buckets: List[dict] = []
        cur = dict(start=0, end=0, ids=set())
        limit = max(1, bucket_bytes // 4)
        for p in order:
            o, n = self.offset[id(p)]
            cur['ids'].add(id(p))
            cur['end'] = o + (n + 3) // 4 * 4
            if cur['end'] - cur['start'] >= limit:
                self.buckets.append(cur)
                cur = dict(start=cur['end'], end=cur['end'], ids=set())
        if cur['ids']:
            self.buckets.append(cur)
        self.bucket_of = {i: b for b, bk in enumerate(self.buckets) for i in bk['ids']}
        self._views = {id(p): self.flat[o:o + n].view(p.shape) for p in self.params for (o, n) in [self.offset[id(p)]]}
        self.comm_stream = torch.cuda.Stream(device=dev) if dev.type == 'cuda' else None
        self.begin_step()

    # -- allocator handed to ops.set_grad_allocator ---------------------------------------------------------------
    def view_for(self, p: torch.Tensor) -> Optional[torch.Tensor]:
        return self._views.get(id(p))

    # -- step protocol ------------------------------------------------------------------------------------------
    def begin_step(self):
        self._ready = [set() for _ in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._handles = []

    def on_params_ready(self, params):
        """called from backward (autograd thread) when the gradients of `params` have been enqueued."""
        touched = set()
        for p in params:
            b = self.bucket_of.get(id(p))
            if b is None:
                continue
            self._ready[b].add(id(p))
            touched.add(b)
        if not self.distributed:
            return
        for b in sorted(touched):
            if not self._launched[b] and len(self._ready[b]) == len(self.buckets[b]['ids']):
                self._launch(b)

    def _launch(self, b: int):
        bk = self.buckets[b]
        self._launched[b] = True
        chunk = self.flat[bk['start']:bk['end']]
        if self.comm_stream is not None:
            from . import ops   # deferred weight gradients run on ops' side stream: the exchange must follow them too
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            side = ops.side_stream(chunk.device)
            ev_side = None
            if side is not None:
                ev_side = torch.cuda.Event()
                ev_side.record(side)
            with torch.cuda.stream(self.comm_stream):
                self.comm_stream.wait_event(ev)
                if ev_side is not None:
                    self.comm_stream.wait_event(ev_side)
                self._handles.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            self._handles.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """after backward: zero the regions of parameters that received no gradient, flush the remaining buckets and
        make the compute stream wait for the exchange."""
        for b, bk in enumerate(self.buckets):
            missing = bk['ids'] - self._ready[b]
            if missing:
                for p in self.params:
                    if id(p) in missing:
                        o, n = self.offset[id(p)]
                        self.flat[o:o + n].zero_()
            if self.distributed and not self._launched[b]:
                self._launch(b)
        for h in self._handles:
            h.wait()
        if self.comm_stream is not None and self._handles:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        self._handles = []

    def attach_grads(self):
        """point every parameter's .grad at its arena view.  The arena is the source of truth: the wgrad kernels wrote
        into it and the all-reduce ran on it (autograd may have kept a private copy of the pre-reduction values; a
        parameter autograd never touched has a zeroed region).  Requires zero_grad(set_to_none=True) before backward,
        as the reference does (nnUNetTrainer.py:901)."""
        for p in self.params:
            p.grad = self._views[id(p)]
