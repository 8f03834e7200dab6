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
    """(batch_size, oversample_percent) of one rank, the rule of MVDTrainer._set_batch_size_and_oversample
    (MVDTrainer.py:316-361): every rank takes ceil(global / world) samples, the trailing ranks give back the excess;
    the foreground-oversampled tail of the GLOBAL batch (its last ``oversample_foreground_percent``) is mapped onto the
    ranks whose sample range [lo, hi) reaches into it."""
    if global_batch_size < world_size:
        raise AssertionError('Cannot run DDP if the batch size is smaller than the number of GPUs... Duh.')
    per_rank = -(-global_batch_size // world_size)
    ends = (np.arange(world_size) + 1) * per_rank
    sizes = per_rank - np.maximum(ends - global_batch_size, 0)
    hi = np.cumsum(sizes) / global_batch_size
    lo = hi - sizes / global_batch_size
    plain_share = 1.0 - oversample_foreground_percent     # leading fraction of the global batch that is NOT oversampled
    if hi[rank] < plain_share:
        pct = 0.0
    elif lo[rank] > plain_share:
        pct = 1.0
    else:
        pct = 1.0 - (plain_share - lo[rank]) / (hi[rank] - lo[rank])
    return int(sizes[rank]), float(pct)


def broadcast_parameters(modules: Sequence[torch.nn.Module], src: int = 0, group=None) -> None:
    """what the DistributedDataParallel constructor does for the reference (MVDTrainer.py:236-238): every rank starts
    from rank ``src``'s parameters and buffers.  One flat broadcast per dtype; no-op without an initialised group."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) <= 1:
        return
    seen, tensors = set(), []
    for mod in modules:
        for t in list(mod.parameters()) + list(mod.buffers()):
            if id(t) not in seen:
                seen.add(id(t))
                tensors.append(t.detach())
    by_dtype: Dict[torch.dtype, List[torch.Tensor]] = {}
    for t in tensors:
        by_dtype.setdefault(t.dtype, []).append(t)
    for group_tensors in by_dtype.values():
        flat = torch.cat([t.reshape(-1) for t in group_tensors])
        dist.broadcast(flat, src=src, group=group)
        off = 0
        for t in group_tensors:
            t.copy_(flat[off:off + t.numel()].view_as(t))
            off += t.numel()


def parameter_checksums(modules: Sequence[torch.nn.Module]) -> torch.Tensor:
    """[sum, sum of squares] over all parameters in fp64 (replica-consistency check: bench.py, tests)."""
    acc = None
    for mod in modules:
        for p in mod.parameters():
            v = p.detach().double()
            cur = torch.stack([v.sum(), (v * v).sum()])
            acc = cur if acc is None else acc + cur
    return acc


def replicas_identical(modules: Sequence[torch.nn.Module], group=None) -> bool:
    """all-gathers the parameter checksums; True when every rank holds bit-identical sums."""
    cs = parameter_checksums(modules)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) <= 1:
        return True
    parts = [torch.empty_like(cs) for _ in range(dist.get_world_size(group))]
    dist.all_gather(parts, cs, group=group)
    return all(torch.equal(parts[0], q) for q in parts[1:])


class GradArena:
    """flat fp32 gradient storage + bucketed, backward-overlapped all-reduce."""

    def __init__(self, params: Sequence[torch.nn.Parameter], bucket_bytes: int = 32 << 20,
                 process_group=None, world_size: Optional[int] = None, tail_bytes: int = 2 << 20):
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
        # buckets: contiguous ranges of ~bucket_bytes in backward order, except the LAST one: the parameters whose
        # gradients arrive at the very end of backward (the first encoder stages: a few hundred KB) form a small tail
        # bucket of >= tail_bytes, so that the only all-reduce that cannot hide behind remaining backward work is
        # latency-sized instead of up to bucket_bytes long
        self.buckets: List[dict] = []
        limit = max(1, bucket_bytes // 4)
        tail_limit = max(1, min(tail_bytes, bucket_bytes) // 4)
        split, acc = len(order), 0
        while split > 1 and acc < tail_limit:
            split -= 1
            acc += (order[split].numel() + 3) // 4 * 4
        if acc >= limit or split <= 0:
            split = len(order)           # everything is tiny anyway: no separate tail
        cur = dict(start=0, end=0, ids=set())
        for k, p in enumerate(order):
            if k == split and cur['ids']:
                self.buckets.append(cur)
                cur = dict(start=cur['end'], end=cur['end'], ids=set())
            o, n = self.offset[id(p)]
            cur['ids'].add(id(p))
            cur['end'] = o + (n + 3) // 4 * 4
            if k < split and cur['end'] - cur['start'] >= limit:
                self.buckets.append(cur)
                cur = dict(start=cur['end'], end=cur['end'], ids=set())
        if cur['ids']:
            self.buckets.append(cur)
        self.bucket_of = {i: b for b, bk in enumerate(self.buckets) for i in bk['ids']}
        self._views = {id(p): self.flat[o:o + n].view(p.shape) for p in self.params for (o, n) in [self.offset[id(p)]]}
        self.comm_stream = torch.cuda.Stream(device=dev) if dev.type == 'cuda' else None
        # exchange timeline (bench.py / profiles): when armed, CUDA events bracket every bucket's all-reduce on the comm
        # stream and the compute stream's wait for the exchange in finish(); eager steps only (events are not captured)
        self.timeline: Optional[list] = None
        self.begin_step()

    def arm_timeline(self, on: bool = True):
        self.timeline = [] if on else None

    def timeline_summary(self):
        """per recorded step: bytes, bucket count, summed all-reduce time on the comm stream, span from the first
        all-reduce start to the last end, and the EXPOSED part = how long the compute stream sat in finish() waiting
        for the exchange after backward had been enqueued."""
        if not self.timeline:
            return None
        torch.cuda.synchronize()
        out = []
        for rec in self.timeline:
            if not rec['buckets'] or 'wait0' not in rec:
                continue
            ar = [a.elapsed_time(b) for (a, b, _) in rec['buckets']]
            first, last = rec['buckets'][0][0], rec['buckets'][-1][1]
            out.append(dict(buckets=len(ar), bytes=int(sum(n for (_, _, n) in rec['buckets'])), allreduce_ms_sum=float(sum(ar)),
                            allreduce_span_ms=float(first.elapsed_time(last)),
                            exposed_ms=float(rec['wait0'].elapsed_time(rec['wait1'])),
                            first_allreduce_after_step_start_ms=float(rec['t0'].elapsed_time(first))))
        return out

    # -- allocator handed to ops.set_grad_allocator ---------------------------------------------------------------
    def view_for(self, p: torch.Tensor) -> Optional[torch.Tensor]:
        return self._views.get(id(p))

    def prezero(self, params) -> List[torch.Tensor]:
        """register parameters whose gradient region is cleared at the top of every step (kernels accumulate into them,
        or they may receive no gradient at all); returns their arena views.  One launch clears all of them."""
        seen, rows, views = set(), [], []
        for p in params:
            ent = self.offset.get(id(p))
            if ent is None or id(p) in seen:
                continue
            seen.add(id(p))
            rows.append(ent)
            views.append(self._views[id(p)])
        self._prezero_ids = seen
        self._prezero_rows = rows
        self._prezero_table = None
        if rows and self.flat.is_cuda:
            self._prezero_table = torch.tensor([v for r in rows for v in r], dtype=torch.int64, device=self.flat.device)
        return views

    # -- step protocol ------------------------------------------------------------------------------------------
    def begin_step(self):
        self._ready = [set() for _ in self.buckets]
        self._launched = [False] * len(self.buckets)
        self._handles = []
        if self.timeline is not None and self.comm_stream is not None:
            t0 = torch.cuda.Event(enable_timing=True)
            t0.record()
            self.timeline.append(dict(t0=t0, buckets=[]))
        rows = getattr(self, '_prezero_rows', None)
        if rows:
            if self._prezero_table is not None:
                from ._lib import lib
                lib.zero_regions(self.flat.data_ptr(), self._prezero_table.data_ptr(), len(rows),
                                 torch.cuda.current_stream(self.flat.device).cuda_stream)
            else:
                for o, n in rows:
                    self.flat[o:o + n].zero_()

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
                if self.timeline:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record(self.comm_stream)
                self._handles.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
                if self.timeline:
                    self._handles[-1].wait()          # stream-side wait: orders e1 behind the collective on comm_stream
                    e1.record(self.comm_stream)
                    self.timeline[-1]['buckets'].append((e0, e1, chunk.numel() * 4))
        else:
            self._handles.append(dist.all_reduce(chunk, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self):
        """after backward: zero the regions of parameters that received no gradient, flush the remaining buckets and
        make the compute stream wait for the exchange."""
        for b, bk in enumerate(self.buckets):
            missing = bk['ids'] - self._ready[b] - getattr(self, '_prezero_ids', set())
            if missing:
                for p in self.params:
                    if id(p) in missing:
                        o, n = self.offset[id(p)]
                        self.flat[o:o + n].zero_()
            if self.distributed and not self._launched[b]:
                self._launch(b)
        rec = self.timeline[-1] if self.timeline else None
        if rec is not None and self.comm_stream is not None:
            rec['wait0'] = torch.cuda.Event(enable_timing=True)
            rec['wait0'].record()
        for h in self._handles:
            h.wait()
        if self.comm_stream is not None and self._handles:
            torch.cuda.current_stream().wait_stream(self.comm_stream)
        if rec is not None and self.comm_stream is not None:
            rec['wait1'] = torch.cuda.Event(enable_timing=True)
            rec['wait1'].record()
        self._handles = []

    def attach_grads(self):
        """point every parameter's .grad at its arena view.  The arena is the source of truth: the wgrad kernels wrote
        into it and the all-reduce ran on it (autograd may have kept a private copy of the pre-reduction values; a
        parameter autograd never touched has a zeroed region).  Requires zero_grad(set_to_none=True) before backward,
        as the reference does (nnUNetTrainer.py:901)."""
        for p in self.params:
            p.grad = self._views[id(p)]
