"""torch.autograd Functions over the C ABI of libmvdseg.so.

PyTorch is plumbing here (device memory, streams, the autograd tape); every device computation below is a call into
the hand-written sm_100a library.  Activations travel as bf16 tensors of logical shape [B, D, H, W, C] ("NDHWC"),
possibly a channel slice of a wider buffer (voxel pitch ld > C), see include/mvdseg.h.
"""
import ctypes
import os
import weakref
import numpy as np
from typing import Optional, Sequence, Tuple

import torch

from ._lib import NormBwdStatsArgs, ConvArgs, DiceCESegment, MvdError, lib

BF16 = torch.bfloat16


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def require_cuda(t: torch.Tensor, who: str):
    if not t.is_cuda:
        raise MvdError(f'{who}: tensors must live on a CUDA device (got {t.device}); libmvdseg has no CPU path')


def cl_pitch(t: torch.Tensor) -> int:
    """voxel pitch (elements) of an NDHWC tensor; raises if the voxels are not dense."""
    assert t.dim() == 5, t.shape
    B, D, H, W, C = t.shape
    st = t.stride()
    if C > 1 and st[4] != 1:
        raise MvdError('NDHWC tensor must have unit channel stride')
    if W > 1:
        ld = st[3]
    elif H > 1:
        ld = st[2]
    elif D > 1:
        ld = st[1]
    elif B > 1:
        ld = st[0]
    else:
        ld = C
    exp = (D * H * W * ld, H * W * ld, W * ld, ld)
    for i, n in enumerate((B, D, H, W)):
        if n > 1 and st[i] != exp[i]:
            raise MvdError(f'NDHWC tensor has non-dense voxels: shape {tuple(t.shape)} strides {st}')
    if ld < C:
        raise MvdError('bad pitch')
    return ld


def cl_pitch_ok(t: torch.Tensor) -> bool:
    """is t a pitched NDHWC tensor the kernels can address in place (16-byte aligned rows)?"""
    try:
        return t.dim() == 5 and cl_pitch(t) % 8 == 0 and t.data_ptr() % 16 == 0
    except MvdError:
        return False


def as_cl(t: torch.Tensor) -> torch.Tensor:
    """returns t if it is a valid pitched NDHWC bf16 tensor, else a dense copy."""
    if t.dtype != BF16:
        t = t.to(BF16)
    try:
        cl_pitch(t)
        return t
    except MvdError:
        return t.contiguous()


def ncdhw_view(t_cl: torch.Tensor) -> torch.Tensor:
    """[B,D,H,W,C] -> logical [B,C,D,H,W] view (channels-last-3d memory format)."""
    return t_cl.permute(0, 4, 1, 2, 3)


def to_cl_view(t: torch.Tensor) -> torch.Tensor:
    """logical [B,C,D,H,W] (any strides) -> NDHWC bf16 tensor usable by the kernels (view when possible)."""
    return as_cl(t.permute(0, 2, 3, 4, 1))


# ---------------------------------------------------------------------------------------------------------------
# layout at the module edge
# ---------------------------------------------------------------------------------------------------------------
def input_to_cl(x: torch.Tensor) -> torch.Tensor:
    """fp32 NCDHW batch -> bf16 NDHWC (no gradient: the network input is data).  A channel slice of a wider batch
    (data[:, 0:1]: dense [C][D][H][W] per sample, larger batch stride) is read in place."""
    require_cuda(x, 'input_to_cl')
    if x.dtype != torch.float32:
        x = x.float()
    B, C, D, H, W = x.shape
    V = D * H * W
    st = x.stride()
    per_sample_dense = st[4] == 1 and st[3] == W and st[2] == H * W and (C == 1 or st[1] == V)
    if not per_sample_dense or (B > 1 and st[0] < C * V):
        x = x.contiguous()
        st = x.stride()
    out = torch.empty((B, D, H, W, C), dtype=BF16, device=x.device)
    lib.ncdhw_f32_to_ndhwc_bf16(x.data_ptr(), st[0] if B > 1 else 0, out.data_ptr(), B, C, V, C, _stream())
    return out


def cl_to_ncdhw_f32(t_cl: torch.Tensor) -> torch.Tensor:
    B, D, H, W, C = t_cl.shape
    ld = cl_pitch(t_cl)
    out = torch.empty((B, C, D, H, W), dtype=torch.float32, device=t_cl.device)
    lib.ndhwc_bf16_to_ncdhw_f32(t_cl.data_ptr(), ld, out.data_ptr(), B, C, D * H * W, _stream())
    return out


# ---------------------------------------------------------------------------------------------------------------
# convolution plumbing
# ---------------------------------------------------------------------------------------------------------------
class Slot:
    """holds a preallocated output tensor; passed to Functions as a non-tensor argument so that autograd does not
    treat the destination buffer as an input."""
    __slots__ = ('t',)

    def __init__(self, t):
        self.t = t


class ConvGeom:
    """geometry of a conv (or of the conv a transposed conv is the adjoint of)."""
    __slots__ = ('k', 's', 'p')

    def __init__(self, k, s, p):
        self.k, self.s, self.p = tuple(k), tuple(s), tuple(p)

    def out_size(self, in_size):
        return tuple((i + 2 * p - k) // s + 1 for i, k, s, p in zip(in_size, self.k, self.s, self.p))


_stem_mode = 'fused'   # 'im2col' selects the explicit X_col form (csrc/stem.cu); test / A-B hook


def set_stem_mode(mode: str):
    global _stem_mode
    assert mode in ('fused', 'im2col')
    _stem_mode = mode


_ALGO = {'auto': 0, 'generic': 1, 'tc': 2}
_default_algo = 0


def set_conv_algo(name: str):
    """'auto' (tcgen05 where covered), 'generic' (CUDA-core tiles), 'tc' (tcgen05 or error). Test hook."""
    global _default_algo
    _default_algo = _ALGO[name]


def _conv_args(geom: ConvGeom, x_cl: torch.Tensor, y_cl: torch.Tensor, w_packed=None, bias=None, stats=None, dw=None,
               dbias=None, accumulate=False, workspace=None) -> ConvArgs:
    B, Di, Hi, Wi, Cin = x_cl.shape
    Bo, Do, Ho, Wo, Cout = y_cl.shape
    assert B == Bo
    a = ConvArgs()
    a.B, a.Di, a.Hi, a.Wi, a.Cin = B, Di, Hi, Wi, Cin
    a.Do, a.Ho, a.Wo, a.Cout = Do, Ho, Wo, Cout
    a.kd, a.kh, a.kw = geom.k
    a.sd, a.sh, a.sw = geom.s
    a.pd, a.ph, a.pw = geom.p
    a.x, a.ldx = x_cl.data_ptr(), cl_pitch(x_cl)
    a.y, a.ldy = y_cl.data_ptr(), cl_pitch(y_cl)
    a.w = _ptr(w_packed)
    a.bias = _ptr(bias)
    a.stats = _ptr(stats)
    a.dw = _ptr(dw)
    a.dbias = _ptr(dbias)
    a.workspace = _ptr(workspace)
    a.workspace_bytes = 0 if workspace is None else workspace.numel() * workspace.element_size()
    a.algo = _default_algo
    a.accumulate = 1 if accumulate else 0
    return a


class WeightPacker:
    """bf16 GEMM layouts of a whole set of conv weights, refreshed by ONE kernel launch (mvd_pack_conv_weights_multi).

    entries: iterable of (weight fp32 [Cout][Cin][kd][kh][kw] (or [CinT][CoutT][k..] for a transposed conv: same
    indexing), want_fprop, want_dgrad).  The packed buffers and the device-side descriptor table are persistent
    (CUDA-graph friendly); ``run()`` re-reads the current weight values."""

    _DESC = np.dtype([('w', '<u8'), ('wf', '<u8'), ('wd', '<u8'), ('Cout', '<i4'), ('Cin', '<i4'), ('taps', '<i4'),
                      ('block_begin', '<i4'), ('stem_kpad', '<i4'), ('reserved', '<i4')])

    def __init__(self, entries):
        entries = list(entries)
        assert entries
        self.packed = {}
        rows = np.zeros(len(entries), dtype=self._DESC)
        blocks = 0
        self._keep = []
        for i, ent in enumerate(entries):
            w, want_f, want_d = ent[:3]
            stem_kpad = int(ent[3]) if len(ent) > 3 else 0
            require_cuda(w, 'WeightPacker')
            wd_ = w.detach()
            if wd_.dtype != torch.float32 or not wd_.is_contiguous():
                raise MvdError('WeightPacker: weights must be contiguous fp32')
            Cout, Cin = wd_.shape[0], wd_.shape[1]
            taps = int(np.prod(wd_.shape[2:]))
            if taps > 27:
                raise MvdError('WeightPacker: at most 27 taps')
            if stem_kpad:      # the stem's [1][Cout][kpad] im2col-column layout (mvd_stem_conv_fprop), no dgrad layout
                if Cin > 16 or taps * Cin > stem_kpad:
                    raise MvdError('WeightPacker: stem layout needs Cin <= 16 and taps * Cin <= kpad')
                wf, wdg = torch.empty((1, Cout, stem_kpad), dtype=BF16, device=w.device), None
            else:
                wf = torch.empty((taps, Cout, Cin), dtype=BF16, device=w.device) if want_f else None
                wdg = torch.empty((taps, Cin, Cout), dtype=BF16, device=w.device) if want_d else None
            rows[i] = (wd_.data_ptr(), _ptr(wf) or 0, _ptr(wdg) or 0, Cout, Cin, taps, blocks, stem_kpad, 0)
            blocks += lib.pack_blocks(Cout, Cin)
            self.packed[(id(w), stem_kpad)] = (wf, wdg)
            self._keep.append(w)
        self.n, self.blocks = len(entries), blocks
        self.rows = rows          # host copy of the table (the fused optimiser kernel extends it, optim.SGDNesterovClip)
        self.table = torch.from_numpy(rows.view(np.uint8).copy()).to(entries[0][0].device)
        self._fresh_versions = None

    def run(self):
        lib.pack_conv_weights_multi(self.table.data_ptr(), self.n, self.blocks, _stream())
        self.mark_fresh()

    # The packed layouts stay valid across steps when the optimiser refreshes them itself (mvd_sgd_pack_conv_weights):
    # it calls mark_fresh() after its update.  Any other in-place change of a weight (load_state_dict, broadcast, a torch
    # optimiser) bumps the tensor's version counter and is noticed here; raw-pointer updates that do NOT refresh the
    # layouts must call invalidate().
    def mark_fresh(self):
        self._fresh_versions = [w._version for w in self._keep]

    def invalidate(self):
        self._fresh_versions = None

    def is_fresh(self) -> bool:
        fv = self._fresh_versions
        return fv is not None and all(w._version == v for w, v in zip(self._keep, fv))

    def get(self, w, stem_kpad=0):
        return self.packed[(id(w), stem_kpad)]


_single_packers = {}
_active_packer: Optional[WeightPacker] = None   # set by PlainConvUNet.forward for the duration of one forward pass


def set_active_packer(pk: Optional[WeightPacker]):
    global _active_packer
    _active_packer = pk


def _packed_for(weight, want_fprop=True, want_dgrad=True, stem_kpad=0):
    """the bf16 layouts of `weight`: from the network-wide packer when a forward pass armed one (already refreshed by
    its single launch), else packed here."""
    if _active_packer is not None:
        hit = _active_packer.packed.get((id(weight), int(stem_kpad)))
        if hit is not None:
            return hit
    return pack_weights(weight, want_fprop, want_dgrad, stem_kpad)


def pack_weights(w: torch.Tensor, want_fprop=True, want_dgrad=True, stem_kpad=0) \
        -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """fp32 [Cout][Cin][kd][kh][kw] -> bf16 [tap][Cout][Cin] and [tap][Cin][Cout] for ONE layer (stem_kpad > 0: the
    stem's [1][Cout][kpad] layout instead).  The packed buffers are cached per (storage address, shape): repeated calls
    refresh and return the same tensors."""
    require_cuda(w, 'pack_weights')
    w = w.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    key = (w.data_ptr(), tuple(w.shape), bool(want_fprop), bool(want_dgrad), int(stem_kpad), w.device)
    pk = _single_packers.get(key)
    if pk is None:
        if len(_single_packers) > 512:
            _single_packers.clear()
        pk = _single_packers[key] = (WeightPacker([(w, want_fprop, want_dgrad, stem_kpad)]), w)
    pk[0].run()
    return next(iter(pk[0].packed.values()))


def stem_kpad_for(weight: torch.Tensor, stride) -> int:
    """> 0 when a conv with this weight is run by the stem kernels (csrc/stem_tc.cu: 1 or 2 input modalities -> 32
    features, 3x3x3, stride 1); the value is the zero-padded im2col width its packed weight uses."""
    Cout, Cin = weight.shape[0], weight.shape[1]
    if Cin in (1, 2) and Cout == 32 and tuple(weight.shape[2:]) == (3, 3, 3) and tuple(stride) == (1, 1, 1):
        return 32 * Cin
    return 0


class ConvTimer:
    """optional CUDA-event bracket around every conv launch (bench.py's live roofline measurement): per pass
    ('fprop' | 'dgrad' | 'wgrad') the algorithmic FLOPs 2*B*Vout*Cout*Cin*taps and the event-timed duration."""

    def __init__(self, lead_cycles: int = 300000):
        self.records = []       # (pass, flops, start_event, end_event, tag)
        self.mem_records = []   # (kernel family, algorithmic bytes, start_event, end_event): the HBM-bound launches
        # every bracket is preceded by a ~0.1 ms busy-wait kernel on the same stream: the host has then enqueued
        # start event, launches and end event before the GPU gets there, so the interval is device execution time only
        # (without it the eager pass is launch-bound and each bracket also contains the GPU's wait for the ctypes call:
        # the 32->32 128^3 wgrad read 0.55 ms in the step against 0.28 ms in ncu)
        self.lead_cycles = int(lead_cycles)

    def lead(self):
        if self.lead_cycles > 0:
            lib.spin(self.lead_cycles, _stream())

    def mem_summary(self):
        """per HBM-bound kernel family: algorithmic bytes, event-timed ms, launches."""
        torch.cuda.synchronize()
        out = {}
        for name, nbytes, e0, e1 in self.mem_records:
            d = out.setdefault(name, dict(bytes=0.0, ms=0.0, launches=0))
            d['bytes'] += nbytes
            d['ms'] += e0.elapsed_time(e1)
            d['launches'] += 1
        return out

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for kind, flops, e0, e1, tag in self.records:
            d = out.setdefault(kind, dict(flops=0.0, ms=0.0, launches=0))
            d['flops'] += flops
            d['ms'] += e0.elapsed_time(e1)
            d['launches'] += 1
        return out

    def per_layer(self):
        torch.cuda.synchronize()
        out = {}
        for kind, flops, e0, e1, tag in self.records:
            d = out.setdefault((kind, tag), dict(flops=0.0, ms=0.0, launches=0))
            d['flops'] += flops
            d['ms'] += e0.elapsed_time(e1)
            d['launches'] += 1
        return out


_conv_timer: Optional[ConvTimer] = None


def set_conv_timer(t: Optional[ConvTimer]):
    global _conv_timer
    _conv_timer = t


def _timed(kind, a: ConvArgs, fn):
    if _conv_timer is None:
        fn(ctypes.byref(a), _stream())
        return
    flops = 2.0 * a.B * a.Do * a.Ho * a.Wo * a.Cout * a.Cin * a.kd * a.kh * a.kw
    tag = f'{a.Cin}->{a.Cout} k{a.kd}{a.kh}{a.kw} s{a.sd}{a.sh}{a.sw} out{a.Do}x{a.Ho}x{a.Wo}'
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _conv_timer.lead()
    e0.record()
    fn(ctypes.byref(a), _stream())
    e1.record()
    _conv_timer.records.append((kind, flops, e0, e1, tag))


def _timed_call(kind, flops, tag, fn):
    """same bracket for conv launches that do not go through mvd_conv3d_* (the fused stem)."""
    if _conv_timer is None:
        fn()
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _conv_timer.lead()
    e0.record()
    fn()
    e1.record()
    _conv_timer.records.append((kind, flops, e0, e1, tag))


def _timed_mem(name: str, nbytes: float, fn, *a):
    """CUDA-event bracket around one HBM-bound launch when bench.py's timer is armed (roofline_hbm)."""
    if _conv_timer is None:
        return fn(*a)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _conv_timer.lead()
    e0.record()
    r = fn(*a)
    e1.record()
    _conv_timer.mem_records.append((name, float(nbytes), e0, e1))
    return r


def _maybe_workspace(a: ConvArgs, which: int, dev):
    """fprop / dgrad of layers with a small produced lattice run split-K when given scratch for the fp32 partials."""
    nbytes = lib.conv3d_workspace_bytes(ctypes.byref(a), which)
    if not nbytes:
        return None
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    a.workspace, a.workspace_bytes = ws.data_ptr(), nbytes
    return ws


def conv_fprop(geom, x_cl, y_cl, wf, bias=None, stats=None, accumulate=False):
    a = _conv_args(geom, x_cl, y_cl, w_packed=wf, bias=bias, stats=stats, accumulate=accumulate)
    ws = _maybe_workspace(a, 0, x_cl.device)
    _timed('fprop', a, lib.conv3d_fprop)
    del ws


def _norm_bwd_args(nb, bstats) -> NormBwdStatsArgs:
    """nb = (y, stats, gamma, beta, eps, slope) of the ConvNormAct block whose output the conv consumed."""
    y, stats, gamma, beta, eps, slope = nb
    n = NormBwdStatsArgs()
    n.y, n.ldy = y.data_ptr(), cl_pitch(y)
    n.stats, n.gamma, n.beta = stats.data_ptr(), _ptr(gamma), _ptr(beta)
    n.eps, n.slope = eps, slope
    n.bstats = bstats.data_ptr()
    return n


def conv_dgrad_fuses_norm_bwd(geom, x_cl_out, y_cl, wd, nb) -> bool:
    """Would conv_dgrad(..., norm_bwd=nb) form the statistics in its epilogue (no extra pass)?"""
    a = _conv_args(geom, x_cl_out, y_cl, w_packed=wd)
    ws = _maybe_workspace(a, 1, y_cl.device)
    n = _norm_bwd_args(nb, nb[1])      # the output pointer only has to be non-null for the query
    a.norm_bwd = ctypes.addressof(n)
    ok = bool(lib.conv3d_dgrad_fuses_norm_bwd(ctypes.byref(a)))
    del ws
    return ok


def conv_dgrad(geom, x_cl_out, y_cl, wd, bias=None, accumulate=False, stats=None, norm_bwd=None, bstats=None):
    """norm_bwd / bstats: also leave the backward statistics of the InstanceNorm + LeakyReLU in front of x_cl_out in
    ``bstats`` ([B][Cin][2] fp64, zeroed by the caller) -- mvd_conv3d_args.norm_bwd."""
    a = _conv_args(geom, x_cl_out, y_cl, w_packed=wd, bias=bias, accumulate=accumulate, stats=stats)
    ws = _maybe_workspace(a, 1, y_cl.device)
    n = None
    if norm_bwd is not None:
        n = _norm_bwd_args(norm_bwd, bstats)
        a.norm_bwd = ctypes.addressof(n)
    _timed('dgrad', a, lib.conv3d_dgrad)
    del ws, n


def conv_wgrad(geom, x_cl, y_cl, dw, dbias=None):
    a = _conv_args(geom, x_cl, y_cl, dw=dw, dbias=dbias)
    nbytes = lib.conv3d_workspace_bytes(ctypes.byref(a), 2)
    ws = None
    if nbytes:
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=x_cl.device)
        a.workspace, a.workspace_bytes = ws.data_ptr(), nbytes
    _timed('wgrad', a, lib.conv3d_wgrad)


_grad_alloc = None


def set_grad_allocator(fn):
    """fn(param) -> fp32 tensor view (same shape) the weight gradient is written into (DDP bucket arena), or None."""
    global _grad_alloc
    _grad_alloc = fn
    _arena_ids.clear()
    _prezeroed_ids.clear()


_arena_ids = set()

# ---- per-step pool of zero-initialised scratch (InstanceNorm sums, loss accumulators): ONE memset per step instead of
# ---- ~80 tiny fill kernels.  Armed by begin_step() (the trainer calls it at the top of every step); without it, or
# ---- when the pool runs out, zeros() falls back to torch.zeros.
_ZERO_POOL_BYTES = 2 << 20
_zero_pool = {}   # device -> [uint8 tensor, offset]


def begin_step(device):
    ent = _zero_pool.get(device)
    if ent is None:
        ent = _zero_pool[device] = [torch.empty((_ZERO_POOL_BYTES,), dtype=torch.uint8, device=device), 0]
    lib.zero_bytes(ent[0].data_ptr(), _ZERO_POOL_BYTES, torch.cuda.current_stream(device).cuda_stream)   # a memset node
    ent[1] = 0


def zeros(shape, dtype, device):
    ent = _zero_pool.get(device)
    n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
    if ent is None or ent[1] + n > _ZERO_POOL_BYTES:
        return torch.zeros(shape, dtype=dtype, device=device)
    off = ent[1]
    ent[1] = (off + n + 255) // 256 * 256
    return ent[0][off:off + n].view(dtype).view(shape)


_prezeroed_ids = set()   # arena views the trainer clears at the top of every step (GradArena.prezero): no fill here


def set_prezeroed(views):
    _prezeroed_ids.clear()
    _prezeroed_ids.update(id(v) for v in views)


def _grad_like(p: torch.Tensor, zero: bool = False) -> torch.Tensor:
    """destination of a parameter gradient; ``zero``: the kernel accumulates into it (bias / head sums)."""
    if _grad_alloc is not None:
        g = _grad_alloc(p)
        if g is not None:
            _arena_ids.add(id(g))   # the arena's views are persistent objects
            if zero and id(g) not in _prezeroed_ids:
                g.zero_()
            return g
    return (torch.zeros if zero else torch.empty)(p.shape, dtype=torch.float32, device=p.device)


def _ret(g: Optional[torch.Tensor], p: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """what a Function's backward hands to autograd for a parameter gradient.  When the kernel wrote it into the
    gradient arena, the arena view is attached as ``p.grad`` right here and autograd gets nothing: handing the view
    over would make AccumulateGrad clone it (108 device-to-device copies per step).  Consequence: with an arena a
    second backward without zero_grad overwrites instead of accumulating (the trainer never does that,
    nnUNetTrainer.py:901)."""
    if g is None or id(g) not in _arena_ids:
        return g
    if p is not None and p.grad is not g:
        p.grad = g
    return None


_backward_hooks = []


def add_param_grad_ready_hook(fn):
    """fn(list_of_params) is called from backward as soon as those parameters' gradients have been written."""
    _backward_hooks.append(fn)
    return fn


def clear_param_grad_ready_hooks():
    _backward_hooks.clear()


def _notify(params):
    for fn in _backward_hooks:
        fn(params)


# ---------------------------------------------------------------------------------------------------------------
# Skip-connection gradient: an encoder stage output feeds the next encoder stage AND the decoder's concat buffer, so
# autograd would add the two gradients with a separate elementwise kernel (read 2, write 1 full-resolution tensors).
# Instead the decoder-side gradient (a channel slice of the concat buffer's gradient, produced first) is parked here
# and the next encoder stage's dgrad ACCUMULATES into it in its epilogue; ConcatViewFn then reports no gradient of its
# own for the skip.  Only armed when the forward pass saw exactly one dgrad consumer of that tensor.
# ---------------------------------------------------------------------------------------------------------------
_dx_consumers = {}     # data_ptr of a conv input that needs a gradient -> number of conv consumers (this forward)
_pending_skip = {}     # data_ptr -> decoder-side gradient view, between ConcatViewFn.backward and the consumer's dgrad


def reset_skip_registry():
    _dx_consumers.clear()
    _pending_skip.clear()
    _head_inputs.clear()
    _concat_bufs.clear()
    _dx_colsum.clear()
    _norm_outputs.clear()
    _dx_bstats.clear()
    _pair_pending.clear()      # entries whose head never ran backward (zero-weighted scale) must not pile up


# ---------------------------------------------------------------------------------------------------------------
# Decoder stage outputs have TWO consumers: the stage's segmentation head and the next up-convolution
# (UNetDecoder.py:104-121), so autograd would sum the two gradients with an elementwise kernel.  In backward the
# up-convolution runs first (it was created later); it hands its gradient to autograd as usual and leaves a reference
# here, and the head's backward then ADDS its share into that very tensor inside its own kernel (mvd_head_bwd
# accumulate_dz) and reports no gradient of its own.  If the head runs first, or never (zero-weighted scale), nothing
# is parked and autograd's own accumulation applies: correct in every order, fused in the usual one.
# ---------------------------------------------------------------------------------------------------------------
_concat_bufs = set()   # data_ptr of the decoder's concat buffers (this forward): their gradient's first half is the
                       # gradient of an up-convolution's output, whose bias gradient is its per-channel sum
_dx_colsum = {}        # data_ptr of such a gradient tensor -> (statistics buffer its dgrad epilogue filled, channels)
_colsum_fusion = True
# A ConvNormAct block whose output feeds exactly one 3x3x3 / stride-1 conv (the first conv of every stage): that conv's
# data-gradient epilogue already holds each gradient value it writes, so it also forms the block's InstanceNorm-backward
# sums there (mvd_conv3d_args.norm_bwd) and the block skips its statistics pass over the gradient.
_norm_outputs = {}     # data_ptr of a block's activation -> (weakref to it, y, stats, gamma, beta, eps, slope) of the block
_dx_bstats = {}        # data_ptr of a produced gradient -> (bstats it comes with, data_ptr of the block's raw conv output)
_norm_bwd_fusion = os.environ.get('MVD_NO_NORM_BWD_FUSION', '0') != '1'


import os as _os
_head_fusion = False


def set_head_fusion(on: bool):
    """fold InstanceNorm + LeakyReLU of the last decoder block into its segmentation head (csrc/norm_head.cu).  OFF by
    default: measured on the B200 at 2 x 128^3 the fused kernels move 40 % fewer bytes but are issue-bound (the per-voxel
    recomputation of the head's data gradient and of the activation adds ~12 % instructions to kernels that already
    keep the issue slots 74 % busy): forward 0.120 ms vs 0.159 ms unfused, backward 0.407 ms vs 0.362 ms -- no net gain
    (profiles/r2_norm_head_fusion.txt).  MVD_HEAD_FUSION=1 or this switch turns it on."""
    global _head_fusion
    _head_fusion = bool(on)


_head_fusion = _os.environ.get('MVD_HEAD_FUSION', '0') == '1'


def head_fusion_ok(channels: int, head) -> bool:
    return (_head_fusion and head is not None and _default_algo != 1 and head.weight.is_cuda
            and bool(lib.inorm_lrelu_head_supported(int(channels), int(head.weight.shape[0]))))


def set_norm_bwd_fusion(on: bool):
    """InstanceNorm-backward sums out of the consumer conv's data-gradient epilogue (default on)."""
    global _norm_bwd_fusion
    _norm_bwd_fusion = bool(on)


def set_colsum_fusion(on: bool):
    """A/B and test hook: take the up-convolution's bias gradient from the consumer's dgrad epilogue sums (default) or
    from a streaming pass over the gradient tensor."""
    global _colsum_fusion
    _colsum_fusion = bool(on)

_head_inputs = {}      # data_ptr of a head's input (this forward) -> the HeadFn ctx
_pair_pending = {}     # pair key -> gradient tensor the up-convolution's backward already returned
_pair_seq = [0]


# ---------------------------------------------------------------------------------------------------------------
# Backward overlap.  The weight gradient of a conv is needed by nobody until the optimiser (or the bucket all-reduce),
# so it is DEFERRED onto a side stream: the main stream runs the dependency chain
#     IN-backward(L) -> dgrad(L) -> IN-backward(L-1) -> dgrad(L-1) -> ...
# while the side stream runs wgrad(L), wgrad(L-1), ... one after the other, each forked after its layer's dgrad.  The
# tensor-pipe-bound wgrad kernels (one 190-thread CTA per SM, ~10 K registers) leave room on every SM for the HBM-bound
# InstanceNorm blocks of the NEXT layer, so the two run side by side; dgrad and wgrad (both one big CTA per SM) simply
# time-share.  Buffers a deferred wgrad reads stay referenced in _pending until the main stream has waited for it; the
# whole queue is joined by an autograd end-of-backward callback (so .grad is complete when backward() returns, also
# inside CUDA graph capture).  Off while bench.py's per-kernel conv timer is armed.
# ---------------------------------------------------------------------------------------------------------------
import os as _os
_bwd_overlap = _os.environ.get('MVD_NO_BWD_OVERLAP', '0') != '1'
_side_streams = {}
_pending = []            # (completion event on the side stream, main stream, tensors kept alive)
_callback_queued = False
_MAX_PENDING = int(_os.environ.get('MVD_MAX_PENDING_WGRAD', '64'))   # deferred wgrads in flight before the main stream waits for the oldest


def set_backward_overlap(on: bool):
    global _bwd_overlap
    _bwd_overlap = bool(on)


def side_stream(dev, create: bool = False):
    """the stream deferred weight gradients run on (None if none was ever forked on this device)."""
    key = torch.device(dev)
    if key.index is None:
        key = torch.device('cuda', torch.cuda.current_device())
    st = _side_streams.get(key)
    if st is None and create:
        st = _side_streams[key] = torch.cuda.Stream(device=key, priority=0)    # lowest: the weight gradients fill gaps
    return st


def _flush_pending(keep: int = 0):
    while len(_pending) > keep:
        ev, main, _refs = _pending.pop(0)
        main.wait_event(ev)


def join_pending():
    """make the compute stream wait for every deferred weight gradient (end of backward)."""
    global _callback_queued
    _callback_queued = False
    _flush_pending(0)


def _defer_wgrad(dev, launch, keepalive, dw=None):
    """run `launch()` (the wgrad launches) on the side stream after everything enqueued so far on the current stream;
    returns False (nothing launched) when overlap is off.

    Only gradients that live in the trainer's arena are deferred: those bypass autograd (``_ret`` attaches them as
    ``p.grad`` itself).  A free-standing ``dw`` is handed to autograd's AccumulateGrad, which -- because the keep-alive
    list below still references the tensor -- cannot steal it and CLONES it on the compute stream, i.e. possibly before
    the side stream has written it (found by the full-size block parity test: garbage conv.weight gradients at 128^3,
    where the wgrad kernels run long enough to lose that race)."""
    global _callback_queued
    if not (_bwd_overlap and _conv_timer is None) or dw is None or id(dw) not in _arena_ids:
        return False
    main = torch.cuda.current_stream(dev)
    side = side_stream(dev, create=True)
    ev = torch.cuda.Event()
    ev.record(main)
    with torch.cuda.stream(side):
        side.wait_event(ev)
        launch()
        done = torch.cuda.Event()
        done.record(side)
    _pending.append((done, main, keepalive))
    if not _callback_queued:
        _callback_queued = True
        try:
            torch.autograd.Variable._execution_engine.queue_callback(join_pending)
        except RuntimeError:      # not inside a backward pass (a Function's backward called by hand): join right away
            join_pending()
            return True
    _flush_pending(_MAX_PENDING)
    return True


# ---------------------------------------------------------------------------------------------------------------
# Conv3d -> InstanceNorm3d(affine) -> LeakyReLU   (one ConvDropoutNormReLU block)
# ---------------------------------------------------------------------------------------------------------------
class ConvNormActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_cl, weight, bias, gamma, beta, geom: ConvGeom, eps: float, slope: float,
                out_slot: Optional['Slot'], params_for_hook, head_w=None, head_b=None, private_input=False):
        """head_w / head_b: the 1x1x1 segmentation head that is the ONLY consumer of this block's output (last decoder
        stage): the function then returns the head's logits and the normalised activation never exists in HBM
        (csrc/norm_head.cu).
        private_input: x_cl is the output of another ConvNormAct block and this conv is its ONLY consumer (conv i > 0 of a
        StackedConvBlocks): the data gradient this conv produces is then that block's complete incoming gradient, and
        its InstanceNorm-backward sums are formed in our data-gradient epilogue (see _norm_outputs)."""
        require_cuda(x_cl, 'ConvNormAct')
        out = out_slot.t if out_slot is not None else None
        x_cl = as_cl(x_cl)
        B, Di, Hi, Wi, Cin = x_cl.shape
        Cout = weight.shape[0]
        Do, Ho, Wo = geom.out_size((Di, Hi, Wi))
        dev = x_cl.device
        need_dx = ctx.needs_input_grad[0]
        if need_dx:
            _dx_consumers[x_cl.data_ptr()] = _dx_consumers.get(x_cl.data_ptr(), 0) + 1
        ctx.x_ptr = x_cl.data_ptr()
        ctx.want_colsum = _colsum_fusion and need_dx and x_cl.data_ptr() in _concat_bufs and Cin in (32, 64) and geom.k == (3, 3, 3) \
            and geom.s == (1, 1, 1) and geom.p == (1, 1, 1)
        y = torch.empty((B, Do, Ho, Wo, Cout), dtype=BF16, device=dev)
        stats = zeros((B, Cout, 2), torch.float64, dev)
        V = Do * Ho * Wo
        # stem (1-2 input modalities): tensor-core GEMM over an im2col tile built in shared memory (csrc/stem_tc.cu);
        # other thin stems: explicit im2col + single-tap tensor-core GEMM (csrc/stem.cu)
        stem = (Cin <= 4 and not need_dx and geom.s == (1, 1, 1) and Cout % 32 == 0 and _default_algo != 1)
        fused_stem = (stem and Cin in (1, 2) and Cout == 32 and geom.k == (3, 3, 3) and geom.p == (1, 1, 1)
                      and cl_pitch(x_cl) == Cin and x_cl.data_ptr() % 4 == 0 and _stem_mode != 'im2col')
        ctx.stem, ctx.fused_stem = stem, fused_stem
        if stem:
            taps = geom.k[0] * geom.k[1] * geom.k[2]
            kpad = 32 * Cin if fused_stem else (32 if taps * Cin <= 32 else 64)
            assert taps * Cin <= kpad
            # [Cout][Cin][taps] -> [1 tap][Cout][(tap, ci) zero-padded]: written by the weight-pack launch
            wcol, _ = _packed_for(weight, True, False, stem_kpad=kpad)
            if fused_stem:
                _timed_call('fprop', 2.0 * B * V * Cout * Cin * taps, f'stem {Cin}->{Cout} k333 out{Do}x{Ho}x{Wo}',
                            lambda: lib.stem_conv_fprop(x_cl.data_ptr(), B, Di, Hi, Wi, Cin, wcol.data_ptr(), _ptr(bias),
                                                        y.data_ptr(), cl_pitch(y), stats.data_ptr(), _stream()))
                wd = None
            else:
                x_col = torch.empty((B, Di, Hi, Wi, kpad), dtype=BF16, device=dev)
                lib.im2col_small(x_cl.data_ptr(), cl_pitch(x_cl), B, Di, Hi, Wi, Cin, *geom.k, *geom.p,
                                 x_col.data_ptr(), kpad, _stream())
                g1 = ConvGeom((1, 1, 1), (1, 1, 1), (0, 0, 0))
                conv_fprop(g1, x_col, y, wcol, bias=bias, stats=stats)
                x_cl, wd, geom = x_col, None, g1
        else:
            wf, wd = _packed_for(weight, True, need_dx)
            conv_fprop(geom, x_cl, y, wf, bias=bias, stats=stats)   # InstanceNorm sums come out of the conv epilogue
        ctx.prev_norm = None
        if _norm_bwd_fusion and private_input and need_dx and not stem:
            ent = _norm_outputs.pop(x_cl.data_ptr(), None)
            if ent is not None and ent[0]() is not None and tuple(ent[1].shape) == tuple(x_cl.shape) and \
                    conv_dgrad_fuses_norm_bwd(geom, x_cl, y, wd, ent[1:]):
                ctx.prev_norm = ent[1:]
        ctx.geom, ctx.eps, ctx.slope = geom, eps, slope
        ctx.params_for_hook = params_for_hook
        ctx.has_head = head_w is not None
        if ctx.has_head:
            K = head_w.shape[0]
            assert not stem and out is None and lib.inorm_lrelu_head_supported(Cout, K)
            hw = head_w.detach().reshape(K, Cout)
            logits = torch.empty((B, Do, Ho, Wo, K), dtype=BF16, device=dev)
            _timed_mem('inorm_lrelu_head_fwd', B * V * (2.0 * Cout + 2.0 * K), lib.inorm_lrelu_head_fwd, y.data_ptr(),
                       cl_pitch(y), stats.data_ptr(), _ptr(gamma), _ptr(beta), hw.data_ptr(), _ptr(head_b),
                       logits.data_ptr(), B, V, Cout, K, eps, slope, _stream())
            ctx.save_for_backward(x_cl, y, stats, wd if need_dx else None, weight, bias, gamma, beta, head_w, head_b)
            return logits
        z = out if out is not None else torch.empty_like(y)
        _timed_mem('inorm_lrelu_fwd', 4.0 * B * V * Cout, lib.inorm_lrelu_fwd, y.data_ptr(), cl_pitch(y), z.data_ptr(),
                   cl_pitch(z), stats.data_ptr(), _ptr(gamma), _ptr(beta), B, V, Cout, eps, slope, _stream())
        ctx.save_for_backward(x_cl, y, stats, wd if need_dx else None, weight, bias, gamma, beta, None, None)
        if _norm_bwd_fusion and out is None:
            _norm_outputs[z.data_ptr()] = (weakref.ref(z), y, stats, gamma, beta, eps, slope)
        return z

    @staticmethod
    def backward(ctx, dz):
        x_cl, y, stats, wd, weight, bias, gamma, beta, head_w, head_b = ctx.saved_tensors
        geom, eps, slope = ctx.geom, ctx.eps, ctx.slope
        B, Do, Ho, Wo, Cout = y.shape
        V = Do * Ho * Wo
        dev = y.device
        st = _stream()
        fused = None if ctx.has_head else _dx_bstats.pop(dz.data_ptr(), None)
        if fused is not None and fused[1] == y.data_ptr() and tuple(dz.shape) == tuple(y.shape):
            bstats = fused[0]          # left by the data-gradient epilogue of the conv that consumed our activation
        else:
            fused = None
            bstats = zeros((B, Cout, 2), torch.float64, dev)
        dy = torch.empty_like(y)
        dgamma = _grad_like(gamma) if gamma is not None and ctx.needs_input_grad[3] else None
        dbeta = _grad_like(beta) if beta is not None and ctx.needs_input_grad[4] else None
        db = _grad_like(bias, zero=True) if bias is not None and ctx.needs_input_grad[2] else None
        dhw = dhb = None
        if ctx.has_head:
            # the incoming gradient is d loss / d logits: the head's data gradient dz = dlogits W is recomputed per voxel
            # inside both InstanceNorm-backward passes, the head's own dW / db come out of the first one
            dl = as_cl(dz)
            K = head_w.shape[0]
            if cl_pitch(dl) != K or dl.data_ptr() % 8:
                dl = dl.contiguous()
            hw = head_w.detach().reshape(K, Cout)
            dhw = _grad_like(head_w, zero=True) if ctx.needs_input_grad[10] else None
            dhb = _grad_like(head_b, zero=True) if head_b is not None and ctx.needs_input_grad[11] else None
            _timed_mem('inorm_lrelu_head_bwd_stats', B * V * (2.0 * Cout + 2.0 * K), lib.inorm_lrelu_head_bwd_stats,
                       dl.data_ptr(), y.data_ptr(), cl_pitch(y), stats.data_ptr(), _ptr(gamma), _ptr(beta), hw.data_ptr(),
                       B, V, Cout, K, eps, slope, bstats.data_ptr(), _ptr(dhw), _ptr(dhb), st)
            _timed_mem('inorm_lrelu_head_bwd_apply', B * V * (4.0 * Cout + 2.0 * K), lib.inorm_lrelu_head_bwd_apply,
                       dl.data_ptr(), y.data_ptr(), cl_pitch(y), dy.data_ptr(), cl_pitch(dy), stats.data_ptr(),
                       bstats.data_ptr(), _ptr(gamma), _ptr(beta), hw.data_ptr(), B, V, Cout, K, eps, slope, _ptr(dgamma),
                       _ptr(dbeta), _ptr(db), st)
        else:
            dz = as_cl(dz)
            if fused is None:
                _timed_mem('inorm_lrelu_bwd_stats', 4.0 * B * V * Cout, lib.inorm_lrelu_bwd_stats, dz.data_ptr(),
                           cl_pitch(dz), y.data_ptr(), cl_pitch(y), stats.data_ptr(), _ptr(gamma), _ptr(beta), B, V, Cout,
                           eps, slope, bstats.data_ptr(), st)
            _timed_mem('inorm_lrelu_bwd_apply', 6.0 * B * V * Cout, lib.inorm_lrelu_bwd_apply, dz.data_ptr(), cl_pitch(dz),
                       y.data_ptr(), cl_pitch(y), dy.data_ptr(), cl_pitch(dy), stats.data_ptr(), bstats.data_ptr(),
                       _ptr(gamma), _ptr(beta), B, V, Cout, eps, slope, _ptr(dgamma), _ptr(dbeta), _ptr(db), st)
        dw = _grad_like(weight) if ctx.needs_input_grad[1] else None
        plain_wgrad = False
        if dw is not None:
            if ctx.stem:
                Cin_w, taps = weight.shape[1], weight.shape[2] * weight.shape[3] * weight.shape[4]
                if ctx.fused_stem:       # writes dw in the torch layout [Cout][Cin][27] itself
                    Bx, Dx, Hx, Wx, _ = x_cl.shape
                    _timed_call('wgrad', 2.0 * Bx * Dx * Hx * Wx * Cout * Cin_w * taps,
                                f'stem {Cin_w}->{Cout} k333 out{Dx}x{Hx}x{Wx}',
                                lambda: lib.stem_conv_wgrad(x_cl.data_ptr(), Bx, Dx, Hx, Wx, Cin_w, dy.data_ptr(),
                                                            cl_pitch(dy), dw.data_ptr(), st))
                else:
                    kpad = x_cl.shape[-1]
                    dw_col = torch.empty((Cout, kpad, 1, 1, 1), dtype=torch.float32, device=dev)
                    conv_wgrad(geom, x_cl, dy, dw_col, None)
                    dw.copy_(dw_col.reshape(Cout, kpad)[:, :taps * Cin_w].reshape(Cout, taps, Cin_w).permute(0, 2, 1)
                             .reshape(weight.shape))
            else:
                plain_wgrad = True
        dx = None
        if ctx.needs_input_grad[0]:
            pend = _pending_skip.pop(ctx.x_ptr, None)
            if pend is not None and tuple(pend.shape) == tuple(x_cl.shape):
                dx = pend     # the decoder's share of this skip tensor's gradient: add ours in the dgrad epilogue
                conv_dgrad(geom, dx, dy, wd, accumulate=True)
            else:
                dx = torch.empty(x_cl.shape, dtype=BF16, device=dev)
                colsum = None
                if ctx.want_colsum:      # channel sums of dx out of the dgrad epilogue: the up-convolution's bias gradient
                    colsum = zeros((x_cl.shape[0], x_cl.shape[-1], 2), torch.float64, dev)
                    _dx_colsum[dx.data_ptr()] = (colsum, x_cl.shape[-1])
                nb = ctx.prev_norm if colsum is None else None
                nb_out = None
                if nb is not None:
                    nb_out = zeros((x_cl.shape[0], x_cl.shape[-1], 2), torch.float64, dev)
                    _dx_bstats[dx.data_ptr()] = (nb_out, nb[0].data_ptr())
                conv_dgrad(geom, dx, dy, wd, stats=colsum, norm_bwd=nb, bstats=nb_out)
        if plain_wgrad:   # after the dgrad: deferred onto the side stream (see "Backward overlap"), else right here
            if not _defer_wgrad(dev, lambda: conv_wgrad(geom, x_cl, dy, dw, None), (x_cl, dy, dw), dw):
                conv_wgrad(geom, x_cl, dy, dw, None)
        if ctx.params_for_hook:
            _notify(ctx.params_for_hook)
        return (dx, _ret(dw, weight), _ret(db, bias), _ret(dgamma, gamma), _ret(dbeta, beta), None, None, None, None, None,
                _ret(dhw, head_w), _ret(dhb, head_b), None)


# ---------------------------------------------------------------------------------------------------------------
# ConvTranspose3d with kernel == stride (the decoder's upsampling), expressed through the adjoint conv
# ---------------------------------------------------------------------------------------------------------------
class ConvTransposeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_cl, weight, bias, stride, out_slot: Optional['Slot'], params_for_hook):
        require_cuda(x_cl, 'ConvTranspose')
        out = out_slot.t if out_slot is not None else None
        x_cl = as_cl(x_cl)
        B, d, h, w, CinT = x_cl.shape
        CoutT = weight.shape[1]
        s = tuple(stride)
        geom = ConvGeom(s, s, (0, 0, 0))  # adjoint conv: hi-res (CoutT) -> lo-res (CinT)
        dev = x_cl.device
        wf, wd = _packed_for(weight, True, True)
        up = out if out is not None else torch.empty((B, d * s[0], h * s[1], w * s[2], CoutT), dtype=BF16, device=dev)
        conv_dgrad(geom, up, x_cl, wd, bias=bias)
        ctx.geom = geom
        ctx.params_for_hook = params_for_hook
        ctx.pair_key = None
        head = _head_inputs.get(x_cl.data_ptr())
        if head is not None and ctx.needs_input_grad[0] and head[1] == tuple(x_cl.shape) and head[0].pair_key is None:
            _pair_seq[0] += 1
            ctx.pair_key = head[0].pair_key = _pair_seq[0]
        ctx.save_for_backward(x_cl, wf, weight, bias)
        return up

    @staticmethod
    def backward(ctx, dup):
        x_cl, wf, weight, bias = ctx.saved_tensors
        geom = ctx.geom
        dup = as_cl(dup)
        dev = dup.device
        dw = _grad_like(weight) if ctx.needs_input_grad[1] else None
        db = None
        if bias is not None and ctx.needs_input_grad[2]:
            db = _grad_like(bias)
            B, D, H, W, C = dup.shape
            ent = _dx_colsum.pop(dup.data_ptr(), None)
            if ent is not None and ent[1] >= C:    # the producer of `dup` left its channel sums: no pass over the tensor
                lib.stats_channel_sum(ent[0].data_ptr(), B, ent[1], 0, C, db.data_ptr(), _stream())
            else:
                lib.channel_sum(dup.data_ptr(), cl_pitch(dup), B * D * H * W, C, db.data_ptr(), _stream())
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(x_cl.shape, dtype=BF16, device=dev)
            conv_fprop(geom, dup, dx, wf)
            if ctx.pair_key is not None:
                _pair_pending[ctx.pair_key] = dx     # the stage's head adds its share in place (see _head_inputs)
        if dw is not None:
            if not _defer_wgrad(dev, lambda: conv_wgrad(geom, dup, x_cl, dw, None), (dup, x_cl, dw), dw):
                conv_wgrad(geom, dup, x_cl, dw, None)
        if ctx.params_for_hook:
            _notify(ctx.params_for_hook)
        return dx, _ret(dw, weight), _ret(db, bias), None, None, None


# ---------------------------------------------------------------------------------------------------------------
# zero-copy concat: `up` and `skip` are the two channel slices of one buffer
# ---------------------------------------------------------------------------------------------------------------
class ConcatViewFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, up, skip, buf_slot: 'Slot'):
        buf = buf_slot.t
        c1 = up.shape[-1]
        assert up.data_ptr() == buf.data_ptr() and skip.data_ptr() == buf.data_ptr() + c1 * buf.element_size()
        assert buf.shape[-1] == c1 + skip.shape[-1]
        ctx.c1 = c1
        ctx.skip_ptr = skip.data_ptr()
        # exactly one conv takes a gradient w.r.t. the skip tensor (the next encoder stage): it will fold our share in
        ctx.defer = ctx.needs_input_grad[1] and _dx_consumers.get(ctx.skip_ptr, 0) == 1
        if ctx.needs_input_grad[0]:
            _concat_bufs.add(buf.data_ptr())
        return buf.view(buf.shape)  # a fresh tensor object aliasing the buffer

    @staticmethod
    def backward(ctx, g):
        if ctx.defer and g.dtype == BF16 and cl_pitch_ok(g[..., ctx.c1:]):
            _pending_skip[ctx.skip_ptr] = g[..., ctx.c1:]
            return g[..., :ctx.c1], None, None
        return g[..., :ctx.c1], g[..., ctx.c1:], None


# ---------------------------------------------------------------------------------------------------------------
# 1x1x1 segmentation head
# ---------------------------------------------------------------------------------------------------------------
class HeadFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z_cl, weight, bias, params_for_hook):
        require_cuda(z_cl, 'Head')
        z_cl = as_cl(z_cl)
        B, D, H, W, C = z_cl.shape
        K = weight.shape[0]
        logits = torch.empty((B, D, H, W, K), dtype=BF16, device=z_cl.device)
        w2 = weight.detach().reshape(K, C)
        if w2.dtype != torch.float32 or not w2.is_contiguous():
            w2 = w2.float().contiguous()
        _timed_mem('head_fwd', 2.0 * B * D * H * W * (C + K), lib.head_fwd, z_cl.data_ptr(), cl_pitch(z_cl),
                   w2.data_ptr(), _ptr(bias), logits.data_ptr(), K, B * D * H * W, C, K, _stream())
        ctx.params_for_hook = params_for_hook
        ctx.pair_key = None
        if ctx.needs_input_grad[0]:
            _head_inputs[z_cl.data_ptr()] = (ctx, tuple(z_cl.shape))
        ctx.save_for_backward(z_cl, weight, bias)
        return logits

    @staticmethod
    def backward(ctx, dl):
        z_cl, weight, bias = ctx.saved_tensors
        dl = as_cl(dl)
        B, D, H, W, C = z_cl.shape
        K = weight.shape[0]
        dev = z_cl.device
        w2 = weight.detach().reshape(K, C)
        if w2.dtype != torch.float32 or not w2.is_contiguous():
            w2 = w2.float().contiguous()
        parked = _pair_pending.pop(ctx.pair_key, None) if ctx.pair_key is not None else None
        acc_dz = parked is not None and ctx.needs_input_grad[0] and tuple(parked.shape) == (B, D, H, W, C) \
            and cl_pitch_ok(parked)
        if acc_dz:
            dz = parked        # the up-convolution's share, already in autograd's hands: add ours in the kernel
        else:
            dz = torch.empty((B, D, H, W, C), dtype=BF16, device=dev) if ctx.needs_input_grad[0] else None
        dw = db = None
        if ctx.needs_input_grad[1]:
            dw = _grad_like(weight, zero=True)
            db = _grad_like(bias, zero=True) if bias is not None else None
        _timed_mem('head_bwd', 2.0 * B * D * H * W * (K + C + (C if dz is not None else 0) + (C if acc_dz else 0)),
                   lib.head_bwd, dl.data_ptr(), cl_pitch(dl), z_cl.data_ptr(), cl_pitch(z_cl), w2.data_ptr(), _ptr(dz),
                   cl_pitch(dz) if dz is not None else 0, _ptr(dw), _ptr(db), B * D * H * W, C, K, int(acc_dz), _stream())
        if acc_dz:
            dz = None          # reported through the parked tensor
        if ctx.params_for_hook:
            _notify(ctx.params_for_hook)
        return dz, _ret(dw, weight), _ret(db, bias), None


# ---------------------------------------------------------------------------------------------------------------
# losses
# ---------------------------------------------------------------------------------------------------------------
def _target_f32(t: torch.Tensor) -> torch.Tensor:
    """(B,1,D,H,W) or (B,D,H,W) class ids -> contiguous fp32 [B][V]."""
    if t.dim() == 5:
        assert t.shape[1] == 1
        t = t[:, 0]
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


_counters = {}    # device -> persistent zero-initialised uint32 ticket counter of the fused loss kernel


def _ticket_counter(dev):
    c = _counters.get(dev)
    if c is None:
        c = _counters[dev] = torch.zeros((1,), dtype=torch.int32, device=dev)
    return c


def _quad_ok(lg: torch.Tensor, tg: torch.Tensor) -> bool:
    """the production shape of the fused loss kernels: C = 4 dense logits, V % 4 == 0, 16-byte aligned."""
    B, D, H, W, C = lg.shape
    return (C == 4 and cl_pitch(lg) == 4 and (D * H * W) % 4 == 0 and lg.data_ptr() % 16 == 0
            and tg.data_ptr() % 16 == 0)


class DiceCEMultiScaleFn(torch.autograd.Function):
    """sum over networks n and scales i of weight_i * (w_ce*CE + w_dice*SoftDice)(logits_n_i, target_i) --
    DeepSupervisionWrapper(DC_and_CE_loss) for one or several networks that share the targets.

    tensors = logits of network 0 (all scales), logits of network 1, ..., then the targets (one per scale).
    Production shape: ONE forward launch for everything (sums + last-block-done scalar algebra) and ONE backward launch
    (csrc/losses_multi.cu); other shapes run scale by scale (csrc/losses.cu)."""

    @staticmethod
    def forward(ctx, cfg: dict, *tensors):
        n_nets = int(cfg.get('n_nets', 1))
        n = len(tensors) // (n_nets + 1)
        targets = tensors[n_nets * n:]
        weights = cfg['weights']
        dev = tensors[0].device
        require_cuda(tensors[0], 'DC_and_CE_loss')
        st = _stream()
        items = []      # (tensor slot, scale, logits NDHWC, target fp32)
        tg_cache = {}
        for k in range(n_nets):
            for i in range(n):
                if weights[i] == 0:
                    continue
                if i not in tg_cache:
                    tg_cache[i] = _target_f32(targets[i])
                items.append((k * n + i, i, to_cl_view(tensors[k * n + i]), tg_cache[i]))
        ddp_bd = bool(cfg['batch_dice'] and cfg['ddp'] and torch.distributed.is_available()
                      and torch.distributed.is_initialized())
        gscale = float(torch.distributed.get_world_size()) if ddp_bd else 1.0
        B = items[0][2].shape[0]
        C = items[0][2].shape[-1]
        fused = (len(items) <= lib.DICE_CE_MAX_SEGMENTS and all(it[2].shape[0] == B and _quad_ok(it[2], it[3]) for it in items))
        ctx.cfg, ctx.items_meta, ctx.n_tensors, ctx.fused, ctx.gscale = cfg, [(it[0], it[1]) for it in items], len(tensors), fused, gscale
        ign = cfg.get('ignore_label')
        if ign is not None:
            # the kernels ignore every voxel whose target is outside [0, C): nnU-Net's ignore label sits behind the last class
            if 0 <= int(ign) < C:
                raise NotImplementedError('ignore_label inside the class range [0, %d) is outside the built hot path' % C)
            if not fused:
                raise NotImplementedError('ignore_label is built for the production shape of the loss kernels only '
                                          '(4 classes, dense logits, voxels % 4 == 0)')
        if fused:
            ns = len(items)
            segs = (DiceCESegment * ns)()
            for j, (_, i, lg, tg) in enumerate(items):
                V = lg.shape[1] * lg.shape[2] * lg.shape[3]
                segs[j].logits, segs[j].target, segs[j].dlogits, segs[j].V, segs[j].weight = \
                    lg.data_ptr(), tg.data_ptr(), None, V, float(weights[i])
            stride = B * C * 3 + 2          # per-(b,c) sums, CE sum, number of valid (not ignored) voxels
            # (the data-parallel batch_dice branch all-reduces and edits `acc` with in-place torch ops: those bump the
            # version counter of the whole tensor they alias, so it must not be a view of the shared zero pool, whose
            # other views are saved for backward elsewhere)
            acc = torch.zeros((ns, stride), dtype=torch.float64, device=dev) if ddp_bd else zeros((ns, stride), torch.float64, dev)
            coef = torch.empty((ns, B * C * 2 + 1), dtype=torch.float32, device=dev)   # (A, E) per (b, c) + 1 / valid
            loss = torch.empty((), dtype=torch.float32, device=dev)
            nbytes = sum(B * s.V * (2.0 * C + 4) for s in segs)
            args = (segs, ns, B, C, cfg['smooth'], int(cfg['do_bg']), int(cfg['batch_dice']), cfg['weight_ce'],
                    cfg['weight_dice'], acc.data_ptr(), coef.data_ptr(), loss.data_ptr())
            if not ddp_bd:
                _timed_mem('dice_ce_fwd', nbytes, lib.dice_ce_multi_fwd, *args, _ticket_counter(dev).data_ptr(), st)
            else:
                _timed_mem('dice_ce_fwd', nbytes, lib.dice_ce_multi_fwd, *args, None, st)
                ce = acc[:, stride - 2:].clone()         # the CE mean stays per rank; only the Dice sums are gathered
                torch.distributed.all_reduce(acc)        # AllGatherGrad(...).sum(0), ddp_allgather.py:35-48
                acc[:, stride - 2:] = ce
                lib.dice_ce_multi_finalize(*args, st)
            ctx.save_for_backward(coef, *[it[2] for it in items], *[it[3] for it in items])
            return loss
        # ---- generic shapes: scale by scale
        loss = zeros((1,), torch.float32, dev)[0]
        saved = []
        for (_, i, lg, tg) in items:
            Bi, D, H, W, Ci = lg.shape
            V = D * H * W
            acc = torch.zeros((Bi * Ci * 3 + 1,), dtype=torch.float64, device=dev) if ddp_bd else \
                zeros((Bi * Ci * 3 + 1,), torch.float64, dev)
            _timed_mem('dice_ce_fwd', Bi * V * (2.0 * Ci + 4), lib.dice_ce_fwd, lg.data_ptr(), cl_pitch(lg), tg.data_ptr(),
                       Bi, V, Ci, acc.data_ptr(), st)
            if ddp_bd:
                torch.distributed.all_reduce(acc[:Bi * Ci * 3])
            coef = torch.empty((Bi, Ci, 2), dtype=torch.float32, device=dev)
            lib.dice_ce_finalize(acc.data_ptr(), Bi, V, Ci, cfg['smooth'], int(cfg['do_bg']), int(cfg['batch_dice']),
                                 cfg['weight_ce'], cfg['weight_dice'], float(weights[i]), coef.data_ptr(),
                                 loss.data_ptr(), st)
            if gscale != 1.0:
                coef.mul_(gscale)
            saved += [lg, tg, coef]
        ctx.save_for_backward(*saved)
        return loss

    @staticmethod
    def backward(ctx, gout):
        cfg = ctx.cfg
        saved = ctx.saved_tensors
        st = _stream()
        gout = gout.contiguous().float()
        grads = [None] * ctx.n_tensors
        weights = cfg['weights']
        if ctx.fused:
            ns = len(ctx.items_meta)
            coef, lgs, tgs = saved[0], saved[1:1 + ns], saved[1 + ns:]
            B, C = lgs[0].shape[0], lgs[0].shape[-1]
            segs = (DiceCESegment * ns)()
            outs = []
            for j, ((slot, i), lg, tg) in enumerate(zip(ctx.items_meta, lgs, tgs)):
                dl = torch.empty(lg.shape, dtype=BF16, device=lg.device)
                outs.append(dl)
                segs[j].logits, segs[j].target, segs[j].dlogits = lg.data_ptr(), tg.data_ptr(), dl.data_ptr()
                segs[j].V, segs[j].weight = lg.shape[1] * lg.shape[2] * lg.shape[3], float(weights[i])
                if ctx.needs_input_grad[1 + slot]:
                    grads[slot] = ncdhw_view(dl)
            nbytes = sum(B * s.V * (4.0 * C + 4) for s in segs)
            _timed_mem('dice_ce_bwd', nbytes, lib.dice_ce_multi_bwd, segs, ns, B, C, coef.data_ptr(), cfg['weight_ce'],
                       ctx.gscale, gout.data_ptr(), st)
            return (None, *grads)
        for j, (slot, i) in enumerate(ctx.items_meta):
            if not ctx.needs_input_grad[1 + slot]:
                continue
            lg, tg, coef = saved[3 * j:3 * j + 3]
            B, D, H, W, C = lg.shape
            V = D * H * W
            dl = torch.empty(lg.shape, dtype=BF16, device=lg.device)
            _timed_mem('dice_ce_bwd', B * V * (4.0 * C + 4), lib.dice_ce_bwd, lg.data_ptr(), cl_pitch(lg), tg.data_ptr(),
                       B, V, C, coef.data_ptr(), cfg['weight_ce'], float(weights[i]), gout.data_ptr(),
                       dl.data_ptr(), C, st)
            grads[slot] = ncdhw_view(dl)
        return (None, *grads)


class KLFn(torch.autograd.Function):
    """distill_kl(y_s, y_t, T) of other_loss.py:51-64 on logits of logical shape [B,C,...].

    ``upstream``: the factor the caller is going to multiply the result with before it enters the total loss (lambda1 in
    MVDTrainer.py:925).  When given (and the logits have the production shape) the forward pass is ONE kernel that also
    writes both gradients, scaled for an upstream gradient of exactly ``upstream``; backward then only launches a rescale
    by gout / upstream, which returns immediately on the device when that ratio is 1."""

    @staticmethod
    def forward(ctx, ys, yt, T: float, upstream=None):
        require_cuda(ys, 'distill_kl')
        a, b = to_cl_view(ys), to_cl_view(yt)
        assert a.shape == b.shape
        B, D, H, W, C = a.shape
        NV = B * D * H * W
        width = 2 if C == 1 else C
        st = _stream()
        dev = a.device
        acc = zeros((1,), torch.float64, dev)
        loss = torch.empty((), dtype=torch.float32, device=dev)
        scale = float(T) ** 2 / float(NV * width)
        ctx.T, ctx.scale = float(T), scale
        need_a, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        ctx.fused = (upstream is not None and float(upstream) != 0.0 and (need_a or need_b) and C == 4
                     and cl_pitch(a) == 4 and cl_pitch(b) == 4 and NV % 4 == 0
                     and a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0)
        if ctx.fused:
            da = torch.empty(a.shape, dtype=BF16, device=dev) if need_a else None
            db = torch.empty(b.shape, dtype=BF16, device=dev) if need_b else None
            ctx.upstream = float(upstream)
            _timed_mem('kl_fwd_bwd', 2.0 * NV * C * (2 + need_a + need_b), lib.kl_fused, a.data_ptr(), b.data_ptr(), NV, C,
                       float(T), ctx.upstream * scale, acc.data_ptr(), _ptr(da), _ptr(db), st)
            lib.scalar_axpy(acc.data_ptr(), scale, loss.data_ptr(), 0, st)
            ctx.grads = (da, db)
            return loss
        _timed_mem('kl_fwd', 4.0 * NV * C, lib.kl_fwd, a.data_ptr(), cl_pitch(a), b.data_ptr(), cl_pitch(b), NV, C,
                   float(T), acc.data_ptr(), st)
        lib.scalar_axpy(acc.data_ptr(), scale, loss.data_ptr(), 0, st)
        ctx.save_for_backward(a, b)
        return loss

    @staticmethod
    def backward(ctx, gout):
        gout = gout.contiguous().float()
        if ctx.fused:
            da, db = ctx.grads
            ctx.grads = None
            n = (da if da is not None else db).numel()
            lib.rescale_bf16_pair(_ptr(da), _ptr(db), n, gout.data_ptr(), ctx.upstream, _stream())
            return (None if da is None else ncdhw_view(da)), (None if db is None else ncdhw_view(db)), None, None
        a, b = ctx.saved_tensors
        B, D, H, W, C = a.shape
        NV = B * D * H * W
        da = torch.empty(a.shape, dtype=BF16, device=a.device) if ctx.needs_input_grad[0] else None
        db = torch.empty(b.shape, dtype=BF16, device=a.device) if ctx.needs_input_grad[1] else None
        if da is not None or db is not None:
            _timed_mem('kl_bwd', 2.0 * NV * C * (2 + (da is not None) + (db is not None)), lib.kl_bwd, a.data_ptr(),
                       cl_pitch(a), b.data_ptr(), cl_pitch(b), NV, C, ctx.T, ctx.scale, gout.data_ptr(), _ptr(da), C,
                       _ptr(db), C, _stream())
        return (None if da is None else ncdhw_view(da)), (None if db is None else ncdhw_view(db)), None, None


class SoftmaxChannelFn(torch.autograd.Function):
    """softmax(logits, 1)[:, ch:ch+1] as fp32 [B,1,D,H,W] (the clDice term's prediction, MVDTrainer.py:904-908); with a
    ``target`` also the one-hot ground-truth channel (target == ch) as fp32, from the same launch (:904-905)."""

    @staticmethod
    def forward(ctx, logits, channel: int, target=None):
        require_cuda(logits, 'softmax_channel')
        lg = to_cl_view(logits)
        B, D, H, W, C = lg.shape
        prob = torch.empty((B, 1, D, H, W), dtype=torch.float32, device=lg.device)
        onehot = tg = None
        if target is not None:
            tg = _target_f32(target)
            onehot = torch.empty((B, 1, D, H, W), dtype=torch.float32, device=lg.device)
        lib.softmax_channel_fwd(lg.data_ptr(), cl_pitch(lg), _ptr(tg), B * D * H * W, C, channel, prob.data_ptr(),
                                _ptr(onehot), _stream())
        ctx.channel = channel
        ctx.save_for_backward(lg)
        if onehot is None:
            return prob
        ctx.mark_non_differentiable(onehot)
        return prob, onehot

    @staticmethod
    def backward(ctx, dprob, *unused):
        (lg,) = ctx.saved_tensors
        B, D, H, W, C = lg.shape
        dprob = dprob.contiguous().float()
        dl = torch.empty(lg.shape, dtype=BF16, device=lg.device)
        lib.softmax_channel_bwd(lg.data_ptr(), cl_pitch(lg), dprob.data_ptr(), B * D * H * W, C, ctx.channel,
                                dl.data_ptr(), C, _stream())
        return ncdhw_view(dl), None, None


def _vol_dims(t: torch.Tensor):
    """(B,1,D,H,W) / (B,C,D,H,W) fp32 -> (B*C, D, H, W)."""
    assert t.dim() == 5
    return t.shape[0] * t.shape[1], t.shape[2], t.shape[3], t.shape[4]


def _f32c(t):
    return t.contiguous().float() if (t.dtype != torch.float32 or not t.is_contiguous()) else t


class SoftErodeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img):
        require_cuda(img, 'soft_erode')
        img = _f32c(img)
        out = torch.empty_like(img)
        lib.soft_erode(img.data_ptr(), out.data_ptr(), *_vol_dims(img), _stream())
        ctx.save_for_backward(img)
        return out

    @staticmethod
    def backward(ctx, g):
        (img,) = ctx.saved_tensors
        g = _f32c(g)
        gin = torch.zeros_like(img)
        lib.soft_erode_bwd(img.data_ptr(), g.data_ptr(), gin.data_ptr(), *_vol_dims(img), _stream())
        return gin


class SoftDilateFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img):
        require_cuda(img, 'soft_dilate')
        img = _f32c(img)
        out = torch.empty_like(img)
        lib.soft_dilate(img.data_ptr(), out.data_ptr(), *_vol_dims(img), _stream())
        ctx.save_for_backward(img)
        return out

    @staticmethod
    def backward(ctx, g):
        (img,) = ctx.saved_tensors
        g = _f32c(g)
        gin = torch.zeros_like(img)
        lib.soft_dilate_bwd(img.data_ptr(), g.data_ptr(), 1.0, gin.data_ptr(), *_vol_dims(img), _stream())
        return gin


_SKEL_LEVELS_PER_PASS = 4   # levels kept on chip per launch (csrc/softskel.cu: kSkelMaxLevels)


def _skel_forward(img: torch.Tensor, iters: int, keep: bool):
    """runs soft_skel's iters + 1 levels with the fused on-chip kernel (<= 4 levels per launch); returns
    (skel, E stack [L] = E_1..E_L, skel stack [L]) -- the stacks (what the backward needs; E_0 is the input itself and
    the deltas are recomputed there) only if ``keep``."""
    dims = _vol_dims(img)
    N = img.numel()
    L = iters + 1
    st = _stream()
    dev = img.device
    flat = img.reshape(-1)
    if keep:
        E = torch.empty((L, N), dtype=torch.float32, device=dev)
        skel = torch.empty((L, N), dtype=torch.float32, device=dev)
    else:
        E = skel = None
        tmpE = [torch.empty((N,), dtype=torch.float32, device=dev) for _ in range(2 if L > _SKEL_LEVELS_PER_PASS else 0)]
        tmpS = [torch.empty((N,), dtype=torch.float32, device=dev) for _ in range(2 if L > _SKEL_LEVELS_PER_PASS else 1)]
    PtrArr = ctypes.c_void_p * _SKEL_LEVELS_PER_PASS
    e_in, s_in, last = flat, None, None
    for pi, j0 in enumerate(range(0, L, _SKEL_LEVELS_PER_PASS)):
        n = min(_SKEL_LEVELS_PER_PASS, L - j0)
        more = j0 + n < L
        en, dl, sk = PtrArr(), PtrArr(), PtrArr()
        e_out = None
        for l in range(n):
            if keep:
                en[l], sk[l] = E[j0 + l].data_ptr(), skel[j0 + l].data_ptr()
            elif l == n - 1:
                last = tmpS[pi % len(tmpS)]
                sk[l] = last.data_ptr()
                if more:
                    e_out = tmpE[pi % 2]
                    en[l] = e_out.data_ptr()
        lib.soft_skel_fused(e_in.data_ptr(), _ptr(s_in), n, en, dl, sk, *dims, st)
        if keep:
            e_in, s_in = E[j0 + n - 1], skel[j0 + n - 1]
        else:
            e_in, s_in = e_out, last
    out = skel[L - 1] if keep else last
    return out.view(img.shape), E, skel


_SKEL_BWD_LEVELS = 2     # levels per backward launch (csrc/softskel.cu: kSkelBwdLevels)


def _skel_backward(g_skel: torch.Tensor, img_flat: torch.Tensor, E, skel, dims):
    """gradient of soft_skel w.r.t. its input, given d/d(skel): fused launches of <= 2 levels from the deepest level
    down, scatter-adds in shared memory; between launches only the chain state G and the partial gradient of the
    topmost E volume of the next launch travel through HBM."""
    L, N = E.shape
    st = _stream()
    dev = g_skel.device
    Ev = [img_flat] + [E[j] for j in range(L)]          # E_0 .. E_L
    PtrE = ctypes.c_void_p * (_SKEL_BWD_LEVELS + 1)
    PtrS = ctypes.c_void_p * _SKEL_BWD_LEVELS
    G, g_top, hi = g_skel, None, L
    while hi > 0:
        n = min(_SKEL_BWD_LEVELS, hi)
        a = hi - n
        pe, ps = PtrE(), PtrS()
        for l in range(n + 1):
            pe[l] = Ev[a + l].data_ptr()
        for l in range(n):
            ps[l] = skel[a + l - 1].data_ptr() if a + l >= 1 else None
        g_out = torch.empty((N,), dtype=torch.float32, device=dev)
        G_out = torch.empty((N,), dtype=torch.float32, device=dev) if a > 0 else None
        lib.soft_skel_bwd_fused(pe, ps, n, int(a == 0), G.data_ptr(), _ptr(g_top), g_out.data_ptr(), _ptr(G_out), *dims, st)
        G, g_top, hi = G_out, g_out, a
    return g_top


class SoftSkelFn(torch.autograd.Function):
    """soft_skel(img, iter_) of soft_skeleton.py:29-37."""

    @staticmethod
    def forward(ctx, img, iters: int):
        require_cuda(img, 'soft_skel')
        img = _f32c(img)
        need = ctx.needs_input_grad[0]
        sk, E, skel = _skel_forward(img, iters, need)
        ctx.dims, ctx.shape = _vol_dims(img), img.shape
        if need:
            ctx.save_for_backward(img, E, skel)
        return sk.clone() if need else sk

    @staticmethod
    def backward(ctx, g):
        img, E, skel = ctx.saved_tensors
        g = _f32c(g).reshape(-1)
        return _skel_backward(g, img.reshape(-1), E, skel, ctx.dims).view(ctx.shape), None


class SoftClDiceFn(torch.autograd.Function):
    """soft_cldice(iter_, smooth)(y_true, y_pred) -- fused: both skeletons, the four sums, the scalar and (in
    backward) the whole chain down to d/d(y_pred)."""

    @staticmethod
    def forward(ctx, y_true, y_pred, iters: int, smooth: float):
        require_cuda(y_pred, 'soft_cldice')
        y_true, y_pred = _f32c(y_true), _f32c(y_pred)
        assert y_true.shape == y_pred.shape
        dev = y_pred.device
        st = _stream()
        N = y_pred.numel()
        need = ctx.needs_input_grad[1]
        sp, E, skel = _skel_forward(y_pred, iters, need)
        stv, _, _ = _skel_forward(y_true, iters, False)
        sums = zeros((4,), torch.float64, dev)
        lib.dot_sum(sp.data_ptr(), y_true.data_ptr(), N, sums.data_ptr(), st)
        lib.dot_sum(stv.data_ptr(), y_pred.data_ptr(), N, sums[2:].data_ptr(), st)
        out4 = torch.empty((4,), dtype=torch.float32, device=dev)
        lib.cldice_finalize(sums.data_ptr(), float(smooth), out4.data_ptr(), st)
        ctx.dims, ctx.shape = _vol_dims(y_pred), y_pred.shape
        if need:
            ctx.save_for_backward(y_pred, E, skel, y_true, stv.reshape(-1), out4)
        return out4[0]

    @staticmethod
    def backward(ctx, gout):
        y_pred, E, skel, y_true, stv, out4 = ctx.saved_tensors
        st = _stream()
        N = y_true.numel()
        gout = gout.contiguous().float()
        g_skel = torch.empty((N,), dtype=torch.float32, device=y_true.device)
        lib.cldice_seed(y_true.data_ptr(), out4.data_ptr(), gout.data_ptr(), g_skel.data_ptr(), N, st)
        gE0 = _skel_backward(g_skel, y_pred.reshape(-1), E, skel, ctx.dims)
        dp = torch.empty((N,), dtype=torch.float32, device=y_true.device)
        lib.cldice_combine(gE0.data_ptr(), stv.data_ptr(), out4.data_ptr(), gout.data_ptr(), dp.data_ptr(), N, st)
        return None, dp.view(ctx.shape), None, None


def argmax_tp_fp_fn(logits: torch.Tensor, target: torch.Tensor):
    """validation_step's hard tp/fp/fn over axes [0,2,3,4] (nnUNetTrainer.py:973-1004). Returns three [C] tensors."""
    lg = to_cl_view(logits)
    B, D, H, W, C = lg.shape
    tg = _target_f32(target)
    out = torch.zeros((C, 3), dtype=torch.float64, device=lg.device)
    lib.argmax_tp_fp_fn(lg.data_ptr(), cl_pitch(lg), tg.data_ptr(), B, D * H * W, C, out.data_ptr(), _stream())
    return out[:, 0], out[:, 1], out[:, 2]
