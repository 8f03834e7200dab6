"""ctypes binding of libmvdseg.so (the C ABI declared in include/mvdseg.h).

The library is the product path: there is no CPU or PyTorch fallback.  Importing this module only needs the shared
object to exist; calling any compute entry point without a CUDA device raises.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, c_char_p, c_double, c_float, c_int, c_longlong, c_size_t, c_uint64,
                    c_ulonglong, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libmvdseg.so')


class MvdError(RuntimeError):
    pass


class ConvArgs(Structure):
    """mirror of mvd_conv3d_args (include/mvdseg.h)."""
    _fields_ = [
        ('B', c_int),
        ('Di', c_int), ('Hi', c_int), ('Wi', c_int), ('Cin', c_int),
        ('Do', c_int), ('Ho', c_int), ('Wo', c_int), ('Cout', c_int),
        ('kd', c_int), ('kh', c_int), ('kw', c_int),
        ('sd', c_int), ('sh', c_int), ('sw', c_int),
        ('pd', c_int), ('ph', c_int), ('pw', c_int),
        ('x', c_void_p), ('ldx', c_int),
        ('y', c_void_p), ('ldy', c_int),
        ('w', c_void_p),
        ('bias', c_void_p),
        ('stats', c_void_p),
        ('dw', c_void_p),
        ('dbias', c_void_p),
        ('workspace', c_void_p), ('workspace_bytes', c_size_t),
        ('algo', c_int),
        ('accumulate', c_int),
        ('norm_bwd', c_void_p),
    ]


class NormBwdStatsArgs(Structure):
    """mirror of mvd_norm_bwd_stats_args (include/mvdseg.h)."""
    _fields_ = [('y', c_void_p), ('ldy', c_int), ('stats', c_void_p), ('gamma', c_void_p), ('beta', c_void_p),
                ('eps', c_float), ('slope', c_float), ('bstats', c_void_p)]


class DiceCESegment(Structure):
    """mirror of mvd_dice_ce_segment (include/mvdseg.h)."""
    _fields_ = [('logits', c_void_p), ('target', c_void_p), ('dlogits', c_void_p), ('V', c_longlong),
                ('weight', c_float)]


P = c_void_p
I = c_int
LL = c_longlong
F = c_float
S = c_void_p  # stream

# name -> (restype, argtypes); every int-returning compute call is wrapped by _checked
_SIGNATURES = {
    'mvd_version': (c_int, []),
    'mvd_last_error': (c_char_p, []),
    'mvd_spin': (c_int, [LL, S]),
    'mvd_launch_count': (c_ulonglong, []),
    'mvd_reset_launch_count': (None, []),
    'mvd_fallback_count': (c_ulonglong, []),
    'mvd_set_deterministic': (c_int, [I]),
    'mvd_get_deterministic': (c_int, []),
    'mvd_reset_fallback_count': (None, []),
    'mvd_shutdown': (c_int, []),
    'mvd_ncdhw_f32_to_ndhwc_bf16': (c_int, [P, LL, P, I, I, LL, I, S]),
    'mvd_ndhwc_bf16_to_ncdhw_f32': (c_int, [P, I, P, I, I, LL, S]),
    'mvd_sw_accumulate': (c_int, [P, I, P, F, P, P, I, I, I, I, I, I, I, I, I, I, I, S]),
    'mvd_sw_finalize': (c_int, [P, P, I, LL, S]),
    'mvd_downsample_seg_nearest': (c_int, [P, I, I, I, I, I, P, P, S]),
    'mvd_stem_conv_fprop': (c_int, [P, I, I, I, I, I, P, P, P, I, P, S]),
    'mvd_stem_conv_wgrad': (c_int, [P, I, I, I, I, I, P, I, P, S]),
    'mvd_pack_conv_weights_multi': (c_int, [P, I, I, S]),
    'mvd_pack_blocks': (c_int, [I, I]),
    'mvd_conv3d_workspace_bytes': (c_size_t, [POINTER(ConvArgs), I]),
    'mvd_conv3d_fprop': (c_int, [POINTER(ConvArgs), S]),
    'mvd_conv3d_dgrad': (c_int, [POINTER(ConvArgs), S]),
    'mvd_conv3d_wgrad': (c_int, [POINTER(ConvArgs), S]),
    'mvd_conv3d_dgrad_fuses_norm_bwd': (c_int, [POINTER(ConvArgs)]),
    'mvd_inorm_stats': (c_int, [P, I, I, LL, I, P, S]),
    'mvd_inorm_lrelu_fwd': (c_int, [P, I, P, I, P, P, P, I, LL, I, F, F, S]),
    'mvd_inorm_lrelu_bwd_stats': (c_int, [P, I, P, I, P, P, P, I, LL, I, F, F, P, S]),
    'mvd_inorm_lrelu_bwd_apply': (c_int, [P, I, P, I, P, I, P, P, P, P, I, LL, I, F, F, P, P, P, S]),
    'mvd_inorm_lrelu_head_supported': (c_int, [I, I]),
    'mvd_inorm_lrelu_head_fwd': (c_int, [P, I, P, P, P, P, P, P, I, LL, I, I, F, F, S]),
    'mvd_inorm_lrelu_head_bwd_stats': (c_int, [P, P, I, P, P, P, P, I, LL, I, I, F, F, P, P, P, S]),
    'mvd_inorm_lrelu_head_bwd_apply': (c_int, [P, P, I, P, I, P, P, P, P, P, I, LL, I, I, F, F, P, P, P, S]),
    'mvd_head_fwd': (c_int, [P, I, P, P, P, I, LL, I, I, S]),
    'mvd_head_bwd': (c_int, [P, I, P, I, P, P, I, P, P, LL, I, I, I, S]),
    'mvd_dice_ce_multi_fwd': (c_int, [P, I, I, I, F, I, I, F, F, P, P, P, P, S]),
    'mvd_dice_ce_multi_finalize': (c_int, [P, I, I, I, F, I, I, F, F, P, P, P, S]),
    'mvd_dice_ce_multi_bwd': (c_int, [P, I, I, I, P, F, F, P, S]),
    'mvd_kl_fused': (c_int, [P, P, LL, I, F, F, P, P, P, S]),
    'mvd_rescale_bf16_pair': (c_int, [P, P, LL, P, F, S]),
    'mvd_dice_ce_fwd': (c_int, [P, I, P, I, LL, I, P, S]),
    'mvd_dice_ce_finalize': (c_int, [P, I, LL, I, F, I, I, F, F, F, P, P, S]),
    'mvd_dice_ce_bwd': (c_int, [P, I, P, I, LL, I, P, F, F, P, P, I, S]),
    'mvd_argmax_tp_fp_fn': (c_int, [P, I, P, I, LL, I, P, S]),
    'mvd_kl_fwd': (c_int, [P, I, P, I, LL, I, F, P, S]),
    'mvd_kl_bwd': (c_int, [P, I, P, I, LL, I, F, F, P, P, I, P, I, S]),
    'mvd_soft_erode': (c_int, [P, P, I, I, I, I, S]),
    'mvd_soft_dilate': (c_int, [P, P, I, I, I, I, S]),
    'mvd_soft_erode_bwd': (c_int, [P, P, P, I, I, I, I, S]),
    'mvd_soft_dilate_bwd': (c_int, [P, P, F, P, I, I, I, I, S]),
    'mvd_skel_update': (c_int, [P, P, P, P, P, I, I, I, I, I, S]),
    'mvd_soft_skel_fused': (c_int, [P, P, I, P, P, P, I, I, I, I, S]),
    'mvd_soft_skel_bwd_fused': (c_int, [P, P, I, I, P, P, P, P, I, I, I, I, S]),
    'mvd_skel_chain_bwd': (c_int, [P, P, P, P, I, LL, S]),
    'mvd_skel_level_bwd': (c_int, [P, P, P, P, P, I, I, I, I, S]),
    'mvd_dot_sum': (c_int, [P, P, LL, P, S]),
    'mvd_softmax_channel_fwd': (c_int, [P, I, P, LL, I, I, P, P, S]),
    'mvd_softmax_channel_bwd': (c_int, [P, I, P, LL, I, I, P, I, S]),
    'mvd_cldice_seed': (c_int, [P, P, P, P, LL, S]),
    'mvd_cldice_combine': (c_int, [P, P, P, P, P, LL, S]),
    'mvd_cldice_finalize': (c_int, [P, F, P, S]),
    'mvd_grad_sqnorm': (c_int, [P, P, P, P, I, P, S]),
    'mvd_sgd_pack_conv_weights': (c_int, [P, I, I, P, F, F, F, F, F, S]),
    'mvd_sgd_nesterov_clip': (c_int, [P, P, P, P, I, P, F, F, F, F, F, S]),
    'mvd_stats_channel_sum': (c_int, [P, I, I, I, I, P, S]),
    'mvd_zero_regions': (c_int, [P, P, I, S]),
    'mvd_zero_bytes': (c_int, [P, c_size_t, S]),
    'mvd_channel_sum': (c_int, [P, I, LL, I, P, S]),
    'mvd_scalar_axpy': (c_int, [P, F, P, I, S]),
    'mvd_add_bf16': (c_int, [P, I, P, I, LL, I, S]),
    'mvd_aug_spline_prefilter': (c_int, [P, I, I, I, I, P, S]),
    'mvd_aug_spatial': (c_int, [P, I, I, I, I, I, P, I, I, I, P, P, I, F, I, S]),
    'mvd_aug_gaussian_noise': (c_int, [P, LL, I, P, c_ulonglong, S]),
    'mvd_aug_gaussian_blur': (c_int, [P, P, I, I, I, I, P, S]),
    'mvd_aug_plane_stats': (c_int, [P, LL, I, P, S]),
    'mvd_aug_intensity': (c_int, [P, LL, I, I, P, P, P, I, S]),
    'mvd_aug_simulate_lowres': (c_int, [P, I, I, I, I, P, P, LL, P, S]),
    'mvd_aug_mirror': (c_int, [P, P, I, I, I, I, I, P, S]),
    'mvd_im2col_small': (c_int, [P, I, I, I, I, I, I, I, I, I, I, I, I, P, I, S]),
}

_UNCHECKED = {'mvd_version', 'mvd_last_error', 'mvd_launch_count', 'mvd_reset_launch_count', 'mvd_fallback_count',
              'mvd_reset_fallback_count', 'mvd_get_deterministic', 'mvd_inorm_lrelu_head_supported',
              'mvd_conv3d_workspace_bytes', 'mvd_pack_blocks', 'mvd_conv3d_dgrad_fuses_norm_bwd'}


def _load():
    if not os.path.exists(LIB_PATH):
        raise MvdError(f'{LIB_PATH} is missing: build it with `python -c "import __graft_entry__ as g; g.build()"` '
                       f'(or `make -C multimodal_mvd_seg_b200/csrc`). There is no fallback path.')
    return ctypes.CDLL(LIB_PATH)


_cdll = _load()


class _Lib:
    def __init__(self, cdll):
        self._cdll = cdll
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(cdll, name)
            fn.restype = res
            fn.argtypes = args
            short = name[len('mvd_'):]
            if name in _UNCHECKED or res is not c_int:
                setattr(self, short, fn)
            else:
                setattr(self, short, self._checked(fn, name))

    def _checked(self, fn, name):
        last_error = self._cdll.mvd_last_error

        def call(*a):
            rc = fn(*a)
            if rc != 0:
                msg = last_error()
                raise MvdError(f'{name} failed ({rc}): {msg.decode() if msg else ""}')
            return rc
        call.__name__ = name
        return call


lib = _Lib(_cdll)
lib.DICE_CE_MAX_SEGMENTS = 16   # MVD_DICE_CE_MAX_SEGMENTS (include/mvdseg.h)


def exported_symbols():
    return list(_SIGNATURES.keys())
