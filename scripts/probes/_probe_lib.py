"""ctypes loader of scripts/probes/libmvdseg_probe.so (`make -C multimodal_mvd_seg_b200/csrc probe`)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(_HERE, 'libmvdseg_probe.so')
if not os.path.exists(_PATH):
    raise RuntimeError(f'{_PATH} missing: run `make -C multimodal_mvd_seg_b200/csrc probe`')
_cdll = ctypes.CDLL(_PATH)
P, I, S = ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p
_cdll.mvd_tc_probe.argtypes = [P, P, I, I, I, I, I, I, I, P, S]
_cdll.mvd_tc_probe.restype = I
_cdll.mvd_tc_mma_bench.argtypes = [I] * 11 + [P, S]
_cdll.mvd_tc_mma_bench.restype = I
_cdll.mvd_last_error.restype = ctypes.c_char_p


def _checked(fn):
    def call(*a):
        rc = fn(*a)
        if rc != 0:
            raise RuntimeError(f'{fn.__name__} failed ({rc}): {_cdll.mvd_last_error().decode()}')
    return call


tc_probe = _checked(_cdll.mvd_tc_probe)
tc_mma_bench = _checked(_cdll.mvd_tc_mma_bench)
