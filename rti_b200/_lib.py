"""ctypes binding of libptm_b200.so (C ABI in include/ptm_b200.h).

There is no fallback: if the CUDA library has not been built, or no B200 is
present, every entry point raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PTM_B200_LIB") or os.path.join(HERE, "lib", "libptm_b200.so")   # override: tuning builds only

NCOEF = 6
NKEYS = 12
U8, U16, U32 = 0, 1, 2
SYNTH_YCBCR, SYNTH_RGB = 0, 1
FIT_LUM, FIT_RGB_INPUT = 1, 2


class PtmError(RuntimeError):
    pass


class ScaleBias(C.Structure):
    _fields_ = [("scale", C.c_float * NCOEF), ("bias", C.c_int32 * NCOEF)]


_vp, _sz, _u32, _int = C.c_void_p, C.c_size_t, C.c_uint32, C.c_int

_SIGNATURES = {
    "ptmb_last_error": ([], C.c_char_p),
    "ptmb_device_count": ([C.POINTER(C.c_int)], _int),
    "ptmb_create": ([_int, _vp, C.POINTER(_vp)], _int),
    "ptmb_destroy": ([_vp], None),
    "ptmb_set_stream": ([_vp, _vp], _int),
    "ptmb_synchronize": ([_vp], _int),
    "ptmb_sm_count": ([_vp], _int),
    "ptmb_stream": ([_vp], _vp),
    "ptmb_device": ([_vp], _int),
    "ptmb_malloc": ([_vp, _sz, C.POINTER(_vp)], _int),
    "ptmb_free": ([_vp, _vp], _int),
    "ptmb_upload": ([_vp, _vp, _vp, _sz], _int),
    "ptmb_download": ([_vp, _vp, _vp, _sz], _int),
    "ptmb_host_register": ([_vp, _sz], _int),
    "ptmb_host_unregister": ([_vp], _int),
    "ptmb_host_alloc": ([_sz, _int, C.POINTER(_vp)], _int),
    "ptmb_host_free": ([_vp], _int),
    "ptmb_set_matrix": ([_vp, _vp, C.c_uint], _int),
    "ptmb_set_fit_engine": ([_vp, _int], _int),
    "ptmb_mma_launches": ([], C.c_ulonglong),
    "ptmb_mma_pack_digits": ([_vp, C.c_uint, _vp, _sz, _vp, _vp], _int),
    "ptmb_keys_init": ([_vp, _vp], _int),
    "ptmb_fit_lrgb": ([_vp, _vp, _int, _sz, _sz, _vp, _vp, _vp], _int),
    "ptmb_set_stack_stable": ([_vp, _int], _int),
    "ptmb_set_coeffs_scratch": ([_vp, _int], _int),
    "ptmb_set_fit_mode": ([_vp, _int], _int),
    "ptmb_fit_rgb": ([_vp, _vp, _int, _sz, _sz, _vp, _sz, _vp], _int),
    "ptmb_fit_channel": ([_vp, _vp, _int, _sz, _sz, _sz, _vp], _int),
    "ptmb_chroma_avg": ([_vp, _vp, _sz, _sz, C.c_uint, _vp], _int),
    "ptmb_lrgb_colour_normalise": ([_vp, _vp, _vp, _sz], _int),
    "ptmb_minmax": ([_vp, _vp, _sz, _vp], _int),
    "ptmb_quantise": ([_vp, _vp, _sz, _vp, _vp, _vp], _int),
    "ptmb_quantise_consume": ([_vp, _vp, _sz, _vp, _vp, _vp], _int),
    "ptmb_set_relight_mode": ([_vp, _int], _int),
    "ptmb_relight_lrgb": ([_vp, _vp, _vp, _sz, _sz, _vp, _vp, _vp, C.c_uint, _vp], _int),
    "ptmb_relight_rgb": ([_vp, _vp, _vp, _vp, _sz, _sz, _vp, _vp, _vp, C.c_uint, _vp], _int),
    "ptmb_relight_lum": ([_vp, _vp, _vp, _sz, _sz, _vp, _vp, _vp, C.c_uint, _vp], _int),
    "ptmb_normal_map": ([_vp, _vp, _sz, _vp], _int),
    "ptmb_normal_map_blocks": ([_vp, _vp, _vp, _vp, _sz, _vp], _int),
    "ptmb_fit_lsum": ([_vp, _vp, _sz, _sz, _int, _vp, _vp, _vp], _int),
    "ptmb_selftest_div255": ([_vp, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64)], _int),
    "ptmb_synth": ([_vp, _u32, _int, _int, _u32, _sz, C.c_uint, _vp, _vp, _sz], _int),
    "ptmb_encode_host_early_rgb": ([_vp, _vp], _int),
    "ptmb_encode_host_fit": ([_vp, _vp, _int, _sz, _sz, C.c_uint, _vp, _int, _vp], _int),
    "ptmb_encode_host_finish": ([_vp, _vp, C.POINTER(_vp), _vp, _vp], _int),
    "ptmb_encode_host_release": ([_vp], _int),
    "ptmb_encode_host": ([_vp, _vp, _int, _sz, _sz, C.c_uint, _vp, _int, C.POINTER(_vp), _vp, _vp], _int),
    "ptmb_xchg_create": ([_vp, _int, _int, C.POINTER(_vp), _vp], _int),
    "ptmb_xchg_connect": ([_vp, _vp], _int),
    "ptmb_xchg_allreduce_max": ([_vp, _vp, _vp], _int),
    "ptmb_xchg_status": ([_vp, _vp, C.POINTER(C.c_int)], _int),
    "ptmb_xchg_destroy": ([_vp], None),
    "ptmb_encode_lrgb_dev": ([_vp, _vp, _vp, _int, _sz, _sz, _vp, _vp, _vp, _vp], _int),
    "ptmb_multi_create": ([_vp, _int, C.POINTER(_vp)], _int),
    "ptmb_multi_destroy": ([_vp], None),
    "ptmb_multi_device_count": ([_vp], _int),
    "ptmb_multi_ctx": ([_vp, _int], _vp),
    "ptmb_multi_slab": ([_vp, _sz, _int, C.POINTER(_sz), C.POINTER(_sz)], _int),
    "ptmb_multi_host_slab": ([_vp, _sz, _int, C.POINTER(_sz), C.POINTER(_sz)], _int),
    "ptmb_multi_set_host_weights": ([_vp, _vp], _int),
    "ptmb_multi_calibrate_h2d": ([_vp, C.c_double, _vp], _int),
    "ptmb_multi_set_matrix": ([_vp, _vp, C.c_uint], _int),
    "ptmb_multi_set_fit_mode": ([_vp, _int], _int),
    "ptmb_multi_synchronize": ([_vp], _int),
    "ptmb_multi_encode_host": ([_vp, _vp, _int, _sz, _sz, C.c_uint, _vp, _int, C.POINTER(_vp), _vp, _vp], _int),
    "ptmb_multi_encode_lrgb_dev": ([_vp, _vp, _int, _vp, _vp, _vp, _vp, _vp, _vp], _int),
    "ptmb_encode_rgb_dev": ([_vp, _vp, _vp, _int, _sz, _sz, _vp, _vp, _vp], _int),
    "ptmb_probe_arm": ([_vp, C.c_uint], _int),
    "ptmb_probe_read": ([_vp, _vp, C.c_uint, C.POINTER(C.c_uint)], _int),
    "ptmb_jpeg_info": ([_vp, _vp, _sz, C.POINTER(C.c_uint), C.POINTER(C.c_uint), C.POINTER(C.c_uint)], _int),
    "ptmb_jpeg_decode": ([_vp, _vp, _sz, _int, _int, _vp, _int], _int),
    "ptmb_jpeg_encode": ([_vp, _vp, C.c_uint, C.c_uint, C.c_uint, _int, _int, C.POINTER(_vp), C.POINTER(_sz)], _int),
    "ptmb_fit_channel_host": ([_vp, _vp, _int, _sz, _sz, _sz, C.c_uint, _vp, _vp], _int),
    "ptmb_chroma_avg_host": ([_vp, _vp, _sz, C.c_uint, _vp], _int),
    "ptmb_lrgb_colour_normalise_host": ([_vp, _vp, _vp, _sz], _int),
    "ptmb_scale_host": ([_vp, _vp, _sz, _int, C.POINTER(_vp), _vp], _int),
    "ptmb_normal_map_host": ([_vp, _vp, _vp, _vp, _vp, _sz, _vp], _int),
    "ptmb_relight_host": ([_vp, C.POINTER(_vp), _int, _sz, _sz, _vp, _vp, _vp, C.c_uint, _vp], _int),
}

EXPORTS = tuple(sorted(_SIGNATURES))

_lib = None


def load():
    """Load the CUDA library; raises PtmError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PtmError("%s is missing -- build it with `make -C rti_b200/csrc` "
                       "(or __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (argtypes, restype) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header and library disagree
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise PtmError(load().ptmb_last_error().decode("utf-8", "replace"))
