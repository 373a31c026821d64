"""Host-side mirror of the reference's ptmlib.h interface (include/ptmlib.h), bound to
libptm.so (rti_b200/csrc/ptmlib_host.c).  Same function names, argument meaning and error
behaviour as the reference's C API, with numpy arrays as the host buffers:

    ptm_svd, ptm_fit_poly_jsample, ptm_fit_poly_uint, ptm_cbcr_avg, ptm_scale_coefficients,
    ptm_write_ptm / ptm_read_header / ptm_read_ptm, ptm_write_jpeg, ptm_normal, ...

ptm_svd runs on the host (LAPACKE); everything per-pixel runs on the GPU.  There is no CPU
fallback: on a machine without a GPU the per-pixel calls abort with the CUDA error, exactly
like the C library (messages on stderr, exit status 1).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libptm.so")

PTM_COEFFICIENTS = 6
RGB_COEFFICIENTS = 3
MAX_JPEG_STREAMS = 18


class PtmFormat(C.Structure):
    # include/ptmlib.h ptm_format_t
    _fields_ = [("id", C.c_int), ("blocks", C.c_int), ("ptm_blocks", C.c_int), ("color_components", C.c_int),
                ("jpeg_streams", C.c_int), ("name", C.c_char_p)]


class PtmHeader(C.Structure):
    # include/ptmlib.h ptm_header_t
    _fields_ = [("format", C.POINTER(PtmFormat)), ("dimen", C.c_size_t * 2), ("scale", C.c_float * 6),
                ("bias", C.c_int * 6), ("compression_param", C.c_int * 1), ("transforms", C.c_int * 18),
                ("motion_vector_x", C.c_int * 18), ("motion_vector_y", C.c_int * 18), ("order", C.c_int * 18),
                ("reference_planes", C.c_int * 18), ("compressed_size", C.c_size_t * 18),
                ("side_info_sizes", C.c_size_t * 18)]


class PtmImageInfo(C.Structure):
    # include/ptmlib.h ptm_image_info_t
    _fields_ = [("width", C.c_size_t), ("height", C.c_size_t), ("pixels", C.c_size_t),
                ("row_stride", C.c_size_t), ("decoder_stride", C.c_size_t), ("n_decoders", C.c_uint)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("%s is missing -- build with `make -C rti_b200/csrc`" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.ptm_get_format.restype = C.POINTER(PtmFormat)
        L.ptm_get_format.argtypes = [C.c_char_p]
        L.ptm_svd_uv.restype = C.POINTER(C.c_float)
        L.ptm_svd_uv.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.ptm_fit_poly_jsample.argtypes = [C.POINTER(PtmImageInfo), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.ptm_fit_poly_uint.argtypes = [C.POINTER(PtmImageInfo), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.ptm_cbcr_avg.argtypes = [C.POINTER(PtmImageInfo), C.c_void_p, C.c_void_p]
        L.ptm_lrgb_finish.argtypes = [C.POINTER(PtmImageInfo), C.c_void_p, C.c_void_p]
        L.ptm_scale_coefficients.argtypes = [C.POINTER(PtmHeader), C.c_void_p, C.POINTER(C.c_void_p)]
        L.ptm_encode_stack.argtypes = [C.POINTER(PtmHeader), C.POINTER(PtmImageInfo), C.c_void_p, C.c_void_p,
                                       C.POINTER(C.c_void_p)]
        L.ptm_relight.argtypes = [C.POINTER(PtmHeader), C.POINTER(C.c_void_p), C.c_float, C.c_float, C.c_void_p]
        L.ptm_normal.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ptm_normal_map.argtypes = [C.POINTER(PtmHeader), C.POINTER(C.c_void_p), C.c_void_p]
        L.ptm_normal_map.restype = None
        for f in ("ptm_fit_poly_jsample", "ptm_fit_poly_uint", "ptm_cbcr_avg", "ptm_lrgb_finish",
                  "ptm_scale_coefficients", "ptm_encode_stack", "ptm_relight", "ptm_normal"):
            getattr(L, f).restype = None
        _lib = L
    return _lib


def ptm_formats():
    """Names of the supported formats, in table order."""
    arr = (PtmFormat * 6).in_dll(lib(), "ptm_formats")
    return [f.name.decode() for f in arr if f.name]


def ptm_get_format(name):
    """-> PtmFormat pointer, or None for an unknown name (like the C NULL)."""
    p = lib().ptm_get_format(name.encode())
    return p if p else None


def ptm_alloc_header(format_name, width, height):
    fmt = ptm_get_format(format_name)
    if fmt is None:
        raise ValueError("No format by that name: %s" % format_name)
    h = PtmHeader()
    h.format = fmt
    h.dimen[0], h.dimen[1] = width, height
    return h


def ptm_alloc_blocks(header):
    """numpy blocks as ptm_alloc_blocks would size them: 6 or 3 bytes per pixel."""
    f = header.format.contents
    P = header.dimen[0] * header.dimen[1]
    return [np.zeros((P, PTM_COEFFICIENTS if i < f.ptm_blocks else RGB_COEFFICIENTS), np.uint8)
            for i in range(f.blocks)]


def image_info(width, height, n_lights, components=3):
    return PtmImageInfo(width, height, width * height, width * components, width * height * components, n_lights)


def ptm_svd(lights):
    """6 x n pseudo-inverse of the light matrix from (u, v[, w]) rows -- host LAPACKE."""
    l = np.ascontiguousarray(lights, dtype=np.float32)
    u = np.ascontiguousarray(l[:, 0])
    v = np.ascontiguousarray(l[:, 1])
    n = l.shape[0]
    p = lib().ptm_svd_uv(u.ctypes.data, v.ctypes.data, n)
    if not p:
        raise RuntimeError("Error in Singular Value Decomposition")
    M = np.ctypeslib.as_array(p, shape=(PTM_COEFFICIENTS, n)).copy()
    C.CDLL(None).free(C.cast(p, C.c_void_p))
    return M


def ptm_fit_poly_jsample(info, buffer, pixel_stride, M, channel=0):
    """buffer uint8 [n][H][W][pixel_stride]; fits `channel` -> float32 [H*W][6]."""
    buffer = np.ascontiguousarray(buffer, dtype=np.uint8)
    M = np.ascontiguousarray(M, dtype=np.float32)
    out = np.empty((info.pixels, PTM_COEFFICIENTS), np.float32)
    lib().ptm_fit_poly_jsample(C.byref(info), buffer.ctypes.data + channel, pixel_stride, M.ctypes.data,
                               out.ctypes.data)
    return out


def ptm_fit_poly_uint(info, buffer, pixel_stride, M, channel=0):
    buffer = np.ascontiguousarray(buffer, dtype=np.uint32)
    M = np.ascontiguousarray(M, dtype=np.float32)
    out = np.empty((info.pixels, PTM_COEFFICIENTS), np.float32)
    lib().ptm_fit_poly_uint(C.byref(info), buffer.ctypes.data + 4 * channel, pixel_stride, M.ctypes.data,
                            out.ctypes.data)
    return out


def ptm_cbcr_avg(info, buffer):
    buffer = np.ascontiguousarray(buffer, dtype=np.uint8)
    out = np.empty((info.pixels, 3), np.uint8)
    lib().ptm_cbcr_avg(C.byref(info), buffer.ctypes.data, out.ctypes.data)
    return out


def ptm_lrgb_finish(info, ycbcr, coeffs):
    """In place: YCbCr average -> RGB, coefficients *= 256 / Y."""
    assert ycbcr.flags.c_contiguous and coeffs.flags.c_contiguous
    lib().ptm_lrgb_finish(C.byref(info), ycbcr.ctypes.data, coeffs.ctypes.data)


def _block_ptrs(blocks):
    return (C.c_void_p * len(blocks))(*[b.ctypes.data for b in blocks])


def ptm_scale_coefficients(header, unscaled, blocks):
    """Fills header.scale / header.bias and the coefficient blocks."""
    unscaled = np.ascontiguousarray(unscaled, dtype=np.float32)
    lib().ptm_scale_coefficients(C.byref(header), unscaled.ctypes.data, _block_ptrs(blocks))


def ptm_encode_stack(header, info, buffer, M, blocks):
    buffer = np.ascontiguousarray(buffer, dtype=np.uint8)
    M = np.ascontiguousarray(M, dtype=np.float32)
    lib().ptm_encode_stack(C.byref(header), C.byref(info), buffer.ctypes.data, M.ctypes.data, _block_ptrs(blocks))


def ptm_relight(header, blocks, u, v):
    out = np.empty((header.dimen[1], header.dimen[0], 3), np.uint8)
    lib().ptm_relight(C.byref(header), _block_ptrs(blocks), u, v, out.ctypes.data)
    return out


def ptm_normal_map(header, blocks):
    """ptm_normal for every pixel of a PTM's first coefficient block -> float32 [pixels][3]."""
    out = np.empty((header.dimen[0] * header.dimen[1], 3), np.float32)
    lib().ptm_normal_map(C.byref(header), _block_ptrs(blocks), out.ctypes.data)
    return out


def ptm_normal(coeffs):
    c = np.ascontiguousarray(coeffs, dtype=np.float32)
    o = [C.c_float() for _ in range(3)]
    lib().ptm_normal(c.ctypes.data, C.byref(o[0]), C.byref(o[1]), C.byref(o[2]))
    return tuple(x.value for x in o)


def stepwise_buffers(width, height):
    """The caller-owned host buffers of the step-by-step sequence (what the reference's main allocates,
    rti-builder/ptm-encoder.c:270,307): float coefficients + the two LRGB blocks, already touched."""
    header = ptm_alloc_header("PTM_FORMAT_LRGB", width, height)
    blocks = ptm_alloc_blocks(header)
    coeffs = np.zeros((width * height, PTM_COEFFICIENTS), np.float32)
    for b in blocks:
        b.fill(0)
    return dict(coeffs=coeffs, blocks=blocks)


def encode_lrgb_stepwise(stack, M, buffers=None):
    """The reference encoder's LRGB call sequence, step by step, on HOST (pageable) numpy buffers
    (rti-builder/ptm-encoder.c:313-353,406):

        ptm_fit_poly_jsample -> ptm_cbcr_avg -> colour + normalise loop (ptm_lrgb_finish) -> ptm_scale_coefficients

    stack: uint8 [n][H][W][3] (YCbCr, stored row order).  `buffers` (stepwise_buffers) are reused if given.
    Returns dict(coef, rgb, scale, bias, coeffs)."""
    assert stack.dtype == np.uint8 and stack.ndim == 4 and stack.shape[3] == 3 and stack.flags.c_contiguous
    n, H, W, _ = stack.shape
    info = image_info(W, H, n)
    header = ptm_alloc_header("PTM_FORMAT_LRGB", W, H)
    buffers = buffers or stepwise_buffers(W, H)
    blocks, coeffs = buffers["blocks"], buffers["coeffs"]
    M = np.ascontiguousarray(M, dtype=np.float32)
    import time
    L = lib()
    t = [time.perf_counter()]
    L.ptm_fit_poly_jsample(C.byref(info), stack.ctypes.data, 3, M.ctypes.data, coeffs.ctypes.data)
    t.append(time.perf_counter())
    L.ptm_cbcr_avg(C.byref(info), stack.ctypes.data, blocks[1].ctypes.data)
    t.append(time.perf_counter())
    L.ptm_lrgb_finish(C.byref(info), blocks[1].ctypes.data, coeffs.ctypes.data)
    t.append(time.perf_counter())
    L.ptm_scale_coefficients(C.byref(header), coeffs.ctypes.data, _block_ptrs(blocks))
    t.append(time.perf_counter())
    names = ("ptm_fit_poly_jsample", "ptm_cbcr_avg", "ptm_lrgb_finish", "ptm_scale_coefficients")
    return dict(coef=blocks[0].reshape(H, W, PTM_COEFFICIENTS), rgb=blocks[1].reshape(H, W, RGB_COEFFICIENTS),
                scale=np.array(header.scale, np.float32), bias=np.array(header.bias, np.int32), coeffs=coeffs,
                ms={n: 1e3 * (b - a) for n, a, b in zip(names, t, t[1:])})
