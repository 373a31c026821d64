"""ctypes front end for the CPU oracle (oracle/ptm_oracle.c, oracle/synth.c)
and, when it has been built, for the reference's own ptmlib.c
(oracle/_ref/libptmref.so) and command-line mains (oracle/_ref/ptm-*-ref).

TEST INFRASTRUCTURE ONLY -- see the header of ptm_oracle.c.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
NCOEF = 6

_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_u16p = np.ctypeslib.ndpointer(np.uint16, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")

_lib = None


def build(quiet=True):
    """(Re)build liboracle.so and, where /root/reference exists, oracle/_ref."""
    out = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)


def lib():
    global _lib
    if _lib is not None:
        return _lib
    path = os.path.join(HERE, "liboracle.so")
    if not os.path.exists(path):
        build()
    L = C.CDLL(path)
    L.orc_pinv.argtypes = [_f32p, _f32p, C.c_int, _f32p]
    L.orc_pinv.restype = C.c_int
    L.orc_pinv_from_design.argtypes = [_f32p, C.c_int, _f32p]
    L.orc_pinv_from_design.restype = C.c_int
    L.orc_singular_values.argtypes = [_f32p, C.c_int, _f32p]
    L.orc_singular_values.restype = C.c_int
    for name in ("orc_fit_u8", "orc_fit_u8_blas"):
        getattr(L, name).argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint, C.c_size_t, _f32p, _f32p]
        getattr(L, name).restype = None
    L.orc_fit_u32.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint, C.c_size_t, _f32p, _f32p]
    L.orc_fit_u32.restype = None
    L.orc_chroma_avg.argtypes = [_u8p, C.c_size_t, C.c_uint, _u8p]
    L.orc_chroma_avg.restype = None
    L.orc_lrgb_colour_normalise.argtypes = [_u8p, _f32p, C.c_size_t]
    L.orc_lrgb_colour_normalise.restype = None
    L.orc_scale.argtypes = [_f32p, C.c_size_t, C.c_int, _f32p, _i32p, C.POINTER(C.c_void_p), C.c_void_p]
    L.orc_scale.restype = None
    L.orc_relight_lrgb.argtypes = [_u8p, _u8p, C.c_size_t, C.c_size_t, _f32p, _i32p, C.c_float, C.c_float, _u8p]
    L.orc_relight_lrgb.restype = None
    L.orc_relight_rgb.argtypes = [_u8p, _u8p, _u8p, C.c_size_t, C.c_size_t, _f32p, _i32p, C.c_float, C.c_float, _u8p]
    L.orc_relight_rgb.restype = None
    L.orc_lrgb16_finish.argtypes = [_u16p, C.c_size_t, C.c_uint, _f32p, _u8p]
    L.orc_lrgb16_finish.restype = None
    L.orc_normal.argtypes = [_f32p, _f32p, _f32p, _f32p]
    L.orc_normal.restype = None
    L.orc_normal_map.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, _f32p]
    L.orc_normal_map.restype = None
    L.orc_relight_lum.argtypes = [_u8p, _u8p, C.c_size_t, C.c_size_t, _f32p, _i32p, C.c_float, C.c_float, _u8p]
    L.orc_relight_lum.restype = None
    L.orc_rgb_to_ycbcr.argtypes = [_u8p, C.c_size_t]
    L.orc_rgb_to_ycbcr.restype = None
    L.orc_fit_lsum.argtypes = [_u8p, C.c_size_t, C.c_uint, _f32p, C.c_int, _f32p, _u8p]
    L.orc_fit_lsum.restype = None
    L.orc_blas_corename.restype = C.c_char_p
    L.orc_blas_set_threads.argtypes = [C.c_int]
    L.orc_omp_set_threads.argtypes = [C.c_int]
    L.orc_omp_max_threads.restype = C.c_int
    L.orc_omp_team_size.restype = C.c_int
    L.synth_light_q14.argtypes = [_f32p, C.c_int, _i32p]
    L.synth_stack_u8.argtypes = [C.c_uint32, C.c_int, C.c_uint32, C.c_size_t, C.c_uint32, _i32p, _u8p]
    L.synth_stack_u16.argtypes = [C.c_uint32, C.c_int, C.c_uint32, C.c_size_t, C.c_uint32, _i32p, _u16p]
    _lib = L
    return L


# ------------------------------------------------------------------ helpers

def host_cores():
    """CPUs this process may run on."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def set_threads(n=None):
    """OpenMP team size of the oracle AND of the reference's libptmref.so (one libgomp per process), set
    through the library: launchers like torchrun export OMP_NUM_THREADS=1, and an environment variable
    changed after libgomp is loaded is ignored.  Returns the team size a parallel region really gets."""
    lib().orc_omp_set_threads(int(n or host_cores()))
    return int(lib().orc_omp_team_size())


def pinv(lights):
    """6 x n pseudo-inverse of the light matrix (LAPACKE path of the reference)."""
    lights = np.ascontiguousarray(lights, dtype=np.float32)
    n = lights.shape[0]
    u = np.ascontiguousarray(lights[:, 0])
    v = np.ascontiguousarray(lights[:, 1])
    M = np.zeros((NCOEF, n), dtype=np.float32)
    if lib().orc_pinv(u, v, n, M):
        raise RuntimeError("SVD did not converge")
    return M


def pinv_from_design(A):
    A = np.ascontiguousarray(A, dtype=np.float32)
    M = np.zeros((NCOEF, A.shape[0]), dtype=np.float32)
    if lib().orc_pinv_from_design(A, A.shape[0], M):
        raise RuntimeError("SVD did not converge")
    return M


def singular_values(A):
    A = np.ascontiguousarray(A, dtype=np.float32)
    sv = np.zeros(NCOEF, dtype=np.float32)
    lib().orc_singular_values(A, A.shape[0], sv)
    return sv


def pinv_f64(lights):
    """Second opinion: float64 numpy pseudo-inverse."""
    l = np.asarray(lights, dtype=np.float64)
    u, v = l[:, 0], l[:, 1]
    A = np.stack([u * u, v * v, u * v, u, v, np.ones_like(u)], axis=1)
    return np.linalg.pinv(A)


def light_q14(lights):
    lights = np.ascontiguousarray(lights, dtype=np.float32).reshape(-1, 3)
    q = np.zeros(lights.shape, dtype=np.int32)
    lib().synth_light_q14(lights.reshape(-1), lights.shape[0], q.reshape(-1))
    return q


def synth_stack(seed, mode, width, height, lights, dtype=np.uint8, row0=0, rows=None):
    """Synthetic stack [n][rows][width][3] for rows [row0, row0+rows) of a
    width x height image (same bytes as the device generator)."""
    rows = height - row0 if rows is None else rows
    q = light_q14(lights)
    n = q.shape[0]
    out = np.zeros((n, rows, width, 3), dtype=dtype)
    fn = lib().synth_stack_u8 if dtype == np.uint8 else lib().synth_stack_u16
    fn(seed, mode, row0 * width, rows * width, n, q.reshape(-1), out.reshape(-1))
    return out


def fit_u8(stack, M, channel=0, blas=False):
    """stack uint8 [n][H][W][3] -> float32 [H][W][6] for one channel."""
    n, H, W, s = stack.shape
    out = np.zeros((H, W, NCOEF), dtype=np.float32)
    M = np.ascontiguousarray(M, dtype=np.float32)
    fn = lib().orc_fit_u8_blas if blas else lib().orc_fit_u8
    fn(stack.ctypes.data + channel, W, H, n, s, M.reshape(-1), out.reshape(-1))
    return out


def fit_u32(samples, M, channel=0):
    """samples uint32 [n][H][W][s] -> float32 [H][W][6]."""
    n, H, W, s = samples.shape
    out = np.zeros((H, W, NCOEF), dtype=np.float32)
    M = np.ascontiguousarray(M, dtype=np.float32)
    lib().orc_fit_u32(samples.ctypes.data + 4 * channel, W, H, n, s, M.reshape(-1), out.reshape(-1))
    return out


def chroma_avg(stack):
    n, H, W, _ = stack.shape
    out = np.zeros((H, W, 3), dtype=np.uint8)
    lib().orc_chroma_avg(stack.reshape(-1), H * W, n, out.reshape(-1))
    return out


def lrgb_colour_normalise(ycc, coeffs):
    """In place, like the reference's loop; returns (rgb, coeffs)."""
    lib().orc_lrgb_colour_normalise(ycc.reshape(-1), coeffs.reshape(-1), coeffs.size // NCOEF)
    return ycc, coeffs


def scale(coeffs, planes=1):
    """coeffs float32 [planes][P][6] (or [P][6]) -> (blocks uint8 [planes][P][6],
    scale[6], bias[6], minmax[12])."""
    c = np.ascontiguousarray(coeffs, dtype=np.float32).reshape(planes, -1, NCOEF)
    P = c.shape[1]
    blocks = np.zeros((planes, P, NCOEF), dtype=np.uint8)
    sc = np.zeros(NCOEF, dtype=np.float32)
    bi = np.zeros(NCOEF, dtype=np.int32)
    mm = np.zeros(2 * NCOEF, dtype=np.float32)
    ptrs = (C.c_void_p * planes)(*[blocks[i].ctypes.data for i in range(planes)])
    lib().orc_scale(c.reshape(-1), P, planes, sc, bi, ptrs, mm.ctypes.data)
    return blocks, sc, bi, mm


def encode_lrgb(stack, M, blas=False):
    """Full LRGB sequence.  Returns dict(coeffs, coef, rgb, scale, bias, minmax)."""
    n, H, W, _ = stack.shape
    coeffs = fit_u8(stack, M, 0, blas=blas)
    rgb = chroma_avg(stack)
    lrgb_colour_normalise(rgb, coeffs)
    blocks, sc, bi, mm = scale(coeffs.reshape(-1, NCOEF), 1)
    return dict(coeffs=coeffs, coef=blocks[0].reshape(H, W, NCOEF), rgb=rgb, scale=sc, bias=bi, minmax=mm)


def encode_lrgb_u16(stack, M):
    """EXTENSION (no reference counterpart, see orc_lrgb16_finish): 16-bit LRGB encode."""
    n, H, W, _ = stack.shape
    coeffs = fit_u32(stack.astype(np.uint32), M, 0)
    rgb = np.zeros((H, W, 3), np.uint8)
    lib().orc_lrgb16_finish(np.ascontiguousarray(stack).reshape(-1), H * W, n, coeffs.reshape(-1), rgb.reshape(-1))
    blocks, sc, bi, mm = scale(coeffs.reshape(-1, NCOEF), 1)
    return dict(coeffs=coeffs, coef=blocks[0].reshape(H, W, NCOEF), rgb=rgb, scale=sc, bias=bi, minmax=mm)


def encode_rgb_u16(stack, M):
    """RGB-format encode of a 16-bit stack through the reference's > 8 bit entry
    (ptm_fit_poly_uint semantics, rti-builder/ptmlib.c:727-760) and ptm_scale_coefficients."""
    n, H, W, _ = stack.shape
    s32 = stack.astype(np.uint32)
    coeffs = np.stack([fit_u32(s32, M, ch) for ch in range(3)], axis=0)
    blocks, sc, bi, mm = scale(coeffs.reshape(3, -1, NCOEF), 3)
    return dict(coeffs=coeffs, coef=blocks.reshape(3, H, W, NCOEF), scale=sc, bias=bi, minmax=mm)


def encode_rgb(stack, M, blas=False):
    n, H, W, _ = stack.shape
    coeffs = np.stack([fit_u8(stack, M, ch, blas=blas) for ch in range(3)], axis=0)
    blocks, sc, bi, mm = scale(coeffs.reshape(3, -1, NCOEF), 3)
    return dict(coeffs=coeffs, coef=blocks.reshape(3, H, W, NCOEF), scale=sc, bias=bi, minmax=mm)


def header_roundtrip(scale_vals):
    """What a decoder sees after the scale went through the PTM text header:
    written with %f, parsed back (rti-builder/ptmlib.c:156-163,132-136)."""
    return np.asarray([np.float32(float("%f" % float(s))) for s in scale_vals], dtype=np.float32)


def relight_lrgb(coef, rgb, scale_vals, bias, u, v):
    H, W, _ = coef.shape
    out = np.zeros((H, W, 3), dtype=np.uint8)
    lib().orc_relight_lrgb(np.ascontiguousarray(coef).reshape(-1), np.ascontiguousarray(rgb).reshape(-1), W, H,
                           np.ascontiguousarray(scale_vals, dtype=np.float32),
                           np.ascontiguousarray(bias, dtype=np.int32), u, v, out.reshape(-1))
    return out


def relight_rgb(coef3, scale_vals, bias, u, v):
    _, H, W, _ = coef3.shape
    out = np.zeros((H, W, 3), dtype=np.uint8)
    c = np.ascontiguousarray(coef3)
    lib().orc_relight_rgb(c[0].reshape(-1), c[1].reshape(-1), c[2].reshape(-1), W, H,
                          np.ascontiguousarray(scale_vals, dtype=np.float32),
                          np.ascontiguousarray(bias, dtype=np.int32), u, v, out.reshape(-1))
    return out


def normal_map(coeffs=None, coef_bytes=None, scale_vals=None, bias=None):
    """ptm_normal for every pixel -> float32 [P][3]; from float coefficients [P][6] or from a quantised block."""
    if coeffs is not None:
        c = np.ascontiguousarray(coeffs, dtype=np.float32).reshape(-1, NCOEF)
        out = np.zeros((c.shape[0], 3), np.float32)
        lib().orc_normal_map(c.ctypes.data, None, None, None, c.shape[0], out.reshape(-1))
        return out
    b = np.ascontiguousarray(coef_bytes, dtype=np.uint8).reshape(-1, NCOEF)
    sc = np.ascontiguousarray(scale_vals, dtype=np.float32)
    bi = np.ascontiguousarray(bias, dtype=np.int32)
    out = np.zeros((b.shape[0], 3), np.float32)
    lib().orc_normal_map(None, b.ctypes.data, sc.ctypes.data, bi.ctypes.data, b.shape[0], out.reshape(-1))
    return out


def relight_lum(coef, crcb, scale_vals, bias, u, v):
    """PTM_FORMAT_LUM relight (rti-builder/ptmlib.c:601-609): crcb = flat uint8 buffer read as 2-byte (cr, cb)
    records (at least 2 bytes per pixel).  Returns the (Y, Cb, Cr) raster."""
    H, W, _ = coef.shape
    out = np.zeros((H, W, 3), dtype=np.uint8)
    lib().orc_relight_lum(np.ascontiguousarray(coef).reshape(-1), np.ascontiguousarray(crcb).reshape(-1), W, H,
                          np.ascontiguousarray(scale_vals, dtype=np.float32),
                          np.ascontiguousarray(bias, dtype=np.int32), u, v, out.reshape(-1))
    return out


def rgb_to_ycbcr(stack):
    """Per-sample integer JFIF conversion of an RGB stack (copy)."""
    out = np.ascontiguousarray(stack, dtype=np.uint8).copy()
    lib().orc_rgb_to_ycbcr(out.reshape(-1), out.size // 3)
    return out


def encode_lum(stack, M):
    """PTM_FORMAT_LUM sequence (rti-builder/ptm-encoder.c:286-305,406): fit Y, average the triplets, scale --
    no colour conversion, no normalisation.  Returns dict(coeffs, coef, ycc, scale, bias)."""
    n, H, W, _ = stack.shape
    coeffs = fit_u8(stack, M, 0)
    ycc = chroma_avg(stack)
    blocks, sc, bi, mm = scale(coeffs.reshape(-1, NCOEF), 1)
    return dict(coeffs=coeffs, coef=blocks[0].reshape(H, W, NCOEF), ycc=ycc, scale=sc, bias=bi, minmax=mm)


def fit_lsum(stack, M, median=False):
    """The reference encoder's disabled L = R + G + B branch (rti-builder/ptm-encoder.c:356-400) + scale.
    Returns dict(coeffs, coef, colour, scale, bias)."""
    n, H, W, _ = stack.shape
    M = np.ascontiguousarray(M, dtype=np.float32)
    coeffs = np.zeros((H, W, NCOEF), np.float32)
    colour = np.zeros((H, W, 3), np.uint8)
    lib().orc_fit_lsum(np.ascontiguousarray(stack).reshape(-1), H * W, n, M.reshape(-1), 1 if median else 0,
                       coeffs.reshape(-1), colour.reshape(-1))
    blocks, sc, bi, mm = scale(coeffs.reshape(-1, NCOEF), 1)
    return dict(coeffs=coeffs, coef=blocks[0].reshape(H, W, NCOEF), colour=colour, scale=sc, bias=bi, minmax=mm)


def normal(c):
    c = np.ascontiguousarray(c, dtype=np.float32)
    o = [np.zeros(1, dtype=np.float32) for _ in range(3)]
    lib().orc_normal(c, *o)
    return tuple(float(x[0]) for x in o)


# ------------------------------------------------ the compiled reference itself

def ref_available():
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in
               ("libptmref.so", "ptm-encoder-ref", "ptm-decoder-ref", "ptm-exploder-ref"))


def write_rawj(path, img):
    """Write an image [H][W][C] uint8 as a RAWJ stream (see shim/rawjpeg.c)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W, Cn = img.shape
    with open(path, "wb") as fp:
        fp.write(b"RAWJ" + struct.pack("<III", W, H, Cn))
        fp.write(img.tobytes())


def read_rawj(data):
    assert data[:4] == b"RAWJ", "not a RAWJ stream"
    W, H, Cn = struct.unpack("<III", data[4:16])
    return np.frombuffer(data, dtype=np.uint8, count=W * H * Cn, offset=16).reshape(H, W, Cn)


def parse_ptm(data):
    """Parse an uncompressed PTM_1.2 file (layout written by
    rti-builder/ptmlib.c:218-239,333-356).  Returns dict with format, width,
    height, scale (float32, as parsed from text), bias, blocks (list)."""
    pos = 0

    def line():
        nonlocal pos
        end = data.index(b"\n", pos)
        s = data[pos:end].decode("ascii").strip()
        pos = end + 1
        return s

    assert line() == "PTM_1.2"
    fmt = line()
    W = int(line())
    H = int(line())
    sc = np.asarray([np.float32(float(x)) for x in line().split()], dtype=np.float32)
    bi = np.asarray([int(x) for x in line().split()], dtype=np.int32)
    P = W * H
    blocks = []
    if fmt == "PTM_FORMAT_RGB":
        sizes = [6, 6, 6]
    elif fmt == "PTM_FORMAT_LRGB":
        sizes = [6, 3]
    else:
        raise ValueError("parse_ptm handles the uncompressed RGB/LRGB formats only: " + fmt)
    for s in sizes:
        blocks.append(np.frombuffer(data, dtype=np.uint8, count=P * s, offset=pos).reshape(H, W, s))
        pos += P * s
    return dict(format=fmt, width=W, height=H, scale=sc, bias=bi, blocks=blocks, header_len=pos - sum(sizes) * P)


def ref_env(threads=None):
    env = dict(os.environ)
    env["OPENBLAS_NUM_THREADS"] = "1"
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
    return env


def ref_encode(stack_stored, lights, fmt, workdir, threads=None):
    """Run the reference's unmodified ptm-encoder on a stack.

    ``stack_stored`` is in STORED row order (what the fit sees).  The encoder
    flips each decoded image vertically (rti-builder/ptm-encoder.c:252-256), so
    the files are written flipped to land in that order.  Returns parse_ptm()."""
    n = stack_stored.shape[0]
    names = []
    for i in range(n):
        name = "img%03d.rawj" % i
        write_rawj(os.path.join(workdir, name), stack_stored[i, ::-1])
        names.append(name)
    lp = os.path.join(workdir, "sample.lp")
    with open(lp, "w") as fp:
        fp.write("%d\n" % n)
        for name, (u, v, w) in zip(names, np.asarray(lights, dtype=np.float32)):
            fp.write("%s %.9g %.9g %.9g\n" % (name, u, v, w))
    out = os.path.join(workdir, "out.ptm")
    subprocess.run([os.path.join(REF_DIR, "ptm-encoder-ref"), "-f", fmt, "-o", out, lp],
                   check=True, env=ref_env(threads))
    with open(out, "rb") as fp:
        return parse_ptm(fp.read()), out


def ref_decode(ptm_path, u, v):
    """Run the reference's unmodified ptm-decoder; returns the RGB raster."""
    res = subprocess.run([os.path.join(REF_DIR, "ptm-decoder-ref"), ptm_path, repr(float(u)), repr(float(v))],
                         check=True, capture_output=True, env=ref_env())
    return read_rawj(res.stdout)


_reflib = None


class _RefInfo(C.Structure):
    # rti-builder/ptmlib.h:105-115
    _fields_ = [("width", C.c_size_t), ("height", C.c_size_t), ("pixels", C.c_size_t),
                ("row_stride", C.c_size_t), ("decoder_stride", C.c_size_t), ("n_decoders", C.c_uint)]


def reflib():
    global _reflib
    if _reflib is None:
        L = C.CDLL(os.path.join(REF_DIR, "libptmref.so"))
        L.ptm_fit_poly_jsample.argtypes = [C.POINTER(_RefInfo), C.c_void_p, C.c_size_t, _f32p, _f32p]
        L.ptm_fit_poly_jsample.restype = None
        L.ptm_cbcr_avg.argtypes = [C.POINTER(_RefInfo), _u8p, _u8p]
        L.ptm_cbcr_avg.restype = None
        _reflib = L
    return _reflib


def ref_info(n, H, W):
    return _RefInfo(W, H, W * H, W * 3, W * H * 3, n)


def ref_fit_u8(stack, M, channel=0):
    """The reference's own ptm_fit_poly_jsample (rti-builder/ptmlib.c:692)."""
    n, H, W, s = stack.shape
    out = np.zeros((H, W, NCOEF), dtype=np.float32)
    info = ref_info(n, H, W)
    reflib().ptm_fit_poly_jsample(C.byref(info), stack.ctypes.data + channel, s,
                                  np.ascontiguousarray(M, dtype=np.float32).reshape(-1), out.reshape(-1))
    return out


def ref_cbcr_avg(stack):
    n, H, W, _ = stack.shape
    out = np.zeros((H, W, 3), dtype=np.uint8)
    info = ref_info(n, H, W)
    reflib().ptm_cbcr_avg(C.byref(info), stack.reshape(-1), out.reshape(-1))
    return out


# ----------------------------------------------------- timed CPU baseline (bench.py)

class _RefFormat(C.Structure):
    # rti-builder/ptmlib.h:87-99
    _fields_ = [("id", C.c_int), ("blocks", C.c_int), ("ptm_blocks", C.c_int), ("color_components", C.c_int),
                ("jpeg_streams", C.c_int), ("name", C.c_char_p)]


class _RefHeader(C.Structure):
    # rti-builder/ptmlib.h:123-138 (MAX_JPEG_STREAMS expands to 3 * 6 = 18 inside the declarators)
    _fields_ = [("format", C.POINTER(_RefFormat)), ("dimen", C.c_size_t * 2), ("scale", C.c_float * 6),
                ("bias", C.c_int * 6), ("compression_param", C.c_int * 1), ("transforms", C.c_int * 18),
                ("motion_vector_x", C.c_int * 18), ("motion_vector_y", C.c_int * 18), ("order", C.c_int * 18),
                ("reference_planes", C.c_int * 18), ("compressed_size", C.c_size_t * 18),
                ("side_info_sizes", C.c_size_t * 18)]


def cpu_encode_lrgb_timed(stack, M, use_ref=None, threads=None):
    """Time the reference's LRGB sequence on the host (rti-builder/ptm-encoder.c:313-353,406):
    ptm_fit_poly_jsample -> ptm_cbcr_avg -> colour/normalise loop -> ptm_scale_coefficients.

    With oracle/_ref built the three ptmlib.c functions are the reference's OWN compiled code
    (kind "reference"); the colour loop lives in the encoder's main and is always the
    restatement.  Otherwise everything is the port (kind "port").  Returns
    (seconds, kind, result dict)."""
    import time
    if use_ref is None:
        use_ref = os.path.exists(os.path.join(REF_DIR, "libptmref.so"))
    if threads:
        set_threads(threads)
    n, H, W, _ = stack.shape
    P = H * W
    M = np.ascontiguousarray(M, dtype=np.float32)
    coeffs = np.empty((P, NCOEF), np.float32)
    rgb = np.empty((P, 3), np.uint8)
    coef = np.empty((P, NCOEF), np.uint8)
    lib().orc_blas_set_threads(1)
    if use_ref:
        R = reflib()
        R.ptm_get_format.restype = C.POINTER(_RefFormat)
        R.ptm_get_format.argtypes = [C.c_char_p]
        R.ptm_scale_coefficients.argtypes = [C.POINTER(_RefHeader), _f32p, C.POINTER(C.c_void_p)]
        R.ptm_scale_coefficients.restype = None
        hdr = _RefHeader()
        hdr.format = R.ptm_get_format(b"PTM_FORMAT_LRGB")
        hdr.dimen[0], hdr.dimen[1] = W, H
        info = ref_info(n, H, W)
        blocks = (C.c_void_p * 2)(coef.ctypes.data, rgb.ctypes.data)
        t0 = time.perf_counter()
        R.ptm_fit_poly_jsample(C.byref(info), stack.ctypes.data, 3, M.reshape(-1), coeffs.reshape(-1))
        R.ptm_cbcr_avg(C.byref(info), stack.reshape(-1), rgb.reshape(-1))
        lib().orc_lrgb_colour_normalise(rgb.reshape(-1), coeffs.reshape(-1), P)
        R.ptm_scale_coefficients(C.byref(hdr), coeffs.reshape(-1), blocks)
        dt = time.perf_counter() - t0
        sc = np.array(hdr.scale, np.float32)
        bi = np.array(hdr.bias, np.int32)
        kind = "reference"
    else:
        sc = np.zeros(NCOEF, np.float32)
        bi = np.zeros(NCOEF, np.int32)
        ptrs = (C.c_void_p * 1)(coef.ctypes.data)
        t0 = time.perf_counter()
        lib().orc_fit_u8(stack.ctypes.data, W, H, n, 3, M.reshape(-1), coeffs.reshape(-1))
        lib().orc_chroma_avg(stack.reshape(-1), P, n, rgb.reshape(-1))
        lib().orc_lrgb_colour_normalise(rgb.reshape(-1), coeffs.reshape(-1), P)
        lib().orc_scale(coeffs.reshape(-1), P, 1, sc, bi, ptrs, None)
        dt = time.perf_counter() - t0
        kind = "port"
    return dt, kind, dict(coeffs=coeffs.reshape(H, W, NCOEF), coef=coef.reshape(H, W, NCOEF),
                          rgb=rgb.reshape(H, W, 3), scale=sc, bias=bi)
