"""Device-resident PTM pipeline: a thin object layer over the C ABI.

Pointers are plain integers (``tensor.data_ptr()`` of a CUDA tensor, or memory
from ``Context.malloc``).  PyTorch is used by callers for device memory, streams
and the cross-GPU exchange; it is plumbing, not the product.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import NCOEF, NKEYS, U8, U16, U32, PtmError, ScaleBias  # noqa: F401


def _ptr(x):
    """int / None / numpy array / torch tensor -> address."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    raise TypeError("cannot take the address of %r" % type(x))


class Context:
    """One per (process, GPU).  ``stream`` is a raw cudaStream_t handle (int),
    e.g. ``torch.cuda.current_stream().cuda_stream``; None / 0 is the CUDA default stream."""

    def __init__(self, device=0, stream=None):
        self._lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self._lib.ptmb_create(int(device), stream, C.byref(h)))
        self._h = h
        self.device = int(device)
        self.n_lights = 0

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ptmb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------ plumbing
    @property
    def sm_count(self):
        return self._lib.ptmb_sm_count(self._h)

    def set_stream(self, stream):
        _lib.check(self._lib.ptmb_set_stream(self._h, stream))

    def synchronize(self):
        _lib.check(self._lib.ptmb_synchronize(self._h))

    def malloc(self, nbytes):
        p = C.c_void_p()
        _lib.check(self._lib.ptmb_malloc(self._h, nbytes, C.byref(p)))
        return p.value

    def free(self, ptr):
        _lib.check(self._lib.ptmb_free(self._h, ptr))

    def upload(self, dev, host):
        host = np.ascontiguousarray(host)
        _lib.check(self._lib.ptmb_upload(self._h, _ptr(dev), host.ctypes.data, host.nbytes))
        self.synchronize()

    def download(self, dev, shape, dtype):
        out = np.empty(shape, dtype=dtype)
        _lib.check(self._lib.ptmb_download(self._h, out.ctypes.data, _ptr(dev), out.nbytes))
        return out

    # ------------------------------------------------------------- kernels
    def set_matrix(self, M):
        M = np.ascontiguousarray(M, dtype=np.float32)
        assert M.ndim == 2 and M.shape[0] == NCOEF, "M must be 6 x n_lights"
        _lib.check(self._lib.ptmb_set_matrix(self._h, M.ctypes.data, M.shape[1]))
        self.n_lights = M.shape[1]

    def set_fit_engine(self, engine):
        """"default" | "ffma" (FP32 FMA streaming kernels) | "mma" (integer GEMM on the tensor cores)."""
        code = {"default": 0, "ffma": 1, "mma": 2}[engine] if isinstance(engine, str) else int(engine)
        _lib.check(self._lib.ptmb_set_fit_engine(self._h, code))

    def set_coeffs_scratch(self, scratch=True):
        """Declare the float plane of the fused encodes scratch (undefined afterwards): K2 quantises its L2-resident
        end first and discards the lines, so they never travel to HBM and back."""
        _lib.check(self._lib.ptmb_set_coeffs_scratch(self._h, 1 if scratch else 0))

    def set_stack_stable(self, stable=True):
        """Promise that stacks passed to encode_lrgb_dev are not written by earlier work on the stream (lets K1's
        copy warp start before the previous kernel has drained)."""
        _lib.check(self._lib.ptmb_set_stack_stable(self._h, 1 if stable else 0))

    def set_fit_mode(self, lum=False, rgb_input=False):
        """Variants of the LRGB-shaped entries: lum = PTM_FORMAT_LUM (colour block = mean YCbCr, no
        normalisation); rgb_input = the stack holds RGB triplets, converted to YCbCr per sample inside K1."""
        _lib.check(self._lib.ptmb_set_fit_mode(self._h, (1 if lum else 0) | (2 if rgb_input else 0)))

    def set_relight_mode(self, mode):
        """"auto" (exact below 4 directions, else fast) | "exact" (bit-identical to the reference decoder) |
        "fast" (every byte within +-1, HBM-rate sweeps)."""
        code = {"auto": 0, "exact": 1, "fast": 2}[mode] if isinstance(mode, str) else int(mode)
        _lib.check(self._lib.ptmb_set_relight_mode(self._h, code))

    def mma_launches(self):
        return int(self._lib.ptmb_mma_launches())

    def keys_init(self, keys):
        _lib.check(self._lib.ptmb_keys_init(self._h, _ptr(keys)))

    def fit_lrgb(self, stack, image_stride, pixels, coeffs, rgb, keys, sample_type=U8):
        _lib.check(self._lib.ptmb_fit_lrgb(self._h, _ptr(stack), sample_type, image_stride, pixels,
                                           _ptr(coeffs), _ptr(rgb), _ptr(keys)))

    def fit_rgb(self, stack, image_stride, pixels, coeffs, plane_stride, keys, sample_type=U8):
        _lib.check(self._lib.ptmb_fit_rgb(self._h, _ptr(stack), sample_type, image_stride, pixels,
                                          _ptr(coeffs), plane_stride, _ptr(keys)))

    def fit_channel(self, samples, sample_type, pixel_stride, image_stride, pixels, coeffs):
        _lib.check(self._lib.ptmb_fit_channel(self._h, _ptr(samples), sample_type, pixel_stride, image_stride,
                                              pixels, _ptr(coeffs)))

    def chroma_avg(self, stack, image_stride, pixels, n_lights, avg):
        _lib.check(self._lib.ptmb_chroma_avg(self._h, _ptr(stack), image_stride, pixels, n_lights, _ptr(avg)))

    def lrgb_colour_normalise(self, ycc, coeffs, pixels):
        _lib.check(self._lib.ptmb_lrgb_colour_normalise(self._h, _ptr(ycc), _ptr(coeffs), pixels))

    def minmax(self, coeffs, pixels, keys):
        _lib.check(self._lib.ptmb_minmax(self._h, _ptr(coeffs), pixels, _ptr(keys)))

    def quantise(self, coeffs, pixels, keys, out_bytes, scale_bias=None, consume=False):
        """K2.  consume=True: the float plane is scratch and is left undefined (L2 lines are
        discarded instead of written back)."""
        fn = self._lib.ptmb_quantise_consume if consume else self._lib.ptmb_quantise
        _lib.check(fn(self._h, _ptr(coeffs), pixels, _ptr(keys), _ptr(out_bytes), _ptr(scale_bias)))

    def encode_lrgb_dev(self, stack, image_stride, pixels, coeffs, rgb, coef_bytes, scale_bias=None,
                        xchg=None, sample_type=U8):
        """Whole device-resident LRGB encode of one slab (K1 -> key hand-off -> K2, two launches)."""
        _lib.check(self._lib.ptmb_encode_lrgb_dev(self._h, None if xchg is None else xchg._h, _ptr(stack),
                                                  sample_type, image_stride, pixels, _ptr(coeffs), _ptr(rgb),
                                                  _ptr(coef_bytes), _ptr(scale_bias)))

    def encode_rgb_dev(self, stack, image_stride, pixels, coeffs, coef_bytes, scale_bias=None, xchg=None,
                       sample_type=U8):
        """Whole device-resident RGB-format encode of one slab (tensor-core K1 -> key hand-off -> one K2 over the
        three contiguous planes)."""
        _lib.check(self._lib.ptmb_encode_rgb_dev(self._h, None if xchg is None else xchg._h, _ptr(stack), sample_type,
                                                 image_stride, pixels, _ptr(coeffs), _ptr(coef_bytes),
                                                 _ptr(scale_bias)))

    def probe_arm(self, pairs):
        _lib.check(self._lib.ptmb_probe_arm(self._h, pairs))

    def probe_read(self, pairs):
        out = (C.c_float * pairs)()
        n = C.c_uint(0)
        _lib.check(self._lib.ptmb_probe_read(self._h, out, pairs, C.byref(n)))
        return [out[i] for i in range(n.value)]

    def relight_lrgb(self, coef, rgb, width, height, scale, bias, uv, out):
        scale = np.ascontiguousarray(scale, dtype=np.float32)
        bias = np.ascontiguousarray(bias, dtype=np.int32)
        uv = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 2)
        _lib.check(self._lib.ptmb_relight_lrgb(self._h, _ptr(coef), _ptr(rgb), width, height, scale.ctypes.data,
                                               bias.ctypes.data, uv.ctypes.data, uv.shape[0], _ptr(out)))

    def relight_rgb(self, r, g, b, width, height, scale, bias, uv, out):
        scale = np.ascontiguousarray(scale, dtype=np.float32)
        bias = np.ascontiguousarray(bias, dtype=np.int32)
        uv = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 2)
        _lib.check(self._lib.ptmb_relight_rgb(self._h, _ptr(r), _ptr(g), _ptr(b), width, height,
                                              scale.ctypes.data, bias.ctypes.data, uv.ctypes.data, uv.shape[0],
                                              _ptr(out)))

    def relight_lum(self, coef, crcb, width, height, scale, bias, uv, out):
        """PTM_FORMAT_LUM: (Y, Cb, Cr) rasters; crcb = 2-byte (cr, cb) records."""
        scale = np.ascontiguousarray(scale, dtype=np.float32)
        bias = np.ascontiguousarray(bias, dtype=np.int32)
        uv = np.ascontiguousarray(uv, dtype=np.float32).reshape(-1, 2)
        _lib.check(self._lib.ptmb_relight_lum(self._h, _ptr(coef), _ptr(crcb), width, height, scale.ctypes.data,
                                              bias.ctypes.data, uv.ctypes.data, uv.shape[0], _ptr(out)))

    def normal_map(self, coeffs, pixels, normals):
        """float coefficient plane [pixels][6] -> normals float [pixels][3]."""
        _lib.check(self._lib.ptmb_normal_map(self._h, _ptr(coeffs), pixels, _ptr(normals)))

    def normal_map_blocks(self, coef_bytes, scale, bias, pixels, normals):
        scale = np.ascontiguousarray(scale, dtype=np.float32)
        bias = np.ascontiguousarray(bias, dtype=np.int32)
        _lib.check(self._lib.ptmb_normal_map_blocks(self._h, _ptr(coef_bytes), scale.ctypes.data, bias.ctypes.data,
                                                    pixels, _ptr(normals)))

    def fit_lsum(self, stack, image_stride, pixels, coeffs, colour, keys, median=False):
        """The reference encoder's disabled L = R + G + B branch; median=True: median chroma instead of the mean."""
        _lib.check(self._lib.ptmb_fit_lsum(self._h, _ptr(stack), image_stride, pixels, 1 if median else 0,
                                           _ptr(coeffs), _ptr(colour), _ptr(keys)))

    def jpeg_info(self, data):
        w, h, c = C.c_uint(0), C.c_uint(0), C.c_uint(0)
        _lib.check(self._lib.ptmb_jpeg_info(self._h, data, len(data), C.byref(w), C.byref(h), C.byref(c)))
        return w.value, h.value, c.value

    def jpeg_decode(self, data, want_rgb=False, flip=False, out=None):
        """JPEG bytes -> uint8 [h][w][c] (numpy), or into `out` (a CUDA tensor / device address)."""
        w, h, c = self.jpeg_info(data)
        if out is None:
            res = np.empty((h, w, c), np.uint8)
            _lib.check(self._lib.ptmb_jpeg_decode(self._h, data, len(data), int(want_rgb), int(flip),
                                                  res.ctypes.data, 0))
            return res
        _lib.check(self._lib.ptmb_jpeg_decode(self._h, data, len(data), int(want_rgb), int(flip), _ptr(out), 1))
        return out

    def jpeg_encode(self, image, ycbcr=False, quality=90):
        """uint8 [h][w] or [h][w][3] (numpy) -> JPEG bytes."""
        img = np.ascontiguousarray(image, dtype=np.uint8)
        h, w = img.shape[:2]
        comps = 1 if img.ndim == 2 else img.shape[2]
        p, n = C.c_void_p(), C.c_size_t(0)
        _lib.check(self._lib.ptmb_jpeg_encode(self._h, img.ctypes.data, w, h, comps, int(ycbcr), quality,
                                              C.byref(p), C.byref(n)))
        data = C.string_at(p, n.value)
        C.CDLL(None).free(p)
        return data

    def selftest_div255(self, first_bits=0, count=1 << 32):
        """Mismatch count of K3's division / CLIP shortcuts against IEEE / the reference."""
        n = C.c_uint64(0)
        _lib.check(self._lib.ptmb_selftest_div255(self._h, first_bits, count, C.byref(n)))
        return n.value

    def synth(self, seed, mode, sample_type, pixel0, pixels, lights_q14, stack, image_stride):
        lq = np.ascontiguousarray(lights_q14, dtype=np.int32).reshape(-1, 3)
        _lib.check(self._lib.ptmb_synth(self._h, seed, mode, sample_type, pixel0, pixels, lq.shape[0],
                                        lq.ctypes.data, _ptr(stack), image_stride))

    def encode_host_fit(self, stack, shape, sample_type, M, planes=1, keys_out=None, early_rgb=None):
        """First half of the host-buffer encode: stream the HOST stack (address or numpy array,
        [n][H][W][3]) through K1.  Leaves the extrema keys on the device (copied to
        ``keys_out`` if given) so that ranks can exchange them before encode_host_finish.
        ``early_rgb`` (LRGB): host destination of the RGB block, filled slab by slab during the fit;
        pass the same buffer as out[1] of encode_host_finish."""
        n, H, W = shape
        if early_rgb is not None:
            _lib.check(self._lib.ptmb_encode_host_early_rgb(self._h, _ptr(early_rgb)))
        M = np.ascontiguousarray(M, dtype=np.float32)
        self._enc_geom = (H, W, planes)
        _lib.check(self._lib.ptmb_encode_host_fit(self._h, _ptr(stack), sample_type, W, H, n, M.ctypes.data,
                                                  planes, _ptr(keys_out)))

    def encode_host_finish(self, keys=None, out=None, want_coeffs=False):
        """Second half: K2 with ``keys`` (None = the context's own) and D2H of the blocks into
        ``out`` (host arrays / pinned tensors; allocated if None).  Synchronous."""
        H, W, planes = self._enc_geom
        P = H * W
        if out is not None:
            blocks = out
        elif planes == 1:
            blocks = [np.empty((H, W, NCOEF), np.uint8), np.empty((H, W, 3), np.uint8)]
        else:
            blocks = [np.empty((H, W, NCOEF), np.uint8) for _ in range(3)]
        coeffs = np.empty((planes, P, NCOEF), np.float32) if want_coeffs else None
        sb = ScaleBias()
        ptrs = (C.c_void_p * len(blocks))(*[_ptr(b) for b in blocks])
        _lib.check(self._lib.ptmb_encode_host_finish(self._h, _ptr(keys), ptrs,
                                                     None if coeffs is None else coeffs.ctypes.data,
                                                     C.addressof(sb)))
        return dict(blocks=blocks, scale=np.array(sb.scale, dtype=np.float32),
                    bias=np.array(sb.bias, dtype=np.int32), coeffs=coeffs)

    def encode_host(self, stack, M, planes=1, want_coeffs=False, out=None):
        """Whole encode from a HOST numpy stack [n][H][W][3] (uint8, or uint16 for planes=3).
        Returns dict(blocks, scale, bias, coeffs)."""
        assert stack.ndim == 4 and stack.shape[3] == 3 and stack.flags.c_contiguous
        n, H, W, _ = stack.shape
        st = {np.dtype(np.uint8): U8, np.dtype(np.uint16): U16}[stack.dtype]
        self.encode_host_fit(stack, (n, H, W), st, M, planes)
        return self.encode_host_finish(None, out, want_coeffs)

    def encode_host_release(self):
        _lib.check(self._lib.ptmb_encode_host_release(self._h))


class KeyExchange:
    """NVLink peer-memory merge of the extrema keys (ptmb_xchg_* in include/ptm_b200.h).

    Two-step set-up so that every rank can take part in the out-of-band handle swap even if
    its own step failed: ``create`` -> handle bytes (or raises), the host all-gathers the
    handles with any transport, ``connect(list_of_handles)`` maps the peers."""

    HANDLE_BYTES = 64

    def __init__(self, ctx, rank, world):
        self._lib = ctx._lib
        self._ctx = ctx
        self.world = world
        self._h = C.c_void_p()
        handle = (C.c_char * self.HANDLE_BYTES)()
        _lib.check(self._lib.ptmb_xchg_create(ctx._h, rank, world, C.byref(self._h), handle))
        self.handle = bytes(handle)

    def connect(self, handles):
        assert len(handles) == self.world and all(len(h) == self.HANDLE_BYTES for h in handles)
        _lib.check(self._lib.ptmb_xchg_connect(self._h, b"".join(handles)))

    def allreduce_max(self, keys):
        _lib.check(self._lib.ptmb_xchg_allreduce_max(self._ctx._h, self._h, _ptr(keys)))

    def status(self):
        st = C.c_int(0)
        _lib.check(self._lib.ptmb_xchg_status(self._ctx._h, self._h, C.byref(st)))
        return st.value

    def close(self):
        if getattr(self, "_h", None):
            self._lib.ptmb_xchg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def light_q14(lights):
    """Q14 fixed-point light directions for the synthetic generator (round half away)."""
    l = np.asarray(lights, dtype=np.float32).reshape(-1, 3) * np.float32(16384.0)
    return np.where(l >= 0, l + np.float32(0.5), l - np.float32(0.5)).astype(np.int32)


def key_to_float(keys):
    """Decode order-preserving int32 keys (numpy) back to floats."""
    k = np.asarray(keys, dtype=np.int32)
    b = k ^ ((k >> 31) & 0x7FFFFFFF)
    return b.view(np.float32)
