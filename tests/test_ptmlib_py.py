"""GPU tests of the ptmlib.h-shaped host API (rti_b200/ptmlib.py -> libptm.so): the reference's
own call sequence, step by step, with host buffers."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_reference_call_sequence_lrgb(oracle, dome1):
    from rti_b200 import ptmlib as P
    W, H = 90, 41
    stack = oracle.synth_stack(0xAB1, 0, W, H, dome1)
    M = P.ptm_svd(dome1)
    np.testing.assert_array_equal(M, oracle.pinv(dome1))            # same LAPACKE path on the host
    info = P.image_info(W, H, 60)
    hdr = P.ptm_alloc_header("PTM_FORMAT_LRGB", W, H)
    blocks = P.ptm_alloc_blocks(hdr)
    # rti-builder/ptm-encoder.c:313-353,406
    coeffs = P.ptm_fit_poly_jsample(info, stack, 3, M)
    ycc = P.ptm_cbcr_avg(info, stack)
    np.testing.assert_array_equal(ycc.reshape(H, W, 3), oracle.chroma_avg(stack))
    P.ptm_lrgb_finish(info, ycc, coeffs)
    blocks[1][:] = ycc
    P.ptm_scale_coefficients(hdr, coeffs, blocks)
    want = oracle.encode_lrgb(stack, M)
    np.testing.assert_array_equal(blocks[1].reshape(H, W, 3), want["rgb"])
    np.testing.assert_allclose(np.array(hdr.scale), want["scale"], rtol=2e-6)
    assert np.abs(np.array(hdr.bias) - want["bias"]).max() <= 1
    d = np.abs(blocks[0].reshape(H, W, 6).astype(int) - want["coef"].astype(int))
    assert d.max() <= 1 and (d != 0).mean() < 0.01
    # the fused entry gives the same blocks as the step-by-step sequence
    hdr2 = P.ptm_alloc_header("PTM_FORMAT_LRGB", W, H)
    blocks2 = P.ptm_alloc_blocks(hdr2)
    P.ptm_encode_stack(hdr2, info, stack, M, blocks2)
    np.testing.assert_array_equal(blocks2[0], blocks[0])
    np.testing.assert_array_equal(blocks2[1], blocks[1])
    assert list(hdr2.bias) == list(hdr.bias) and list(hdr2.scale) == list(hdr.scale)
    # relight through the header (scale goes through %f in a real file; here it is the float itself)
    out = P.ptm_relight(hdr, blocks, 0.3, 0.1)
    ref = oracle.relight_lrgb(blocks[0].reshape(H, W, 6), blocks[1].reshape(H, W, 3), np.array(hdr.scale, np.float32),
                              np.array(hdr.bias, np.int32), 0.3, 0.1)
    np.testing.assert_array_equal(out, ref)


def test_reference_call_sequence_rgb_and_uint(oracle, dome1):
    from rti_b200 import ptmlib as P
    W, H = 34, 27
    stack = oracle.synth_stack(0xAB2, 1, W, H, dome1)
    M = P.ptm_svd(dome1)
    info = P.image_info(W, H, 60)
    hdr = P.ptm_alloc_header("PTM_FORMAT_RGB", W, H)
    blocks = P.ptm_alloc_blocks(hdr)
    # rti-builder/ptm-encoder.c:272-284: one fit per channel with buffer + r, stride 3
    coeffs = np.stack([P.ptm_fit_poly_jsample(info, stack, 3, M, channel=r) for r in range(3)])
    P.ptm_scale_coefficients(hdr, coeffs, blocks)
    want = oracle.encode_rgb(stack, M)
    np.testing.assert_allclose(np.array(hdr.scale), want["scale"], rtol=2e-6)
    d = np.abs(np.stack(blocks).reshape(3, H, W, 6).astype(int) - want["coef"].astype(int))
    assert d.max() <= 1 and (d != 0).mean() < 0.02
    # ptm_fit_poly_uint (rti-builder/ptmlib.c:727): unsigned int samples, stride 1
    L = stack.astype(np.uint32).sum(-1, dtype=np.uint32)[..., None]          # LUM() of ptmlib.h:160
    got = P.ptm_fit_poly_uint(P.image_info(W, H, 60, 1), L, 1, M)
    want_u = oracle.fit_u32(np.ascontiguousarray(L), M)
    np.testing.assert_allclose(got.reshape(H, W, 6), want_u, rtol=0, atol=2e-2)


def test_stack_mirror_is_never_stale(oracle, dome1):
    """The step-by-step entries mirror the caller's stack on the device for the length of one call sequence
    (rti_b200/csrc/ptm_api_host.cu).  A caller that re-uses its buffer for NEW samples must get new results:
    a second fit of the same channel, or a second averaging pass, uploads again."""
    from rti_b200 import ptmlib as P
    W, H = 70, 37
    M = P.ptm_svd(dome1)
    info = P.image_info(W, H, 60)
    buf = np.ascontiguousarray(oracle.synth_stack(0xAB3, 0, W, H, dome1))
    first = P.ptm_fit_poly_jsample(info, buf, 3, M)
    avg1 = P.ptm_cbcr_avg(info, buf)                      # served from the mirror
    np.testing.assert_array_equal(avg1.reshape(H, W, 3), oracle.chroma_avg(buf))
    other = oracle.synth_stack(0xAB4, 0, W, H, dome1)
    buf[...] = other                                      # same memory, new stack
    second = P.ptm_fit_poly_jsample(info, buf, 3, M)
    assert not np.array_equal(first, second)
    np.testing.assert_allclose(second.reshape(H, W, 6), oracle.fit_u8(other, M), rtol=0, atol=2e-3)
    avg2 = P.ptm_cbcr_avg(info, buf)
    np.testing.assert_array_equal(avg2.reshape(H, W, 3), oracle.chroma_avg(other))
    buf[...] = oracle.synth_stack(0xAB5, 0, W, H, dome1)  # changed again WITHOUT a new fit: a second averaging pass
    avg3 = P.ptm_cbcr_avg(info, buf)                      # must not be answered from the mirror
    np.testing.assert_array_equal(avg3.reshape(H, W, 3), oracle.chroma_avg(buf))
    # channel fits in the reference's order (buffer + r), each from the mirror after the first
    rgb = np.ascontiguousarray(oracle.synth_stack(0xAB6, 1, W, H, dome1))
    for r in range(3):
        got = P.ptm_fit_poly_jsample(info, rgb, 3, M, channel=r)
        np.testing.assert_allclose(got.reshape(H, W, 6), oracle.fit_u8(rgb, M, r), rtol=0, atol=2e-3)
    # ... and in another order (a smaller channel offset after a larger one is a new sequence)
    for r in (2, 0, 1):
        got = P.ptm_fit_poly_jsample(info, rgb, 3, M, channel=r)
        np.testing.assert_allclose(got.reshape(H, W, 6), oracle.fit_u8(rgb, M, r), rtol=0, atol=2e-3)


def test_pinned_and_pageable_buffers_agree(oracle, dome1):
    """Pageable operands are staged through the pinned ring by host threads, pinned ones are copied directly:
    same bytes either way, for an image large enough to span several staging chunks (28 MB per light pair)."""
    import torch
    from rti_b200 import ptmlib as P
    from rti_b200.device import Context
    W, H = 2048, 600
    L = dome1[:24]
    M = P.ptm_svd(L)
    stack = oracle.synth_stack(0xAB7, 0, W, H, L)
    pinned = torch.from_numpy(stack).pin_memory()
    c = Context(0)
    a = c.encode_host(stack, M, planes=1, want_coeffs=True)
    b = c.encode_host(pinned.numpy(), M, planes=1, want_coeffs=True)
    for x, y in zip(a["blocks"], b["blocks"]):
        np.testing.assert_array_equal(x, y)
    np.testing.assert_array_equal(a["coeffs"], b["coeffs"])
    c.close()
    info = P.image_info(W, H, len(L))
    f1 = P.ptm_fit_poly_jsample(info, stack, 3, M)
    f2 = P.ptm_fit_poly_jsample(info, pinned.numpy(), 3, M)
    np.testing.assert_array_equal(f1, f2)
    rows = [0, 299, 599]
    np.testing.assert_allclose(f1.reshape(H, W, 6)[rows], oracle.fit_u8(np.ascontiguousarray(stack[:, rows]), M),
                               rtol=0, atol=2e-3)


def test_scale_after_finish_is_verified_not_trusted(oracle, dome1):
    """With pinned (registered) caller buffers ptm_lrgb_finish keeps its float results on the device and quantises them
    ahead of the ptm_scale_coefficients that normally follows; that call then only verifies, bit for bit, that the host
    buffer still holds those floats while the bytes already travel back (rti_b200/csrc/ptm_api_host.cu).  Unchanged
    floats must give the ordinary path's bytes; a caller that CHANGES a coefficient between the two calls (the
    reference's encoder is free to) must get the bytes of the changed floats -- never the speculative ones."""
    import torch
    from rti_b200 import ptmlib as P
    W, H = 256, 130                                                  # even pixel count: the compare covers every byte
    M = P.ptm_svd(dome1)
    info = P.image_info(W, H, 60)
    stack = np.ascontiguousarray(oracle.synth_stack(0xAB8, 0, W, H, dome1))

    def sequence(pinned, poke):
        hdr = P.ptm_alloc_header("PTM_FORMAT_LRGB", W, H)
        hold = []

        def buf(a):
            if not pinned:
                return a
            t = torch.from_numpy(a).pin_memory()
            hold.append(t)
            return t.numpy()

        blocks = [buf(b) for b in P.ptm_alloc_blocks(hdr)]
        coeffs = buf(np.zeros((W * H, 6), np.float32))
        M32 = np.ascontiguousarray(M, dtype=np.float32)
        L = P.lib()
        import ctypes as C
        L.ptm_fit_poly_jsample(C.byref(info), stack.ctypes.data, 3, M32.ctypes.data, coeffs.ctypes.data)
        L.ptm_cbcr_avg(C.byref(info), stack.ctypes.data, blocks[1].ctypes.data)
        L.ptm_lrgb_finish(C.byref(info), blocks[1].ctypes.data, coeffs.ctypes.data)
        if poke == "extremum":
            coeffs[12345, 2] = 1.0e6          # moves the global maximum of c2: every byte of that coefficient changes
        elif poke == "last":
            coeffs[-1, 5] += 0.75             # the very last float of the plane
        L.ptm_scale_coefficients(C.byref(hdr), coeffs.ctypes.data, P._block_ptrs(blocks))
        return (np.array(blocks[0], copy=True), np.array(blocks[1], copy=True), np.array(hdr.scale, np.float32),
                np.array(hdr.bias, np.int32), np.array(coeffs, copy=True))

    for poke in (None, "extremum", "last"):
        want = sequence(False, poke)           # pageable buffers: the ordinary path
        got = sequence(True, poke)             # pinned buffers: speculation + verification
        for a, b in zip(got, want):
            np.testing.assert_array_equal(a, b)
    plain = sequence(True, None)
    poked = sequence(True, "extremum")
    assert not np.array_equal(plain[0], poked[0]) and poked[2][2] > 100 * plain[2][2]
    # and the whole sequence against the oracle
    ref = oracle.encode_lrgb(stack, M)
    np.testing.assert_array_equal(plain[1].reshape(H, W, 3), ref["rgb"])
    d = np.abs(plain[0].reshape(H, W, 6).astype(int) - ref["coef"].astype(int))
    assert d.max() <= 1 and (d != 0).mean() < 0.01
