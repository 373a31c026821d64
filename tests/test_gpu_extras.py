"""GPU parity of the remaining per-pixel pieces: ptm_normal as a map,
PTM_FORMAT_LUM encode and relight, the encoder's disabled L = R + G + B branch (mean and median chroma),
and the fused RGB -> YCbCr front end.  Each against its oracle function (oracle/ptm_oracle.c cites the
reference lines)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, assert_byte_parity

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ctx():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from rti_b200.device import Context
    torch.cuda.set_device(0)
    c = Context(0, stream=torch.cuda.current_stream().cuda_stream)
    yield c
    c.close()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def gpu_lrgb_like(ctx, stack, M, **mode):
    """fit + quantise through the LRGB-shaped entries with a fit mode set."""
    from rti_b200.device import NCOEF
    n, H, W, _ = stack.shape
    P = H * W
    ctx.set_matrix(M)
    ctx.set_fit_mode(**mode)
    try:
        coeffs = torch.empty((P, NCOEF), dtype=torch.float32, device="cuda")
        colour = torch.empty((P, 3), dtype=torch.uint8, device="cuda")
        coef = torch.empty((P, NCOEF), dtype=torch.uint8, device="cuda")
        keys = torch.empty(12, dtype=torch.int32, device="cuda")
        sb = torch.empty(12, dtype=torch.int32, device="cuda")
        ctx.keys_init(keys)
        ctx.fit_lrgb(dev(stack), P * 3, P, coeffs, colour, keys)
        ctx.quantise(coeffs, P, keys, coef, sb)
        torch.cuda.synchronize()
        # the fused two-launch entry and the host-buffer entry must agree with the step-by-step calls
        coef2 = torch.empty_like(coef)
        colour2 = torch.empty_like(colour)
        sb2 = torch.empty_like(sb)
        ctx.encode_lrgb_dev(dev(stack), P * 3, P, coeffs, colour2, coef2, sb2)
        torch.cuda.synchronize()
        assert torch.equal(coef, coef2) and torch.equal(colour, colour2) and torch.equal(sb, sb2)
        host = ctx.encode_host(stack, M, planes=1)
        np.testing.assert_array_equal(host["blocks"][0].reshape(P, 6), coef.cpu().numpy())
        np.testing.assert_array_equal(host["blocks"][1].reshape(P, 3), colour.cpu().numpy())
    finally:
        ctx.set_fit_mode()
    raw = sb.cpu().numpy()
    return dict(coeffs=coeffs.cpu().numpy().reshape(H, W, NCOEF), colour=colour.cpu().numpy().reshape(H, W, 3),
                coef=coef.cpu().numpy().reshape(H, W, NCOEF), scale=raw[:6].view(np.float32).copy(), bias=raw[6:].copy())


def rel_err(got, want):
    scale = np.abs(want).reshape(-1, 6).max(0)
    return float((np.abs(got - want).reshape(-1, 6).max(0) / scale).max())


# (64, 48): three ring tiles; (37, 29): register + generic kernels; (1040, 3): a full tile + a 16-px-granular tail
SHAPES = [(64, 48), (37, 29), (1040, 3)]


@pytest.mark.parametrize("W,H", SHAPES)
def test_lum_encode_vs_oracle(ctx, oracle, dome1, M60, W, H):
    """PTM_FORMAT_LUM (rti-builder/ptm-encoder.c:286-305): colour block = the averaged YCbCr triplet, identical;
    coefficients unnormalised."""
    stack = oracle.synth_stack(0x10 + W, 0, W, H, dome1)
    want = oracle.encode_lum(stack, M60)
    got = gpu_lrgb_like(ctx, stack, M60, lum=True)
    np.testing.assert_array_equal(got["colour"], want["ycc"])
    assert rel_err(got["coeffs"], want["coeffs"]) <= 1e-5
    np.testing.assert_allclose(got["scale"], want["scale"], rtol=3e-6)
    assert_byte_parity(got["coef"], want["coef"], got["bias"], want["bias"], "LUM bytes %dx%d" % (W, H))


@pytest.mark.parametrize("W,H", SHAPES)
def test_rgb_input_front_end_vs_oracle(ctx, oracle, dome1, M60, W, H):
    """an RGB stack through the fused per-sample JFIF conversion equals the host conversion
    followed by the plain LRGB encode -- colour block identical, the rest at the usual bars."""
    rgb_stack = oracle.synth_stack(0x20 + W, 1, W, H, dome1)
    want = oracle.encode_lrgb(oracle.rgb_to_ycbcr(rgb_stack), M60)
    got = gpu_lrgb_like(ctx, rgb_stack, M60, rgb_input=True)
    np.testing.assert_array_equal(got["colour"], want["rgb"])
    assert rel_err(got["coeffs"], want["coeffs"]) <= 1e-5
    np.testing.assert_allclose(got["scale"], want["scale"], rtol=3e-6)
    assert_byte_parity(got["coef"], want["coef"], got["bias"], want["bias"], "RGB-input LRGB bytes %dx%d" % (W, H))
    # and the plain path on the pre-converted stack gives the very same bytes (the conversion is the only difference)
    plain = gpu_lrgb_like(ctx, oracle.rgb_to_ycbcr(rgb_stack), M60)
    np.testing.assert_array_equal(plain["coef"], got["coef"])
    np.testing.assert_array_equal(plain["colour"], got["colour"])


@pytest.mark.parametrize("median", [False, True])
def test_lsum_branch_vs_oracle(ctx, oracle, dome1, M60, median):
    """rti-builder/ptm-encoder.c:356-400 (disabled there): L = R + G + B fit, colour = 256 * chroma / L_0; chroma =
    truncated mean (the reference's) or the median its FIXME asks for."""
    from rti_b200.device import NCOEF
    W, H = 50, 21
    stack = oracle.synth_stack(0x33, 1, W, H, dome1)
    stack[:, 3, 4] = 0                                      # L_0 == 0: a division by zero in the reference, 0 here
    want = oracle.fit_lsum(stack, M60, median=median)
    P = W * H
    ctx.set_matrix(M60)
    coeffs = torch.empty((P, NCOEF), dtype=torch.float32, device="cuda")
    colour = torch.empty((P, 3), dtype=torch.uint8, device="cuda")
    coef = torch.empty((P, NCOEF), dtype=torch.uint8, device="cuda")
    keys = torch.empty(12, dtype=torch.int32, device="cuda")
    sb = torch.empty(12, dtype=torch.int32, device="cuda")
    ctx.keys_init(keys)
    ctx.fit_lsum(dev(stack), P * 3, P, coeffs, colour, keys, median=median)
    ctx.quantise(coeffs, P, keys, coef, sb)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(colour.cpu().numpy().reshape(H, W, 3), want["colour"])
    assert rel_err(coeffs.cpu().numpy().reshape(H, W, 6), want["coeffs"]) <= 1e-5
    raw = sb.cpu().numpy()
    np.testing.assert_allclose(raw[:6].view(np.float32), want["scale"], rtol=3e-6)
    assert_byte_parity(coef.cpu().numpy().reshape(H, W, 6), want["coef"], raw[6:], want["bias"],
                       "L = R + G + B branch (%s chroma)" % ("median" if median else "mean"))


def test_normal_map_vs_oracle(ctx, oracle, dome1, M60):
    """ptm_normal (rti-builder/ptmlib.c:808-814) for every pixel: bit-identical floats (NaNs where the reference's
    square root goes negative), from float coefficients and from a quantised block."""
    W, H = 61, 33
    stack = oracle.synth_stack(0x44, 0, W, H, dome1)
    enc = oracle.encode_lrgb(stack, M60)
    P = W * H
    out = torch.empty((P, 3), dtype=torch.float32, device="cuda")
    ctx.normal_map(dev(enc["coeffs"]), P, out)
    want = oracle.normal_map(coeffs=enc["coeffs"])
    got = out.cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(want))
    np.testing.assert_array_equal(got[~np.isnan(want)], want[~np.isnan(want)])
    scale = oracle.header_roundtrip(enc["scale"])
    ctx.normal_map_blocks(dev(enc["coef"]), scale, enc["bias"], P, out)
    want = oracle.normal_map(coef_bytes=enc["coef"], scale_vals=scale, bias=enc["bias"])
    got = out.cpu().numpy()
    assert np.array_equal(np.isnan(got), np.isnan(want))
    np.testing.assert_array_equal(got[~np.isnan(want)], want[~np.isnan(want)])
    assert np.isfinite(want).mean() > 0.3, "the test image should give mostly finite normals"
    # the ptmlib.h drop-in: ptm_normal_map over host blocks, and the scalar ptm_normal for one pixel
    from rti_b200 import ptmlib
    hdr = ptmlib.ptm_alloc_header("PTM_FORMAT_LRGB", W, H)
    for k in range(6):
        hdr.scale[k], hdr.bias[k] = float(scale[k]), int(enc["bias"][k])
    nm = ptmlib.ptm_normal_map(hdr, [np.ascontiguousarray(enc["coef"]).reshape(P, 6), np.ascontiguousarray(enc["rgb"]).reshape(P, 3)])
    np.testing.assert_array_equal(nm[~np.isnan(want)], want[~np.isnan(want)])


def test_lum_relight_vs_oracle(ctx, oracle, dome1, M60):
    """rti-builder/ptmlib.c:601-609: (CLIP(POLY), Cb, Cr) rasters with block 1 read as 2-byte (cr, cb) records."""
    rng = np.random.default_rng(5)
    W, H = 52, 35
    stack = oracle.synth_stack(0x55, 0, W, H, dome1)
    enc = oracle.encode_lum(stack, M60)
    scale = oracle.header_roundtrip(enc["scale"])
    crcb = rng.integers(0, 256, (H, W, 2), dtype=np.uint8)
    uv = np.concatenate([[[0, 0], [0.5, 0.5], [1, 0], [0, -1]], rng.uniform(-0.9, 0.9, size=(8, 2))]).astype(np.float32)
    out = torch.empty((len(uv), H, W, 3), dtype=torch.uint8, device="cuda")
    ctx.relight_lum(dev(enc["coef"]), dev(crcb), W, H, scale, enc["bias"], uv, out)
    got = out.cpu().numpy()
    for i, (u, v) in enumerate(uv):
        np.testing.assert_array_equal(got[i], oracle.relight_lum(enc["coef"], crcb, scale, enc["bias"], float(u), float(v)))


def test_lum_command_line_round_trip(oracle, dome1, M60, tmp_path):
    """ptm-encoder -f PTM_FORMAT_LUM -> file -> ptm-decoder: the file carries the reference writer's layout
    (rti-builder/ptmlib.c:333-347: six coefficients, Cr, Cb per pixel, after the colour matrix lines) and the
    decoded raster equals the oracle's LUM relight of the encoder's own blocks.  (The reference's reader cannot
    read its own LUM files, SURVEY.md appendix A.11, so there is no reference decoder to compare with.)"""
    from test_gpu_cli import run, write_inputs
    W, H = 40, 24
    stack = oracle.synth_stack(0x66, 0, W, H, dome1)
    lp = write_inputs(oracle, str(tmp_path), stack, dome1)
    out = os.path.join(str(tmp_path), "lum.ptm")
    r = run([os.path.join(ROOT, "rti_b200", "bin", "ptm-encoder"), "-f", "PTM_FORMAT_LUM", "-o", out, lp])
    assert r.returncode == 0, r.stderr.decode()
    blob = open(out, "rb").read()
    lines = blob.split(b"\n", 7)
    assert lines[0] == b"PTM_1.2" and lines[1] == b"PTM_FORMAT_LUM" and len(lines[6].split()) == 16
    payload = np.frombuffer(lines[7], np.uint8).reshape(H, W, 8)
    want = oracle.encode_lum(stack, M60)
    assert_byte_parity(payload[..., :6], want["coef"], [int(x) for x in lines[5].split()], want["bias"], "LUM file bytes")
    np.testing.assert_array_equal(payload[..., 6], want["ycc"][..., 2])   # Cr first
    np.testing.assert_array_equal(payload[..., 7], want["ycc"][..., 1])
    scale = np.array([float(x) for x in lines[4].split()], np.float32)
    bias = np.array([int(x) for x in lines[5].split()], np.int32)
    r = run([os.path.join(ROOT, "rti_b200", "bin", "ptm-decoder"), out, "0.3", "-0.4"])
    assert r.returncode == 0, r.stderr.decode()
    raster = oracle.read_rawj(r.stdout)
    ref = oracle.relight_lum(np.ascontiguousarray(payload[..., :6]), np.ascontiguousarray(payload[..., 6:8]),
                             scale, bias, 0.3, -0.4)
    np.testing.assert_array_equal(raster, ref)
