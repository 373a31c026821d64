"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every check calls the CUDA
path through the C ABI (ctypes -> libptm_b200.so) and compares with the CPU oracle
on the same seeded inputs, or with the committed outputs of the reference binaries.

Bars (BASELINE.json north_star): float coefficients within 1e-5 relative to the
coefficient scale; integer / byte outputs identical except +-1 LSB where the float
summation order lands on a truncation boundary (count reported); relit bytes +-1.
"""
import numpy as np
import pytest

from conftest import assert_byte_parity, golden_path

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ctx():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from rti_b200.device import Context
    torch.cuda.set_device(0)
    c = Context(0, stream=torch.cuda.current_stream().cuda_stream)
    c.set_relight_mode("exact")   # this file pins the bit-exact kernels; tests/test_gpu_relight_fast.py has the +-1 ones
    yield c
    c.close()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def gpu_encode_lrgb(ctx, stack, M):
    from rti_b200.device import NCOEF, key_to_float
    n, H, W, _ = stack.shape
    P = H * W
    ctx.set_matrix(M)
    d_stack = dev(stack)
    coeffs = torch.empty((P, NCOEF), dtype=torch.float32, device="cuda")
    rgb = torch.empty((P, 3), dtype=torch.uint8, device="cuda")
    coef = torch.empty((P, NCOEF), dtype=torch.uint8, device="cuda")
    keys = torch.empty(12, dtype=torch.int32, device="cuda")
    sb = torch.empty(12, dtype=torch.int32, device="cuda")
    ctx.keys_init(keys)
    ctx.fit_lrgb(d_stack, P * 3, P, coeffs, rgb, keys)
    ctx.quantise(coeffs, P, keys, coef, sb)
    torch.cuda.synchronize()
    raw = sb.cpu().numpy()
    f = key_to_float(keys.cpu().numpy())
    return dict(coeffs=coeffs.cpu().numpy().reshape(H, W, NCOEF), rgb=rgb.cpu().numpy().reshape(H, W, 3),
                coef=coef.cpu().numpy().reshape(H, W, NCOEF), scale=raw[:6].view(np.float32).copy(),
                bias=raw[6:].copy(), minmax=np.concatenate([-f[:6], f[6:]]))


def coeff_error(got, want, M, stack_y, lrgb=False):
    """max |delta| normalised by the coefficient's scale sum_n |M_kn| * y_n;
    for LRGB the coefficients carry the extra factor 256 / floor(mean y)."""
    n = stack_y.shape[0]
    y = stack_y.astype(np.float64).reshape(n, *got.shape[:-1])
    bound = np.einsum("kn,n...->...k", np.abs(M).astype(np.float64), y)
    if lrgb:
        bound = bound * (256.0 / np.floor(y.sum(0) / n))[..., None]
    return (np.abs(got.astype(np.float64) - want.astype(np.float64)) / np.maximum(bound, 1e-30)).max()


def assert_bytes_close(got, want, what, max_frac=0.01):
    d = np.abs(got.astype(int) - want.astype(int))
    n = int((d != 0).sum())
    print("%s: %d of %d bytes differ (max %d)" % (what, n, d.size, d.max() if d.size else 0))
    assert d.max(initial=0) <= 1, what
    assert n <= max_frac * d.size, what
    return n


# ----------------------------------------------------------------------- synth

def test_synth_matches_host(ctx, oracle, dome1):
    from rti_b200 import lights
    from rti_b200.device import U8, U16, light_q14
    for mode in (0, 1):
        want = oracle.synth_stack(0x5EED0001, mode, 70, 33, dome1)
        P = 70 * 33
        got = torch.empty(want.shape, dtype=torch.uint8, device="cuda")
        ctx.synth(0x5EED0001, mode, U8, 0, P, light_q14(dome1), got, P * 3)
        np.testing.assert_array_equal(got.cpu().numpy(), want)
    L100 = lights.dome100_lights()
    want = oracle.synth_stack(0x5EED0004, 1, 24, 16, L100, dtype=np.uint16)
    got = torch.empty(want.shape, dtype=torch.uint16, device="cuda")
    ctx.synth(0x5EED0004, 1, U16, 0, 24 * 16, light_q14(L100), got, 24 * 16 * 6)
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    # a slab starting at a pixel offset equals the same rows of the whole
    want = oracle.synth_stack(0x5EED0001, 0, 70, 33, dome1, row0=10, rows=5)
    got = torch.empty(want.shape, dtype=torch.uint8, device="cuda")
    ctx.synth(0x5EED0001, 0, U8, 10 * 70, 5 * 70, light_q14(dome1), got, 5 * 70 * 3)
    np.testing.assert_array_equal(got.cpu().numpy(), want)


# ------------------------------------------------------------------ LRGB encode

@pytest.mark.parametrize("W,H,seed", [(64, 48, 0x5EED0001), (37, 29, 0x5EED0002), (256, 192, 5), (3, 1, 6), (1, 1, 7)])
def test_lrgb_encode_vs_oracle(ctx, oracle, dome1, M60, W, H, seed):
    stack = oracle.synth_stack(seed, 0, W, H, dome1)
    want = oracle.encode_lrgb(stack, M60)
    got = gpu_encode_lrgb(ctx, stack, M60)
    # the float path
    err = coeff_error(got["coeffs"], want["coeffs"], M60, stack[..., 0], lrgb=True)
    print("normalised coefficient error %.3g" % err)
    assert err <= 1e-5
    # integer + unfused float paths: exact
    np.testing.assert_array_equal(got["rgb"], want["rgb"])
    if W * H > 1:
        np.testing.assert_allclose(got["minmax"], want["minmax"], rtol=2e-6, atol=1e-4)
        np.testing.assert_allclose(got["scale"], want["scale"], rtol=2e-6)
        assert_byte_parity(got["coef"], want["coef"], got["bias"], want["bias"], "LRGB coefficient bytes %dx%d" % (W, H))


@pytest.mark.parametrize("W,H,seed", [(256, 192, 5), (1040, 37, 8), (2000, 752, 9)])
def test_lrgb_ring_tile_sizes_give_the_same_bytes(ctx, oracle, dome1, M60, W, H, seed, monkeypatch):
    """K1's ring kernel exists with 1024-pixel tiles (8 consumer warps, 2 CTAs / SM) and with 512-pixel tiles
    (4 consumer warps, 4 CTAs / SM; chosen for slabs with few tiles per CTA).  Both must give the oracle's result,
    and each other's bit for bit (same per-pixel arithmetic, extrema merged with integer max), for whole tiles,
    a partial last tile (1040 x 37 = 75 x 512 + 80 = 37 x 1024 + 592) and the plain / fused entry points."""
    from rti_b200.device import NCOEF
    stack = oracle.synth_stack(seed, 0, W, H, dome1)
    want = oracle.encode_lrgb(stack, M60)
    P = W * H
    res = {}
    for tile in ("512", "1024"):
        monkeypatch.setenv("PTM_TMA_TILE", tile)
        got = gpu_encode_lrgb(ctx, stack, M60)
        assert coeff_error(got["coeffs"], want["coeffs"], M60, stack[..., 0], lrgb=True) <= 1e-5
        np.testing.assert_array_equal(got["rgb"], want["rgb"])
        assert_byte_parity(got["coef"], want["coef"], got["bias"], want["bias"], "tile %s, %dx%d" % (tile, W, H))
        # the fused two-launch encode on the same tile size
        d_stack = dev(stack)
        coeffs = torch.empty((P, NCOEF), dtype=torch.float32, device="cuda")
        rgb = torch.zeros((P, 3), dtype=torch.uint8, device="cuda")
        coef = torch.zeros((P, NCOEF), dtype=torch.uint8, device="cuda")
        sb = torch.empty(12, dtype=torch.int32, device="cuda")
        for _ in range(2):
            ctx.encode_lrgb_dev(d_stack, P * 3, P, coeffs, rgb, coef, sb)
        torch.cuda.synchronize()
        np.testing.assert_array_equal(coeffs.cpu().numpy().reshape(H, W, NCOEF), got["coeffs"])
        np.testing.assert_array_equal(coef.cpu().numpy().reshape(H, W, NCOEF), got["coef"])
        np.testing.assert_array_equal(rgb.cpu().numpy().reshape(H, W, 3), got["rgb"])
        res[tile] = got
    monkeypatch.delenv("PTM_TMA_TILE")
    for k in ("coeffs", "coef", "rgb", "scale", "bias"):
        np.testing.assert_array_equal(res["512"][k], res["1024"][k])


@pytest.mark.parametrize("planes", [1, 3])
def test_scratch_float_plane_gives_the_same_blocks(oracle, dome1, M60, planes, monkeypatch):
    """ptmb_set_coeffs_scratch: K2 quantises the end of the float plane first and discards the L2 lines it has read
    (the plane is undefined afterwards).  Bytes, colour block and scale / bias must equal those of an encode that keeps
    its floats -- for a plane that is all tail, and (PTM_K2_TAIL_MB=1) for planes split into a 1 MB tail and a head,
    with pixel counts that are and are not multiples of 16 (the latter take the plain forward kernel)."""
    from rti_b200.device import U8, light_q14
    from rti_b200.encoder import PtmEncoder
    for W, H in ((2000, 752), (1040, 37), (1000, 333)):
        P = W * H
        stack = torch.empty((60, H, W, 3), dtype=torch.uint8, device="cuda")
        keep = PtmEncoder(W, H, M60, planes=planes, keep_coeffs=True)
        keep.ctx.synth(0x5EED0010 + W, 0 if planes == 1 else 1, U8, 0, P, light_q14(dome1), stack, P * 3)
        keep.encode(stack)
        torch.cuda.synchronize()
        for tail in (None, "1"):
            if tail:
                monkeypatch.setenv("PTM_K2_TAIL_MB", tail)
            scratch = PtmEncoder(W, H, M60, planes=planes, keep_coeffs=False)
            for _ in range(3):
                scratch.encode(stack)
            torch.cuda.synchronize()
            assert torch.equal(scratch.coef_bytes, keep.coef_bytes), (W, H, tail)
            assert torch.equal(scratch.sb, keep.sb)
            if planes == 1:
                assert torch.equal(scratch.rgb, keep.rgb)
            if tail:
                monkeypatch.delenv("PTM_K2_TAIL_MB")


@pytest.mark.parametrize("engine", ["ffma", "mma"])
def test_plain_and_fused_entry_points_interleave(ctx, oracle, dome1, M60, engine):
    """The dynamic tile schedule keeps a device counter at 0 between launches; the plain fit (cleared by a
    stream-ordered memset) and the fused encode (cleared by the hand-off block) must be able to alternate
    on one context, at a size where most tiles are drawn from the counter."""
    from rti_b200.device import NCOEF, U8, light_q14
    W, H = 2000, 752
    P = W * H
    n = len(dome1)
    ctx.set_fit_engine(engine)
    ctx.set_matrix(M60)
    stack = torch.empty((n, P, 3), dtype=torch.uint8, device="cuda")
    ctx.synth(0x5EED0009, 0, U8, 0, P, light_q14(dome1), stack, P * 3)
    results = []
    for fused in (False, True, False, True, True, False):
        coeffs = torch.empty((P, NCOEF), dtype=torch.float32, device="cuda")
        rgb = torch.zeros((P, 3), dtype=torch.uint8, device="cuda")
        coef = torch.zeros((P, NCOEF), dtype=torch.uint8, device="cuda")
        sb = torch.empty(12, dtype=torch.int32, device="cuda")
        if fused:
            ctx.encode_lrgb_dev(stack, P * 3, P, coeffs, rgb, coef, sb)
        else:
            keys = torch.empty(12, dtype=torch.int32, device="cuda")
            ctx.keys_init(keys)
            ctx.fit_lrgb(stack, P * 3, P, coeffs, rgb, keys)
            ctx.quantise(coeffs, P, keys, coef, sb)
        torch.cuda.synchronize()
        results.append((rgb, coef, coeffs))
    ctx.set_fit_engine("default")
    for rgb, coef, coeffs in results[1:]:
        assert torch.equal(rgb, results[0][0]) and torch.equal(coef, results[0][1])
        assert torch.equal(coeffs, results[0][2])


@pytest.mark.parametrize("engine", ["ffma", "mma"])
def test_fused_encode_loop_is_deterministic(oracle, dome1, M60, engine):
    """Back-to-back fused encodes (every launch programmatically dependent on its predecessor, tiles drawn from
    the device counter, keys handed over through the mailbox): whichever block takes which tile, every result
    equals the first bit for bit.  tools/soak.py runs the same for thousands of encodes at the bench sizes."""
    from rti_b200.device import U8, light_q14
    from rti_b200.encoder import PtmEncoder
    for W, H, reps in ((2000, 37, 300), (2000, 752, 60)):
        P = W * H
        stack = torch.empty((len(dome1), H, W, 3), dtype=torch.uint8, device="cuda")
        enc = PtmEncoder(W, H, M60, engine=engine)
        enc.ctx.synth(0x5EED0001, 0, U8, 0, P, light_q14(dome1), stack, P * 3)
        enc.encode(stack)
        torch.cuda.synchronize()
        ref = (enc.coef_bytes.clone(), enc.rgb.clone(), enc.coeffs.clone(), enc.sb.clone())
        for i in range(reps):
            enc.encode(stack)
            if i % 29 == 0 or i == reps - 1:
                torch.cuda.synchronize()
                assert torch.equal(enc.coef_bytes, ref[0]) and torch.equal(enc.rgb, ref[1])
                assert torch.equal(enc.coeffs, ref[2]) and torch.equal(enc.sb, ref[3])


def test_degenerate_inputs_follow_the_reference(ctx, oracle, dome1, M60):
    """Inputs the reference does not guard against (rti-builder/ptm-encoder.c:342 divides by the mean luma,
    rti-builder/ptmlib.c:860-862 by the coefficient range): a pixel that is black in every image gets NaN
    coefficients, which fminf/fmaxf skip and CLIP turns into byte 0; a constant stack has range 0, scale 0
    and the x86 float->int result INT_MIN as bias.  The GPU path reproduces both."""
    W, H = 32, 8
    stack = oracle.synth_stack(77, 0, W, H, dome1)
    stack[:, 3, 5, :] = 0          # black in every light
    stack[:, 4, 6, 0] = 0          # zero luma, chroma kept
    want = oracle.encode_lrgb(stack, M60)
    for engine in ("ffma", "mma"):
        ctx.set_fit_engine(engine)
        got = gpu_encode_lrgb(ctx, stack, M60)
        np.testing.assert_array_equal(got["rgb"], want["rgb"])
        assert np.isnan(got["coeffs"][3, 5]).all() and np.isnan(got["coeffs"][4, 6]).all()
        assert np.array_equal(np.isnan(got["coeffs"]), np.isnan(want["coeffs"]))
        # 256 pixels: the range is a few hundred ulps of the extrema, which carry the oracle's summation noise
        np.testing.assert_allclose(got["scale"], want["scale"], rtol=5e-6)
        assert np.abs(got["bias"].astype(np.int64) - want["bias"].astype(np.int64)).max() <= 1
        assert (got["coef"][3, 5] == 0).all() and (got["coef"][4, 6] == 0).all()
        assert_byte_parity(got["coef"], want["coef"], got["bias"], want["bias"], "bytes with NaN pixels (%s)" % engine)
    ctx.set_fit_engine("default")
    flat = np.full((60, H, W, 3), 128, dtype=np.uint8)
    flat[..., 0] = 100
    want = oracle.encode_lrgb(flat, M60)
    got = gpu_encode_lrgb(ctx, flat, M60)
    np.testing.assert_array_equal(got["rgb"], want["rgb"])
    # every pixel has the same coefficients whatever the summation order, so the range is exactly 0
    assert (got["scale"] == 0).all() and (want["scale"] == 0).all()
    np.testing.assert_array_equal(got["bias"], want["bias"])
    assert (got["bias"] == -2**31).all()
    # byte = CLIP(c * inf + bias): 255 for positive coefficients, 0 for negative ones; a coefficient whose
    # float sum cancels to +-0 or to rounding noise may differ in sign between summation orders
    sure = np.abs(want["coeffs"]) > 1e-3
    np.testing.assert_array_equal(got["coef"][sure], want["coef"][sure])


def test_lrgb_quantise_is_exact_given_same_floats(ctx, oracle, dome1, M60):
    """K2 alone: fed the ORACLE's float coefficients, bytes, scale and bias are identical."""
    stack = oracle.synth_stack(11, 0, 96, 50, dome1)
    want = oracle.encode_lrgb(stack, M60)
    P = 96 * 50
    coeffs = dev(want["coeffs"].reshape(P, 6))
    keys = torch.empty(12, dtype=torch.int32, device="cuda")
    sb = torch.empty(12, dtype=torch.int32, device="cuda")
    out = torch.empty((P, 6), dtype=torch.uint8, device="cuda")
    ctx.keys_init(keys)
    ctx.minmax(coeffs, P, keys)
    ctx.quantise(coeffs, P, keys, out, sb)
    raw = sb.cpu().numpy()
    np.testing.assert_array_equal(raw[:6].view(np.float32), want["scale"])
    np.testing.assert_array_equal(raw[6:], want["bias"])
    np.testing.assert_array_equal(out.cpu().numpy().reshape(50, 96, 6), want["coef"])


def test_unfused_steps_match_fused_kernel(ctx, oracle, dome1, M60):
    """The ptmlib.h-shaped steps (fit channel, chroma average, colour/normalise, minmax)
    give the same result as the fused K1."""
    from rti_b200.device import U8
    stack = oracle.synth_stack(12, 0, 50, 31, dome1)
    P = 50 * 31
    fused = gpu_encode_lrgb(ctx, stack, M60)
    d_stack = dev(stack)
    coeffs = torch.empty((P, 6), dtype=torch.float32, device="cuda")
    avg = torch.empty((P, 3), dtype=torch.uint8, device="cuda")
    ctx.set_matrix(M60)
    ctx.fit_channel(d_stack, U8, 3, P * 3, P, coeffs)
    ctx.chroma_avg(d_stack, P * 3, P, 60, avg)
    np.testing.assert_array_equal(avg.cpu().numpy().reshape(31, 50, 3), oracle.chroma_avg(stack))
    ctx.lrgb_colour_normalise(avg, coeffs, P)
    np.testing.assert_array_equal(avg.cpu().numpy().reshape(31, 50, 3), fused["rgb"])
    np.testing.assert_array_equal(coeffs.cpu().numpy().reshape(31, 50, 6), fused["coeffs"])


def test_unaligned_and_strided_stack(ctx, oracle, dome1, M60):
    """Image stride larger than an image, base pointer off by one byte: the generic path."""
    stack = oracle.synth_stack(13, 0, 21, 10, dome1)
    P = 210
    want = oracle.encode_lrgb(stack, M60)
    stride = P * 3 + 7
    buf = torch.zeros(60 * stride + 1, dtype=torch.uint8, device="cuda")
    view = buf[1:].reshape(60, stride)
    view[:, :P * 3] = dev(stack.reshape(60, P * 3))
    coeffs = torch.empty((P, 6), dtype=torch.float32, device="cuda")
    rgb = torch.empty((P, 3), dtype=torch.uint8, device="cuda")
    keys = torch.empty(12, dtype=torch.int32, device="cuda")
    ctx.set_matrix(M60)
    ctx.keys_init(keys)
    ctx.fit_lrgb(buf.data_ptr() + 1, stride, P, coeffs, rgb, keys)
    np.testing.assert_array_equal(rgb.cpu().numpy().reshape(10, 21, 3), want["rgb"])
    assert coeff_error(coeffs.cpu().numpy().reshape(10, 21, 6), want["coeffs"], M60, stack[..., 0], lrgb=True) <= 1e-5


@pytest.mark.parametrize("name", ["lrgb_64x48", "lrgb_37x29"])
def test_lrgb_vs_reference_binaries_golden(ctx, oracle, dome1, M60, name):
    """Against what the reference's unmodified ptm-encoder / ptm-decoder produced."""
    g = np.load(golden_path(name + ".npz"))
    W, H = int(g["width"]), int(g["height"])
    stack = oracle.synth_stack(int(g["seed"]), 0, W, H, dome1)
    got = gpu_encode_lrgb(ctx, stack, M60)
    np.testing.assert_array_equal(got["rgb"], g["block1"])
    np.testing.assert_allclose(oracle.header_roundtrip(got["scale"]), g["scale_text"], atol=2e-6)
    assert np.abs(got["bias"].astype(int) - g["bias"]).max() <= 1
    assert_bytes_close(got["coef"], g["block0"], name + " coefficient bytes vs reference encoder")
    # relight the REFERENCE's blocks on the GPU and compare with the reference decoder's rasters
    out = torch.empty((len(g["directions"]), H, W, 3), dtype=torch.uint8, device="cuda")
    ctx.relight_lrgb(dev(g["block0"]), dev(g["block1"]), W, H, g["scale_text"], g["bias"], g["directions"], out)
    d = np.abs(out.cpu().numpy().astype(int) - g["rasters"].astype(int))
    print("relit bytes differing from reference decoder: %d of %d" % ((d != 0).sum(), d.size))
    assert d.max() <= 1


# ------------------------------------------------------------------- RGB format

def gpu_encode_rgb(ctx, stack, M):
    from rti_b200.device import key_to_float
    n, H, W, _ = stack.shape
    P = H * W
    ctx.set_matrix(M)
    coeffs = torch.empty((3, P, 6), dtype=torch.float32, device="cuda")
    coef = torch.empty((3, P, 6), dtype=torch.uint8, device="cuda")
    keys = torch.empty(12, dtype=torch.int32, device="cuda")
    sb = torch.empty(12, dtype=torch.int32, device="cuda")
    ctx.keys_init(keys)
    ctx.fit_rgb(dev(stack), P * 3, P, coeffs, P * 6, keys)
    for b in range(3):
        ctx.quantise(coeffs[b], P, keys, coef[b], sb if b == 0 else None)
    raw = sb.cpu().numpy()
    f = key_to_float(keys.cpu().numpy())
    return dict(coeffs=coeffs.cpu().numpy().reshape(3, H, W, 6), coef=coef.cpu().numpy().reshape(3, H, W, 6),
                scale=raw[:6].view(np.float32).copy(), bias=raw[6:].copy(), minmax=np.concatenate([-f[:6], f[6:]]))


@pytest.mark.parametrize("W,H", [(40, 24), (33, 7)])
def test_rgb_encode_vs_oracle(ctx, oracle, dome1, M60, W, H):
    stack = oracle.synth_stack(0x5EED0003, 1, W, H, dome1)
    want = oracle.encode_rgb(stack, M60)
    got = gpu_encode_rgb(ctx, stack, M60)
    for ch in range(3):
        assert coeff_error(got["coeffs"][ch], want["coeffs"][ch], M60, stack[..., ch]) <= 1e-5
    np.testing.assert_allclose(got["minmax"], want["minmax"], rtol=2e-6, atol=1e-4)   # plane 0 only
    np.testing.assert_allclose(got["scale"], want["scale"], rtol=2e-6)
    assert_byte_parity(got["coef"], want["coef"], got["bias"], want["bias"], "RGB coefficient bytes")


def test_rgb_vs_reference_binaries_golden(ctx, oracle, dome1, M60):
    g = np.load(golden_path("rgb_40x24.npz"))
    W, H = int(g["width"]), int(g["height"])
    stack = oracle.synth_stack(int(g["seed"]), 1, W, H, dome1)
    got = gpu_encode_rgb(ctx, stack, M60)
    ref = np.stack([g["block0"], g["block1"], g["block2"]])
    np.testing.assert_allclose(oracle.header_roundtrip(got["scale"]), g["scale_text"], atol=2e-6)
    assert np.abs(got["bias"].astype(int) - g["bias"]).max() <= 1
    assert_bytes_close(got["coef"], ref, "RGB coefficient bytes vs reference encoder")
    out = torch.empty((len(g["directions"]), H, W, 3), dtype=torch.uint8, device="cuda")
    ctx.relight_rgb(dev(ref[0]), dev(ref[1]), dev(ref[2]), W, H, g["scale_text"], g["bias"], g["directions"], out)
    d = np.abs(out.cpu().numpy().astype(int) - g["rasters"].astype(int))
    print("relit bytes differing from reference decoder: %d of %d" % ((d != 0).sum(), d.size))
    assert d.max() <= 1


# ------------------------------------------------------------------ wider samples

def test_fit_channel_uint16_and_uint32_vs_oracle(ctx, oracle):
    from rti_b200 import lights
    from rti_b200.device import U16, U32
    L = lights.dome100_lights()
    M = oracle.pinv(L)
    s16 = oracle.synth_stack(0x5EED0004, 1, 24, 16, L, dtype=np.uint16)
    P = 24 * 16
    ctx.set_matrix(M)
    for ch in range(3):
        want = oracle.fit_u32(s16.astype(np.uint32), M, ch)   # rti-builder/ptmlib.c:727 is the >8 bit entry
        out = torch.empty((P, 6), dtype=torch.float32, device="cuda")
        d16 = dev(s16.view(np.int16)).view(torch.uint16)
        ctx.fit_channel(d16.data_ptr() + 2 * ch, U16, 3, P * 3, P, out)
        assert coeff_error(out.cpu().numpy().reshape(16, 24, 6), want, M, s16[..., ch]) <= 1e-5
        d32 = dev(s16.astype(np.int32))
        ctx.fit_channel(d32.data_ptr() + 4 * ch, U32, 3, P * 3, P, out)
        assert coeff_error(out.cpu().numpy().reshape(16, 24, 6), want, M, s16[..., ch]) <= 1e-5


# ---------------------------------------------------------------------- relight

def test_relight_matches_oracle_exactly(ctx, oracle, dome1, M60):
    """K3 keeps the reference's association and division, so it is bit-exact against the
    restated loop, including the %f round trip of the scale through the PTM header."""
    stack = oracle.synth_stack(21, 0, 52, 35, dome1)
    enc = oracle.encode_lrgb(stack, M60)
    scale = oracle.header_roundtrip(enc["scale"])
    rng = np.random.default_rng(3)
    uv = np.concatenate([[[0, 0], [0.5, 0.5], [1, 0], [0, -1], [-0.7071, 0.7071]],
                         rng.uniform(-0.7, 0.7, size=(27, 2))]).astype(np.float32)
    out = torch.empty((len(uv), 35, 52, 3), dtype=torch.uint8, device="cuda")
    ctx.relight_lrgb(dev(enc["coef"]), dev(enc["rgb"]), 52, 35, scale, enc["bias"], uv, out)
    got = out.cpu().numpy()
    for i, (u, v) in enumerate(uv):
        np.testing.assert_array_equal(got[i], oracle.relight_lrgb(enc["coef"], enc["rgb"], scale, enc["bias"],
                                                                  float(u), float(v)))


def test_relight_rgb_format_matches_oracle_exactly(ctx, oracle, dome1, M60):
    """RGB-format PTMs (three coefficient planes, CLIP(POLY) per channel, rti-builder/ptmlib.c:610-619):
    the packed 4-pixel kernel (>= 4 directions), the scalar 4-pixel kernel (fewer) and the generic kernel
    (odd width) are all bit-exact against the restated loop."""
    rng = np.random.default_rng(9)
    for W, H in ((52, 35), (37, 9)):
        stack = oracle.synth_stack(23, 1, W, H, dome1)
        enc = oracle.encode_rgb(stack, M60)
        scale = oracle.header_roundtrip(enc["scale"])
        uv = np.concatenate([[[0, 0], [0.5, 0.5], [1, 0], [0, -1]],
                             rng.uniform(-0.9, 0.9, size=(9, 2))]).astype(np.float32)
        planes = [dev(enc["coef"][c]) for c in range(3)]
        for nd in (len(uv), 2):
            out = torch.empty((nd, H, W, 3), dtype=torch.uint8, device="cuda")
            ctx.relight_rgb(planes[0], planes[1], planes[2], W, H, scale, enc["bias"], uv[:nd], out)
            got = out.cpu().numpy()
            for i in range(nd):
                np.testing.assert_array_equal(got[i], oracle.relight_rgb(enc["coef"], scale, enc["bias"],
                                                                         float(uv[i, 0]), float(uv[i, 1])))


def test_clip_edge_values(ctx):
    """CLIP (rti-builder/ptmlib.h:156-158) at its edges through K2: extrema 0 and 256 give
    scale 1, bias 0, so bytes are clip(trunc(c))."""
    vals = np.array([0.0, 256.0, 255.0, 255.5, 254.999, 0.999, 1.0, 128.5, -0.0, 1e-30, 200.25, 77.75],
                    np.float32)
    c = np.tile(vals[:, None], (1, 6))
    P = len(vals)
    keys = torch.empty(12, dtype=torch.int32, device="cuda")
    out = torch.empty((P, 6), dtype=torch.uint8, device="cuda")
    sb = torch.empty(12, dtype=torch.int32, device="cuda")
    d = dev(c)
    ctx.keys_init(keys)
    ctx.minmax(d, P, keys)
    ctx.quantise(d, P, keys, out, sb)
    raw = sb.cpu().numpy()
    assert raw[:6].view(np.float32).tolist() == [1.0] * 6 and raw[6:].tolist() == [0] * 6
    want = [0, 255, 255, 255, 254, 0, 1, 128, 0, 0, 200, 77]
    assert out.cpu().numpy()[:, 0].tolist() == want


# ------------------------------------------------------------- host-buffer entry

def test_encode_host_matches_device_path(ctx, oracle, dome1, M60):
    stack = oracle.synth_stack(31, 0, 130, 75, dome1)
    want = gpu_encode_lrgb(ctx, stack, M60)
    got = ctx.encode_host(stack, M60, planes=1, want_coeffs=True)
    np.testing.assert_array_equal(got["blocks"][0], want["coef"])
    np.testing.assert_array_equal(got["blocks"][1], want["rgb"])
    np.testing.assert_array_equal(got["scale"], want["scale"])
    np.testing.assert_array_equal(got["bias"], want["bias"])
    np.testing.assert_array_equal(got["coeffs"].reshape(75, 130, 6), want["coeffs"])
    st = oracle.synth_stack(32, 1, 44, 30, dome1)
    w3 = gpu_encode_rgb(ctx, st, M60)
    g3 = ctx.encode_host(st, M60, planes=3)
    np.testing.assert_array_equal(np.stack(g3["blocks"]), w3["coef"])
    np.testing.assert_array_equal(g3["bias"], w3["bias"])


# ------------------------------------------------- full size, size-independent checks

def test_full_size_properties(ctx, oracle, dome1, M60):
    """BASELINE.json configs[1] size (6000 x 4000 x 60): properties that need no CPU run of
    24 MP -- slab invariance (any row-slab partition gives byte-identical results, which is
    also the multi-GPU claim), agreement of the extrema with a separate min/max pass, and the
    oracle on a few sampled rows."""
    from rti_b200.device import U8, key_to_float, light_q14
    from rti_b200.encoder import PtmEncoder, slab_rows
    W, H, n = 6000, 4000, 60
    P = W * H
    stack = torch.empty((n, H, W, 3), dtype=torch.uint8, device="cuda")
    ctx.synth(0x5EED0001, 0, U8, 0, P, light_q14(dome1), stack, P * 3)
    enc = PtmEncoder(W, H, M60, planes=1)
    enc.fit(stack)
    enc.exchange()
    enc.quantise()
    torch.cuda.synchronize()
    scale, bias = enc.scale_bias()
    # (0) the fused two-launch entry gives the same bytes, scale and bias, call after call
    fused = PtmEncoder(W, H, M60, planes=1)
    for _ in range(3):
        fused.encode(stack)
    torch.cuda.synchronize()
    assert torch.equal(fused.coef_bytes, enc.coef_bytes) and torch.equal(fused.rgb, enc.rgb)
    assert torch.equal(fused.sb, enc.sb)
    del fused
    # (1) extrema agree with an independent pass over the floats
    keys2 = torch.empty(12, dtype=torch.int32, device="cuda")
    ctx.keys_init(keys2)
    ctx.minmax(enc.coeffs, P, keys2)
    assert torch.equal(keys2, enc.keys)
    # (2) slab invariance: 3 uneven slabs, keys merged by integer max
    parts = []
    keys = []
    for r in range(3):
        r0, r1 = slab_rows(H, 3, r)
        e = PtmEncoder(W, r1 - r0, M60, planes=1)
        slab = stack[:, r0:r1].contiguous()
        e.fit(slab)
        keys.append(e.keys.clone())
        parts.append(e)
    merged = torch.stack(keys).max(0).values
    assert torch.equal(merged, enc.keys)
    for r, e in enumerate(parts):
        r0, r1 = slab_rows(H, 3, r)
        e.keys.copy_(merged)
        e.quantise()
        assert torch.equal(e.coef_bytes[0], enc.coef_bytes[0][r0 * W:r1 * W])
        assert torch.equal(e.rgb, enc.rgb[r0 * W:r1 * W])
    # (3) oracle on sampled rows
    rows = [0, 1234, 3999]
    sub = stack[:, rows].cpu().numpy()
    want_c = oracle.fit_u8(sub, M60)
    want_rgb = oracle.chroma_avg(sub)
    oracle.lrgb_colour_normalise(want_rgb, want_c)
    got_c = enc.coeffs[0].reshape(H, W, 6)[rows].cpu().numpy()
    got_rgb = enc.rgb.reshape(H, W, 3)[rows].cpu().numpy()
    np.testing.assert_array_equal(got_rgb, want_rgb)
    assert coeff_error(got_c, want_c, M60, sub[..., 0], lrgb=True) <= 1e-5
    inv = (np.float32(1.0) / scale).astype(np.float32)
    q = (want_c * inv).astype(np.float32) + bias.astype(np.float32)
    want_b = np.where(q > 255, 255, np.where(q < 0, 0, q)).astype(np.uint8)
    assert_bytes_close(enc.coef_bytes[0].reshape(H, W, 6)[rows].cpu().numpy(), want_b, "24 MP sampled rows")


def test_consuming_quantise_gives_identical_bytes(ctx, oracle, dome1, M60):
    """ptmb_quantise_consume (mirrored walk + L2 discard) must produce the same bytes as ptmb_quantise."""
    stack = oracle.synth_stack(41, 0, 512, 300, dome1)     # pixels % 16 == 0 -> the consuming kernel runs
    P = 512 * 300
    ctx.set_matrix(M60)
    coeffs = torch.empty((P, 6), dtype=torch.float32, device="cuda")
    rgb = torch.empty((P, 3), dtype=torch.uint8, device="cuda")
    keys = torch.empty(12, dtype=torch.int32, device="cuda")
    a = torch.empty((P, 6), dtype=torch.uint8, device="cuda")
    b = torch.empty((P, 6), dtype=torch.uint8, device="cuda")
    ctx.keys_init(keys)
    ctx.fit_lrgb(dev(stack), P * 3, P, coeffs, rgb, keys)
    ctx.quantise(coeffs, P, keys, a)
    ctx.quantise(coeffs, P, keys, b, consume=True)
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    # odd sizes fall back to the forward kernel and still agree
    P2 = 333 * 7
    c2 = torch.randn((P2, 6), device="cuda") * 100
    ctx.keys_init(keys)
    ctx.minmax(c2, P2, keys)
    a2 = torch.empty((P2, 6), dtype=torch.uint8, device="cuda")
    b2 = torch.empty((P2, 6), dtype=torch.uint8, device="cuda")
    ctx.quantise(c2, P2, keys, a2)
    ctx.quantise(c2.clone(), P2, keys, b2, consume=True)
    assert torch.equal(a2, b2)


def test_relight_shortcuts_are_exact_for_every_finite_float(ctx):
    """x / 255.0f as multiply + two FMAs, and CLIP as min/max + add-round-toward-zero, compared with
    the IEEE division and the reference's CLIP for every finite float (both signs, denormals
    included).  Only +-inf differs (NaN instead of inf)."""
    assert ctx.selftest_div255(0x00000000, 0x7F800000) == 0
    assert ctx.selftest_div255(0x80000000, 0x7F800000) == 0
    assert ctx.selftest_div255(0x7F800000, 1) == 1


def test_relight_sweep_360_directions(ctx, oracle, dome1, M60):
    """BASELINE.json configs[4] shape at small size: 360 directions on a ring of radius 0.7,
    batched in one call; every raster equals the oracle."""
    W, H = 64, 36
    stack = oracle.synth_stack(51, 0, W, H, dome1)
    enc = oracle.encode_lrgb(stack, M60)
    scale = oracle.header_roundtrip(enc["scale"])
    th = np.deg2rad(np.arange(360, dtype=np.float64))
    uv = np.stack([0.7 * np.cos(th), 0.7 * np.sin(th)], 1).astype(np.float32)
    out = torch.empty((360, H, W, 3), dtype=torch.uint8, device="cuda")
    ctx.relight_lrgb(dev(enc["coef"]), dev(enc["rgb"]), W, H, scale, enc["bias"], uv, out)
    got = out.cpu().numpy()
    for i in range(0, 360, 7):
        np.testing.assert_array_equal(got[i], oracle.relight_lrgb(enc["coef"], enc["rgb"], scale, enc["bias"],
                                                                  float(uv[i, 0]), float(uv[i, 1])))
    # odd width -> generic kernel, same answer
    W2 = 37
    st2 = oracle.synth_stack(52, 0, W2, 9, dome1)
    e2 = oracle.encode_lrgb(st2, M60)
    out2 = torch.empty((3, 9, W2, 3), dtype=torch.uint8, device="cuda")
    ctx.relight_lrgb(dev(e2["coef"]), dev(e2["rgb"]), W2, 9, e2["scale"], e2["bias"], uv[:3], out2)
    for i in range(3):
        np.testing.assert_array_equal(out2.cpu().numpy()[i],
                                      oracle.relight_lrgb(e2["coef"], e2["rgb"], e2["scale"], e2["bias"],
                                                          float(uv[i, 0]), float(uv[i, 1])))


def test_sixteen_bit_stacks(ctx, oracle):
    """BASELINE.json configs[3] shape at small size: 100 lights, 16-bit linear samples, RGB-format
    PTM (reference semantics: ptm_fit_poly_uint + ptm_scale_coefficients) and LRGB (extension, no
    reference counterpart: checked against the oracle's statement of the same semantics)."""
    from rti_b200 import lights
    from rti_b200.device import U16
    L = lights.dome100_lights()
    M = oracle.pinv(L)
    _sixteen_bit_case(ctx, oracle, L, M, 40, 26)      # 1040 px: one streaming tile + a 144 px partial tile
    _sixteen_bit_case(ctx, oracle, L, M, 37, 29)      # 1073 px: tile + generic remainder
    _sixteen_bit_case(ctx, oracle, L, M, 24, 10)      # 240 px: partial tile only


def _sixteen_bit_case(ctx, oracle, L, M, W, H):
    from rti_b200.device import U16
    P = W * H
    ctx.set_matrix(M)
    keys = torch.empty(12, dtype=torch.int32, device="cuda")
    sb = torch.empty(12, dtype=torch.int32, device="cuda")
    # RGB format
    s16 = oracle.synth_stack(0x5EED0004, 1, W, H, L, dtype=np.uint16)
    want = oracle.encode_rgb_u16(s16, M)
    d16 = dev(s16.view(np.int16)).view(torch.uint16)
    coeffs = torch.empty((3, P, 6), dtype=torch.float32, device="cuda")
    coef = torch.empty((3, P, 6), dtype=torch.uint8, device="cuda")
    ctx.keys_init(keys)
    ctx.fit_rgb(d16, P * 6, P, coeffs, P * 6, keys, U16)
    for b in range(3):
        ctx.quantise(coeffs[b], P, keys, coef[b], sb if b == 0 else None)
    raw = sb.cpu().numpy()
    for ch in range(3):
        assert coeff_error(coeffs[ch].cpu().numpy().reshape(H, W, 6), want["coeffs"][ch], M, s16[..., ch]) <= 1e-5
    np.testing.assert_allclose(raw[:6].view(np.float32), want["scale"], rtol=2e-6)
    assert_byte_parity(coef.cpu().numpy().reshape(3, H, W, 6), want["coef"], raw[6:], want["bias"],
                       "16-bit RGB coefficient bytes")
    # host-buffer entry with a 16-bit stack
    g = ctx.encode_host(s16, M, planes=3)
    np.testing.assert_array_equal(np.stack(g["blocks"]).reshape(3, P, 6), coef.cpu().numpy())
    # LRGB (extension)
    y16 = oracle.synth_stack(0x5EED0005, 0, W, H, L, dtype=np.uint16)
    want = oracle.encode_lrgb_u16(y16, M)
    c1 = torch.empty((P, 6), dtype=torch.float32, device="cuda")
    rgb = torch.empty((P, 3), dtype=torch.uint8, device="cuda")
    ctx.keys_init(keys)
    ctx.fit_lrgb(dev(y16.view(np.int16)).view(torch.uint16), P * 6, P, c1, rgb, keys, U16)
    np.testing.assert_array_equal(rgb.cpu().numpy().reshape(H, W, 3), want["rgb"])
    got = c1.cpu().numpy().reshape(H, W, 6)
    norm = 256.0 / np.floor(y16[..., 0].astype(np.float64).sum(0) / 100)
    bound = np.einsum("kn,nhw->hwk", np.abs(M).astype(np.float64), y16[..., 0].astype(np.float64)) * norm[..., None]
    assert (np.abs(got - want["coeffs"]) / bound).max() <= 1e-5


def test_full_size_sixteen_bit_100_lights(ctx, oracle):
    """BASELINE.json configs[3] at full size: 50 MP (8688 x 5755), 100 lights, uint16 (30 GB stack).
    Size-independent checks: the RGB-format encode of two uneven slabs with merged keys equals the
    whole-image encode byte for byte, and sampled rows agree with the oracle (ptm_fit_poly_uint
    semantics) -- for the LRGB extension too."""
    from rti_b200 import lights
    from rti_b200.device import U16, light_q14
    from rti_b200.encoder import PtmEncoder
    free, _ = torch.cuda.mem_get_info()
    if free < 60e9:
        pytest.skip("needs about 60 GB of free HBM")
    W, H, n = 8688, 5755, 100
    P = W * H
    L = lights.dome100_lights()
    M = oracle.pinv(L)
    stack = torch.empty((n, H, W, 3), dtype=torch.uint16, device="cuda")
    ctx.synth(0x5EED0004, 1, U16, 0, P, light_q14(L), stack, P * 6)
    enc = PtmEncoder(W, H, M, planes=3)
    enc.encode(stack)
    torch.cuda.synchronize()
    scale, bias = enc.scale_bias()
    rows = [0, 2871, H - 1]
    sub = np.stack([stack[:, r].view(torch.int16).cpu().numpy().view(np.uint16) for r in rows], axis=1)
    for ch in range(3):
        want = oracle.fit_u32(sub.astype(np.uint32), M, ch)
        got = enc.coeffs[ch].reshape(H, W, 6)[rows].cpu().numpy()
        assert coeff_error(got, want, M, sub[..., ch]) <= 1e-5
        inv = (np.float32(1.0) / scale).astype(np.float32)
        q = (want * inv).astype(np.float32) + bias.astype(np.float32)
        want_b = np.where(q > 255, 255, np.where(q < 0, 0, q)).astype(np.uint8)
        assert_bytes_close(enc.coef_bytes[ch].reshape(H, W, 6)[rows].cpu().numpy(), want_b, "50 MP u16 rows, plane %d" % ch)
    whole = enc.coef_bytes.clone()
    enc.fit(stack)                    # the fused encode hands its keys over on the device; the plain fit leaves them here
    keys_whole = enc.keys.clone()
    del enc
    # slab invariance (the multi-GPU claim) with uneven slabs
    cut = 2001
    parts = []
    for r0, r1 in ((0, cut), (cut, H)):
        e = PtmEncoder(W, r1 - r0, M, planes=3)
        e.fit(stack[:, r0:r1])        # a row range of a contiguous stack is NOT contiguous per light ...
        parts.append((r0, r1, e))
    merged = torch.stack([e.keys for _, _, e in parts]).max(0).values
    assert torch.equal(merged, keys_whole)
    for r0, r1, e in parts:
        e.keys.copy_(merged)
        e.quantise()
        for ch in range(3):
            assert torch.equal(e.coef_bytes[ch], whole[ch][r0 * W:r1 * W])


def test_full_size_relight_sweep(ctx, oracle, dome1, M60):
    """BASELINE.json configs[4] at full size: a 24 MP LRGB PTM relit for 360 directions in one call
    (26 GB of rasters).  Sampled rows of sampled directions equal the oracle exactly, and every
    raster equals the same direction relit on its own (batching does not change a byte)."""
    from rti_b200.device import U8, light_q14
    from rti_b200.encoder import PtmEncoder
    free, _ = torch.cuda.mem_get_info()
    if free < 40e9:
        pytest.skip("needs about 40 GB of free HBM")
    W, H = 6000, 4000
    P = W * H
    stack = torch.empty((60, H, W, 3), dtype=torch.uint8, device="cuda")
    ctx.synth(0x5EED0001, 0, U8, 0, P, light_q14(dome1), stack, P * 3)
    enc = PtmEncoder(W, H, M60)
    enc.encode(stack)
    torch.cuda.synchronize()
    del stack
    scale, bias = enc.scale_bias()
    scale = oracle.header_roundtrip(scale)
    th = np.deg2rad(np.arange(360, dtype=np.float64))
    uv = np.stack([0.7 * np.cos(th), 0.7 * np.sin(th)], 1).astype(np.float32)
    out = torch.empty((360, H, W, 3), dtype=torch.uint8, device="cuda")
    ctx.relight_lrgb(enc.coef_bytes[0], enc.rgb, W, H, scale, bias, uv, out)
    torch.cuda.synchronize()
    # stored rows for output rows y: H-1-y
    ys = [0, 1777, H - 1]
    coef = enc.coef_bytes[0].reshape(H, W, 6)
    rgb = enc.rgb.reshape(H, W, 3)
    for d in (0, 97, 359):
        for y in ys:
            c_row = coef[H - 1 - y].cpu().numpy()[None]
            r_row = rgb[H - 1 - y].cpu().numpy()[None]
            want = oracle.relight_lrgb(c_row, r_row, scale, bias, float(uv[d, 0]), float(uv[d, 1]))
            np.testing.assert_array_equal(out[d, y].cpu().numpy(), want[0])
    single = torch.empty((1, H, W, 3), dtype=torch.uint8, device="cuda")
    for d in (5, 200):
        ctx.relight_lrgb(enc.coef_bytes[0], enc.rgb, W, H, scale, bias, uv[d:d + 1], single)
        assert torch.equal(single[0], out[d])


@pytest.mark.parametrize("engine", ["ffma", "mma"])
@pytest.mark.parametrize("fmt", ["lrgb", "rgb"])
def test_config0_1024x1024_whole_image_vs_oracle(ctx, oracle, dome1, M60, engine, fmt):
    """BASELINE.json configs[0]: the 1024 x 1024 x dome1.lp stack, EVERY pixel against the oracle, both fit
    engines, both formats (the larger configurations are covered by sampled rows and invariances)."""
    W = H = 1024
    ctx.set_fit_engine(engine)
    try:
        if fmt == "lrgb":
            stack = oracle.synth_stack(0x5EED0001, 0, W, H, dome1)
            want = oracle.encode_lrgb(stack, M60)
            got = gpu_encode_lrgb(ctx, stack, M60)
            np.testing.assert_array_equal(got["rgb"], want["rgb"])
            err = coeff_error(got["coeffs"], want["coeffs"], M60, stack[..., 0], lrgb=True)
        else:
            stack = oracle.synth_stack(0x5EED0001, 1, W, H, dome1)
            want = oracle.encode_rgb(stack, M60)
            got = gpu_encode_rgb(ctx, stack, M60)
            err = max(coeff_error(got["coeffs"][ch], want["coeffs"][ch], M60, stack[..., ch]) for ch in range(3))
        print("1024x1024 %s/%s normalised coefficient error %.3g" % (fmt, engine, err))
        assert err <= 1e-5
        np.testing.assert_allclose(got["scale"], want["scale"], rtol=2e-6)
        assert_byte_parity(got["coef"], want["coef"], got["bias"], want["bias"], "1024x1024 %s/%s bytes" % (fmt, engine),
                           0.01)
    finally:
        ctx.set_fit_engine("default")


def test_light_count_limit_2048(ctx, oracle):
    """ptmb_set_matrix accepts up to 2048 lights; the pinv table of the generic kernels is then 48 KB + of
    dynamic shared memory (needs the opt-in attribute).  One more light is refused."""
    from rti_b200.device import PtmError
    n = 2048
    L = _ring_lights(n)
    M = oracle.pinv(L)
    W, H = 19, 7
    stack = oracle.synth_stack(0x4000, 0, W, H, L)
    want = oracle.encode_lrgb(stack, M)
    got = gpu_encode_lrgb(ctx, stack, M)
    np.testing.assert_array_equal(got["rgb"], want["rgb"])
    assert coeff_error(got["coeffs"], want["coeffs"], M, stack[..., 0], lrgb=True) <= 1e-5
    st3 = oracle.synth_stack(0x4001, 1, W, H, L)
    w3 = oracle.encode_rgb(st3, M)
    g3 = gpu_encode_rgb(ctx, st3, M)
    for ch in range(3):
        assert coeff_error(g3["coeffs"][ch], w3["coeffs"][ch], M, st3[..., ch]) <= 1e-5
    with pytest.raises(PtmError):
        ctx.set_matrix(np.zeros((6, 2049), np.float32))


def _ring_lights(n):
    """n light directions on three rings (any n >= 6 gives a well conditioned design matrix)."""
    th = np.arange(n) * (2 * np.pi / n) * 3.0
    r = np.array([0.25, 0.55, 0.85])[np.arange(n) % 3]
    u, v = r * np.cos(th), r * np.sin(th)
    return np.stack([u, v, np.sqrt(1 - u * u - v * v)], 1).astype(np.float32)


@pytest.mark.parametrize("n", [13, 100, 257, 300])
def test_light_counts_that_do_not_fill_the_ring_stages(ctx, oracle, n):
    """Light counts that are not a multiple of the ring's lights-per-stage (partial last stage),
    the largest count of the 16-bit-lane byte sums (257) and one beyond it (generic kernels)."""
    from rti_b200.device import U16
    L = _ring_lights(n)
    M = oracle.pinv(L)
    W, H = 72, 33                                    # 2376 px: full tiles + a partial tile
    P = W * H
    stack = oracle.synth_stack(0x1000 + n, 0, W, H, L)
    want = oracle.encode_lrgb(stack, M)
    got = gpu_encode_lrgb(ctx, stack, M)
    np.testing.assert_array_equal(got["rgb"], want["rgb"])
    assert coeff_error(got["coeffs"], want["coeffs"], M, stack[..., 0], lrgb=True) <= 1e-5
    # scale = (max - min) / 256 of float sums over n terms: the summation-order noise of the two
    # extrema grows with n and is divided by a range that can be small
    scale_tol = 3e-6 if n <= 100 else 3e-5
    np.testing.assert_allclose(got["scale"], want["scale"], rtol=scale_tol)
    assert_byte_parity(got["coef"], want["coef"], got["bias"], want["bias"], "LRGB bytes, %d lights" % n, 0.03)
    # RGB format, same light set
    st3 = oracle.synth_stack(0x2000 + n, 1, W, H, L)
    w3 = oracle.encode_rgb(st3, M)
    g3 = gpu_encode_rgb(ctx, st3, M)
    for ch in range(3):
        assert coeff_error(g3["coeffs"][ch], w3["coeffs"][ch], M, st3[..., ch]) <= 1e-5
    np.testing.assert_allclose(g3["scale"], w3["scale"], rtol=scale_tol)
    # 16-bit, same light set (stage size 4 there)
    s16 = oracle.synth_stack(0x3000 + n, 1, W, H, L, dtype=np.uint16)
    w16 = oracle.encode_rgb_u16(s16, M)
    ctx.set_matrix(M)
    coeffs = torch.empty((3, P, 6), dtype=torch.float32, device="cuda")
    keys = torch.empty(12, dtype=torch.int32, device="cuda")
    ctx.keys_init(keys)
    ctx.fit_rgb(dev(s16.view(np.int16)).view(torch.uint16), P * 6, P, coeffs, P * 6, keys, U16)
    for ch in range(3):
        assert coeff_error(coeffs[ch].cpu().numpy().reshape(H, W, 6), w16["coeffs"][ch], M, s16[..., ch]) <= 1e-5
