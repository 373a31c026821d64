"""Tensor-core fit engine (rti_b200/csrc/ptm_fit_mma.cu) against the CPU oracle and against the
FFMA engine, through the C ABI.  The engine computes the exact integer dot product with the
pseudo-inverse in 32-bit fixed point and rounds once, so it must meet the same bars as the FFMA
kernels: float coefficients within 1e-5 of the coefficient scale, colour bytes identical,
coefficient bytes +-1 LSB on summation-order ties only.
"""
import numpy as np
import pytest

from conftest import assert_byte_parity
from test_gpu_parity import assert_bytes_close, coeff_error, dev, gpu_encode_lrgb, gpu_encode_rgb

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ctx():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from rti_b200.device import Context
    torch.cuda.set_device(0)
    c = Context(0, stream=torch.cuda.current_stream().cuda_stream)
    c.set_fit_engine("mma")
    yield c
    c.close()


@pytest.fixture(scope="module")
def ctx_ffma():
    from rti_b200.device import Context
    c = Context(0, stream=torch.cuda.current_stream().cuda_stream)
    c.set_fit_engine("ffma")
    yield c
    c.close()


def _ring_lights(n):
    th = np.arange(n) * 2.399963
    r = 0.15 + 0.8 * (np.arange(n) + 0.5) / n
    u, v = r * np.cos(th), r * np.sin(th)
    return np.stack([u, v, np.sqrt(1 - u * u - v * v)], 1).astype(np.float32)


# 128-pixel units: exactly one, one + tail of 16, many + tail, tail only, a width that leaves a
# remainder for the generic kernel (pixels % 16 != 0), a tail whose third column block is partial
# (112 x 2), and enough units for several ring wraps
SHAPES = [(128, 1), (48, 3), (16, 1), (64, 48), (72, 33), (37, 29), (80, 33), (112, 2), (512, 300)]


def eligible(W, H):
    """TMA needs a 16-byte image pitch; at least one 16-pixel-granular unit must exist."""
    return (W * H * 3) % 16 == 0 and W * H >= 16


@pytest.mark.parametrize("W,H", SHAPES)
def test_rgb_mma_vs_oracle(ctx, oracle, dome1, M60, W, H):
    stack = oracle.synth_stack(0x77 + W, 1, W, H, dome1)
    want = oracle.encode_rgb(stack, M60)
    before = ctx.mma_launches()
    got = gpu_encode_rgb(ctx, stack, M60)
    assert ctx.mma_launches() == before + (1 if eligible(W, H) else 0), "the tensor-core kernel did not run"
    for ch in range(3):
        err = coeff_error(got["coeffs"][ch], want["coeffs"][ch], M60, stack[..., ch])
        print("plane %d normalised coefficient error %.3g" % (ch, err))
        assert err <= 1e-5
    np.testing.assert_allclose(got["minmax"], want["minmax"], rtol=2e-6, atol=1e-4)
    np.testing.assert_allclose(got["scale"], want["scale"], rtol=2e-6)
    assert_byte_parity(got["coef"], want["coef"], got["bias"], want["bias"], "RGB coefficient bytes (mma) %dx%d" % (W, H))


@pytest.mark.parametrize("W,H", SHAPES)
def test_lrgb_mma_vs_oracle(ctx, oracle, dome1, M60, W, H):
    stack = oracle.synth_stack(0x99 + W, 0, W, H, dome1)
    want = oracle.encode_lrgb(stack, M60)
    before = ctx.mma_launches()
    got = gpu_encode_lrgb(ctx, stack, M60)
    assert ctx.mma_launches() == before + (1 if eligible(W, H) else 0), "the tensor-core kernel did not run"
    np.testing.assert_array_equal(got["rgb"], want["rgb"])
    err = coeff_error(got["coeffs"], want["coeffs"], M60, stack[..., 0], lrgb=True)
    print("normalised coefficient error %.3g" % err)
    assert err <= 1e-5
    np.testing.assert_allclose(got["minmax"], want["minmax"], rtol=2e-6, atol=1e-4)
    # scale = (max - min) / 256: the two extrema carry the ORACLE's summation-order noise (a few ulp of
    # the coefficient scale, which is larger than the range on small images); the engine itself rounds once
    np.testing.assert_allclose(got["scale"], want["scale"], rtol=5e-6)
    assert_byte_parity(got["coef"], want["coef"], got["bias"], want["bias"], "LRGB coefficient bytes (mma) %dx%d" % (W, H))


@pytest.mark.parametrize("n", [6, 8, 12, 32, 33, 100, 128, 160, 192, 224])
def test_mma_light_counts(ctx, oracle, n):
    """K padding (lights rounded up to 32), one to seven MMA steps per column block; light groups
    of 8 with and without a remainder group."""
    L = _ring_lights(n)
    M = oracle.pinv(L)
    W, H = 80, 33                                     # 2640 px: 16-byte image pitch, 20 units + a tail
    assert eligible(W, H)
    stack = oracle.synth_stack(0x4000 + n, 0, W, H, L)
    want = oracle.encode_lrgb(stack, M)
    before = ctx.mma_launches()
    got = gpu_encode_lrgb(ctx, stack, M)
    assert ctx.mma_launches() == before + 1
    np.testing.assert_array_equal(got["rgb"], want["rgb"])
    assert coeff_error(got["coeffs"], want["coeffs"], M, stack[..., 0], lrgb=True) <= 1e-5
    st3 = oracle.synth_stack(0x5000 + n, 1, W, H, L)
    w3 = oracle.encode_rgb(st3, M)
    g3 = gpu_encode_rgb(ctx, st3, M)
    assert ctx.mma_launches() == before + 2
    for ch in range(3):
        assert coeff_error(g3["coeffs"][ch], w3["coeffs"][ch], M, st3[..., ch]) <= 1e-5


def test_mma_is_the_correctly_rounded_dot_product(ctx, oracle, dome1, M60):
    """The engine's float equals the float64 dot product with M, rounded once, up to the 2^-31
    quantisation of M: far inside one float ulp of the coefficient scale."""
    W, H = 128, 8
    stack = oracle.synth_stack(0xABCD, 1, W, H, dome1)
    got = gpu_encode_rgb(ctx, stack, M60)
    exact = np.einsum("kn,nhwc->chwk", M60.astype(np.float64), stack.astype(np.float64))
    bound = np.einsum("kn,nhwc->chwk", np.abs(M60).astype(np.float64), stack.astype(np.float64))
    err = np.abs(got["coeffs"].astype(np.float64) - exact)
    ulp = np.spacing(np.abs(exact).astype(np.float32)).astype(np.float64)
    assert (err <= 0.5 * ulp + 2.0 ** -30 * bound).all()


def test_mma_matches_ffma_engine_at_scale(ctx, ctx_ffma, oracle, dome1, M60):
    """3 MP slab, both engines: colour bytes identical, floats within the bar, fused device encode
    (K1 -> hand-off -> K2) gives bytes within 1 LSB."""
    from rti_b200 import lights
    from rti_b200.device import NCOEF, U8, light_q14
    W, H = 2000, 1504
    P = W * H
    n = len(dome1)
    stack = torch.empty((n, P, 3), dtype=torch.uint8, device="cuda")
    ctx.synth(0x5EED0001, 0, U8, 0, P, light_q14(dome1), stack, P * 3)
    outs = []
    for c in (ctx, ctx_ffma):
        c.set_matrix(M60)
        coeffs = torch.empty((P, NCOEF), dtype=torch.float32, device="cuda")
        rgb = torch.empty((P, 3), dtype=torch.uint8, device="cuda")
        coef = torch.empty((P, NCOEF), dtype=torch.uint8, device="cuda")
        sb = torch.empty(12, dtype=torch.int32, device="cuda")
        c.encode_lrgb_dev(stack, P * 3, P, coeffs, rgb, coef, sb)
        torch.cuda.synchronize()
        outs.append((coeffs, rgb, coef, sb.cpu().numpy()))
    assert torch.equal(outs[0][1], outs[1][1])
    scale = np.abs(outs[1][0].cpu().numpy()).max(0)
    d = (outs[0][0] - outs[1][0]).abs().amax(0).cpu().numpy()
    print("max |mma - ffma| per coefficient, relative to its range:", d / scale)
    assert (d <= 1e-5 * scale).all()
    np.testing.assert_allclose(outs[0][3][:6].view(np.float32), outs[1][3][:6].view(np.float32), rtol=5e-6)
    assert_byte_parity(outs[0][2].cpu().numpy(), outs[1][2].cpu().numpy(), outs[0][3][6:], outs[1][3][6:],
                       "bytes between engines", 0.01)


@pytest.mark.parametrize("W,H,n", [(64, 1, 60), (80, 33, 60), (72, 33, 100), (512, 150, 100), (48, 5, 224)])
def test_rgb_u16_mma_vs_oracle(ctx, oracle, W, H, n):
    """16-bit stacks (ptm_fit_poly_uint semantics): low and high byte rows combined in integers."""
    from rti_b200.device import U16
    L = _ring_lights(n)
    M = oracle.pinv(L)
    P = W * H
    s16 = oracle.synth_stack(0x6000 + n + W, 1, W, H, L, dtype=np.uint16)
    want = oracle.encode_rgb_u16(s16, M)
    ctx.set_matrix(M)
    coeffs = torch.empty((3, P, 6), dtype=torch.float32, device="cuda")
    keys = torch.empty(12, dtype=torch.int32, device="cuda")
    ctx.keys_init(keys)
    before = ctx.mma_launches()
    ctx.fit_rgb(dev(s16.view(np.int16)).view(torch.uint16), P * 6, P, coeffs, P * 6, keys, U16)
    torch.cuda.synchronize()
    assert ctx.mma_launches() == before + (1 if (P * 6) % 16 == 0 and P >= 16 else 0)
    got = coeffs.cpu().numpy().reshape(3, H, W, 6)
    for ch in range(3):
        err = coeff_error(got[ch], want["coeffs"][ch], M, s16[..., ch])
        print("plane %d normalised coefficient error %.3g" % (ch, err))
        assert err <= 1e-5
    from rti_b200.device import key_to_float
    f = key_to_float(keys.cpu().numpy())
    mm = np.concatenate([-f[:6], f[6:]])
    p0 = want["coeffs"][0].reshape(-1, 6)
    np.testing.assert_allclose(mm, np.concatenate([p0.min(0), p0.max(0)]), rtol=3e-5, atol=1e-2)


@pytest.mark.parametrize("W,H,n", [(64, 1, 60), (80, 33, 60), (72, 33, 100), (512, 150, 100), (48, 5, 224)])
def test_lrgb_u16_mma_vs_oracle(ctx, oracle, W, H, n):
    """LRGB on 16-bit stacks (the extension stated in oracle/ptm_oracle.c:orc_lrgb16_finish): one pixel per
    (low byte, high byte) lane pair, 16-bit means exchanged through shared memory."""
    from rti_b200.device import U16
    L = _ring_lights(n)
    M = oracle.pinv(L)
    P = W * H
    y16 = oracle.synth_stack(0x7000 + n + W, 0, W, H, L, dtype=np.uint16)
    want = oracle.encode_lrgb_u16(y16, M)
    ctx.set_matrix(M)
    c1 = torch.empty((P, 6), dtype=torch.float32, device="cuda")
    rgb = torch.empty((P, 3), dtype=torch.uint8, device="cuda")
    keys = torch.empty(12, dtype=torch.int32, device="cuda")
    ctx.keys_init(keys)
    before = ctx.mma_launches()
    ctx.fit_lrgb(dev(y16.view(np.int16)).view(torch.uint16), P * 6, P, c1, rgb, keys, U16)
    torch.cuda.synchronize()
    assert ctx.mma_launches() == before + (1 if (P * 6) % 16 == 0 and P >= 16 else 0)
    np.testing.assert_array_equal(rgb.cpu().numpy().reshape(H, W, 3), want["rgb"])
    got = c1.cpu().numpy().reshape(H, W, 6)
    norm = 256.0 / np.floor(y16[..., 0].astype(np.float64).sum(0) / n)
    bound = np.einsum("kn,nhw->hwk", np.abs(M).astype(np.float64), y16[..., 0].astype(np.float64)) * norm[..., None]
    err = (np.abs(got - want["coeffs"]) / bound).max()
    print("normalised coefficient error %.3g" % err)
    assert err <= 1e-5
    from rti_b200.device import key_to_float
    f = key_to_float(keys.cpu().numpy())
    flat = want["coeffs"].reshape(-1, 6)
    np.testing.assert_allclose(np.concatenate([-f[:6], f[6:]]), np.concatenate([flat.min(0), flat.max(0)]),
                               rtol=3e-5, atol=1e-2)
