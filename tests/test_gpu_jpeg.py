"""GPU tests of the nvJPEG codec layer (ptmb_jpeg_*): decode against Pillow's libjpeg-turbo on the
same files (YCbCr without colour conversion, the way the reference asks libjpeg for it; RGB), the
upsampling restatement for 4:2:2 / 4:2:0, flipped decode straight into device memory, and encode.
The inverse DCT differs between the two decoders, so samples are compared within 2 LSB."""
import io

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
PIL = pytest.importorskip("PIL")
from PIL import Image  # noqa: E402


@pytest.fixture(scope="module")
def ctx():
    from rti_b200.device import Context
    torch.cuda.set_device(0)
    c = Context(0)
    yield c
    c.close()


def smooth_image(w, h, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.stack([128 + 100 * np.sin(x / 17.0 + seed) * np.cos(y / 23.0),
                    128 + 90 * np.cos(x / 29.0) * np.sin(y / 13.0 + 1),
                    128 + 80 * np.sin((x + y) / 31.0)], -1)
    img += rng.normal(scale=3.0, size=img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


def pil_jpeg(img, subsampling, quality=92):
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format="JPEG", quality=quality, subsampling=subsampling)
    return buf.getvalue()


def pil_decode_ycbcr(data):
    im = Image.open(io.BytesIO(data))
    im.draft("YCbCr", im.size)          # libjpeg out_color_space = JCS_YCbCr: no colour conversion
    assert im.mode == "YCbCr"
    return np.asarray(im)


@pytest.mark.parametrize("subsampling,name", [(0, "4:4:4"), (1, "4:2:2"), (2, "4:2:0")])
@pytest.mark.parametrize("w,h", [(64, 48), (67, 45), (130, 33)])
def test_decode_ycbcr_matches_libjpeg(ctx, subsampling, name, w, h):
    data = pil_jpeg(smooth_image(w, h, 3), subsampling)
    want = pil_decode_ycbcr(data)
    assert ctx.jpeg_info(data) == (w, h, 3)
    got = ctx.jpeg_decode(data, want_rgb=False)
    d = np.abs(got.astype(int) - want.astype(int))
    print("%s %dx%d: max |diff| per channel %s, mean %.3f" % (name, w, h, d.reshape(-1, 3).max(0), d.mean()))
    assert d.max() <= 2 and d.mean() < 0.25


def test_decode_rgb_gray_flip_and_device_output(ctx):
    img = smooth_image(96, 50, 5)
    data = pil_jpeg(img, 2)
    want = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
    got = ctx.jpeg_decode(data, want_rgb=True)
    assert np.abs(got.astype(int) - want.astype(int)).max() <= 3
    # bottom row first, written straight into device memory (one light's image of a stack)
    ycc = ctx.jpeg_decode(data)
    plane = torch.zeros((50, 96, 3), dtype=torch.uint8, device="cuda")
    ctx.jpeg_decode(data, flip=True, out=plane)
    np.testing.assert_array_equal(plane.cpu().numpy(), ycc[::-1])
    # grayscale file
    gbuf = io.BytesIO()
    Image.fromarray(img[..., 0]).save(gbuf, format="JPEG", quality=90)
    g = ctx.jpeg_decode(gbuf.getvalue())
    assert g.shape == (50, 96, 1)
    assert np.abs(g[..., 0].astype(int) - np.asarray(Image.open(io.BytesIO(gbuf.getvalue()))).astype(int)).max() <= 2


def psnr(a, b):
    mse = ((a.astype(np.float64) - b.astype(np.float64)) ** 2).mean()
    return 10 * np.log10(255.0 ** 2 / max(mse, 1e-12))


def test_encode_is_readable_by_libjpeg(ctx):
    img = smooth_image(120, 77, 9)
    data = ctx.jpeg_encode(img, quality=90)
    assert data[:2] == b"\xff\xd8"
    back = np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))
    assert back.shape == img.shape and psnr(back, img) > 30
    gray = ctx.jpeg_encode(img[..., 1], quality=90)
    gb = np.asarray(Image.open(io.BytesIO(gray)))
    assert gb.shape == img.shape[:2] and psnr(gb, img[..., 1]) > 33
    # YCbCr input (what a LUM relight hands to the writer) decodes back to the same YCbCr
    ycc = np.asarray(Image.fromarray(img).convert("YCbCr"))
    d2 = ctx.jpeg_encode(ycc, ycbcr=True, quality=90)
    assert psnr(pil_decode_ycbcr(d2), ycc) > 30
