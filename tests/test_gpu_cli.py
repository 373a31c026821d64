"""GPU tests of the three command lines (rti_b200/bin/ptm-encoder, ptm-decoder, ptm-exploder)
against the reference's UNMODIFIED mains (oracle/_ref/ptm-*-ref, built from /root/reference by
oracle/Makefile; both sides use the same lossless raw image container, so files compare
byte for byte where the arithmetic is exact)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, assert_byte_parity

pytestmark = pytest.mark.gpu

BIN = os.path.join(ROOT, "rti_b200", "bin")
REF = os.path.join(ROOT, "oracle", "_ref")


def have_ref():
    return all(os.path.exists(os.path.join(REF, "ptm-%s-ref" % n)) for n in ("encoder", "decoder", "exploder"))


def write_inputs(oracle, d, stack_stored, lights):
    names = []
    for i in range(stack_stored.shape[0]):
        name = "img%03d.rawj" % i
        oracle.write_rawj(os.path.join(d, name), stack_stored[i, ::-1])   # files are top row first
        names.append(name)
    lp = os.path.join(d, "sample.lp")
    with open(lp, "w") as fp:
        fp.write("%d\n" % len(names))
        for name, (u, v, w) in zip(names, lights):
            fp.write("%s %.9g %.9g %.9g\n" % (name, u, v, w))
    return lp


def run(cmd, codec="raw", env=None, **kw):
    # PTM_IMAGE_CODEC=raw: the lossless container on both sides, so that files compare byte for byte
    return subprocess.run(cmd, capture_output=True,
                          env=dict(os.environ, OPENBLAS_NUM_THREADS="1", PTM_IMAGE_CODEC=codec, **(env or {})), **kw)


@pytest.mark.parametrize("fmt,mode", [("PTM_FORMAT_LRGB", 0), ("PTM_FORMAT_RGB", 1),
                                      ("PTM_FORMAT_JPEG_LRGB", 0), ("PTM_FORMAT_JPEG_RGB", 1)])
def test_encoder_decoder_vs_reference_mains(oracle, dome1, tmp_path, fmt, mode):
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    d = str(tmp_path)
    W, H = 72, 40
    stack = oracle.synth_stack(0xC11, mode, W, H, dome1)
    lp = write_inputs(oracle, d, stack, dome1)
    mine, ref = os.path.join(d, "mine.ptm"), os.path.join(d, "ref.ptm")
    r = run([os.path.join(BIN, "ptm-encoder"), "-f", fmt, "-o", mine, "-v", lp])
    assert r.returncode == 0, r.stderr.decode()
    assert b"60 JPEGs opened" in r.stderr and b"done SVD" in r.stderr
    r = run([os.path.join(REF, "ptm-encoder-ref"), "-f", fmt, "-o", ref, lp])
    assert r.returncode == 0, r.stderr.decode()
    a, b = open(mine, "rb").read(), open(ref, "rb").read()
    assert len(a) == len(b)
    # text header: same lines; scale within float noise, bias within 1
    def header_lines(blob, n):
        return blob.split(b"\n", n)[:n]
    nl = 6 if "JPEG" not in fmt else 14
    ha, hb = header_lines(a, nl), header_lines(b, nl)
    assert ha[:4] == hb[:4]
    np.testing.assert_allclose([float(x) for x in ha[4].split()], [float(x) for x in hb[4].split()], rtol=3e-6, atol=2e-6)
    assert max(abs(int(x) - int(y)) for x, y in zip(ha[5].split(), hb[5].split())) <= 1
    assert ha[6:] == hb[6:]
    body_a = np.frombuffer(a, np.uint8)[-(len(a) - len(b"\n".join(ha)) - 1):]
    body_b = np.frombuffer(b, np.uint8)[-(len(b) - len(b"\n".join(hb)) - 1):]
    if "JPEG" not in fmt:
        # coefficient blocks first ([pixels][6] each), then (LRGB) the colour block, which must be identical
        ncoef = body_a.size * 6 // 9 if "LRGB" in fmt else body_a.size
        assert_byte_parity(body_a[:ncoef].reshape(-1, 6), body_b[:ncoef].reshape(-1, 6),
                           [int(x) for x in ha[5].split()], [int(x) for x in hb[5].split()], fmt + " payload", 0.01)
        assert np.array_equal(body_a[ncoef:], body_b[ncoef:]), "colour block differs"
    elif ha[5] == hb[5]:
        # compressed planes: byte streams of (lossless, in this test) coded planes of equal length
        diff = np.abs(body_a.astype(int) - body_b.astype(int))
        print("%s: %d of %d payload bytes differ" % (fmt, (diff != 0).sum(), diff.size))
        assert diff.max() <= 1 and (diff != 0).mean() < 0.01
    # decoders: each on the REFERENCE's file must agree exactly (relight is bit exact),
    for uv in (("0.25", "-0.5"), ()):
        ra = run([os.path.join(BIN, "ptm-decoder"), ref, *uv])
        rb = run([os.path.join(REF, "ptm-decoder-ref"), ref, *uv])
        assert ra.returncode == 0 and rb.returncode == 0, ra.stderr.decode()
        assert ra.stdout == rb.stdout
    # and my decoder reads my own file
    ra = run([os.path.join(BIN, "ptm-decoder"), mine, "0.1", "0.2"])
    rb = run([os.path.join(REF, "ptm-decoder-ref"), mine, "0.1", "0.2"])
    assert ra.returncode == 0 and ra.stdout == rb.stdout


def test_exploder_vs_reference_main(oracle, dome1, tmp_path):
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    d = str(tmp_path)
    stack = oracle.synth_stack(0xC12, 0, 48, 36, dome1)
    lp = write_inputs(oracle, d, stack, dome1)
    ptm = os.path.join(d, "x.ptm")
    assert run([os.path.join(REF, "ptm-encoder-ref"), "-f", "PTM_FORMAT_LRGB", "-o", ptm, lp]).returncode == 0
    os.makedirs(os.path.join(d, "a"))
    os.makedirs(os.path.join(d, "b"))
    ra = run([os.path.join(BIN, "ptm-exploder"), ptm, lp, os.path.join(d, "a", "out.jpg")])
    rb = run([os.path.join(REF, "ptm-exploder-ref"), ptm, lp, os.path.join(d, "b", "out.jpg")])
    assert ra.returncode == 0 and rb.returncode == 0, ra.stderr.decode()
    fa, fb = sorted(os.listdir(os.path.join(d, "a"))), sorted(os.listdir(os.path.join(d, "b")))
    assert fa == fb and len(fa) == 60 and fa[0] == "out001.jpeg"     # ref :71-83 naming, no dash
    for f in fa:
        assert open(os.path.join(d, "a", f), "rb").read() == open(os.path.join(d, "b", f), "rb").read(), f
    assert ra.stderr.replace(b"/a/", b"/b/") == rb.stderr             # the progress lines


def test_cli_errors_match_reference(tmp_path):
    enc = os.path.join(BIN, "ptm-encoder")
    r = run([enc, "-l"])
    assert r.returncode == 0 and r.stdout.split() == [b"Supported", b"formats:", b"PTM_FORMAT_RGB", b"PTM_FORMAT_JPEG_RGB",
                                                      b"PTM_FORMAT_LRGB", b"PTM_FORMAT_JPEG_LRGB", b"PTM_FORMAT_LUM"]
    r = run([enc, "-f", "NOPE", "x.lp"])
    assert r.returncode == 0 and b"No format by that name: NOPE" in r.stderr
    r = run([enc, os.path.join(str(tmp_path), "missing.lp")])
    assert r.returncode == 1 and b"can't open" in r.stderr
    lp = tmp_path / "few.lp"
    lp.write_text("0\n")
    r = run([enc, str(lp)])
    assert r.returncode == 1 and b"not enough jpegs 0 < 12" in r.stderr
    r = run([os.path.join(BIN, "ptm-decoder")])
    assert r.returncode == 1 and b"Usage:" in r.stderr
    bad = tmp_path / "bad.ptm"
    bad.write_text("hello\n")
    r = run([os.path.join(BIN, "ptm-decoder"), str(bad)])
    assert r.returncode == 1 and b"not a PTM file" in r.stderr
    r = run([os.path.join(BIN, "ptm-exploder"), "a", "b"])
    assert r.returncode == 1 and b"Usage:" in r.stderr


def test_real_jpeg_files_end_to_end(oracle, dome1, tmp_path):
    """Real JPEG inputs and outputs through nvJPEG: Pillow writes the 60 input photographs (4:2:0),
    ptm-encoder reads them, ptm-decoder writes a JPEG that Pillow reads back.  Compared with the
    oracle pipeline run on Pillow's own (libjpeg) YCbCr decode of the same files -- the two JPEG
    decoders differ by an LSB here and there, so this checks agreement, not identity."""
    import io
    from PIL import Image
    d = str(tmp_path)
    W, H = 96, 64
    stack = oracle.synth_stack(0xC13, 0, W, H, dome1)          # YCbCr triplets, stored row order
    names, decoded = [], []
    for i in range(60):
        ycc_top_first = stack[i, ::-1]
        im = Image.fromarray(np.ascontiguousarray(ycc_top_first), mode="YCbCr")
        name = "img%03d.jpg" % i
        im.save(os.path.join(d, name), format="JPEG", quality=95, subsampling=2)
        back = Image.open(os.path.join(d, name))
        back.draft("YCbCr", back.size)
        decoded.append(np.asarray(back)[::-1])                   # what libjpeg would put into the stack
        names.append(name)
    lp = os.path.join(d, "sample.lp")
    with open(lp, "w") as fp:
        fp.write("60\n")
        for name, (u, v, w) in zip(names, dome1):
            fp.write("%s %.9g %.9g %.9g\n" % (name, u, v, w))
    ptm = os.path.join(d, "out.ptm")
    r = run([os.path.join(BIN, "ptm-encoder"), "-f", "PTM_FORMAT_LRGB", "-o", ptm, lp], codec="jpeg")
    assert r.returncode == 0, r.stderr.decode()
    got = oracle.parse_ptm(open(ptm, "rb").read())
    want = oracle.encode_lrgb(np.ascontiguousarray(np.stack(decoded)), oracle.pinv(dome1))
    np.testing.assert_allclose(got["scale"], want["scale"], rtol=0.05)
    rgb_diff = np.abs(got["blocks"][1].astype(int) - want["rgb"].astype(int))
    assert rgb_diff.max() <= 3 and rgb_diff.mean() < 0.3
    r = run([os.path.join(BIN, "ptm-decoder"), ptm, "0.2", "0.1"], codec="jpeg")
    assert r.returncode == 0 and r.stdout[:2] == b"\xff\xd8", r.stderr.decode()
    relit = np.asarray(Image.open(io.BytesIO(r.stdout)).convert("RGB")).astype(np.float64)
    ref = oracle.relight_lrgb(want["coef"], want["rgb"], oracle.header_roundtrip(want["scale"]), want["bias"], 0.2, 0.1)
    mse = ((relit - ref.astype(np.float64)) ** 2).mean()
    psnr = 10 * np.log10(255.0 ** 2 / max(mse, 1e-9))
    print("relit JPEG vs oracle raster: PSNR %.1f dB" % psnr)
    assert psnr > 28
    # the JPEG-compressed PTM format really holds JPEG streams now, and our decoder reads it back
    ptmj = os.path.join(d, "outj.ptm")
    r = run([os.path.join(BIN, "ptm-encoder"), "-f", "PTM_FORMAT_JPEG_LRGB", "-o", ptmj, lp], codec="jpeg")
    assert r.returncode == 0 and b"\xff\xd8\xff" in open(ptmj, "rb").read()
    assert os.path.getsize(ptmj) < os.path.getsize(ptm)
    r2 = run([os.path.join(BIN, "ptm-decoder"), ptmj, "0.2", "0.1"], codec="raw")
    assert r2.returncode == 0
    relit2 = oracle.read_rawj(r2.stdout).astype(np.float64)
    mse2 = ((relit2 - ref.astype(np.float64)) ** 2).mean()
    assert 10 * np.log10(255.0 ** 2 / max(mse2, 1e-9)) > 24


def have_dropin():
    return all(os.path.exists(os.path.join(REF, "ptm-%s-dropin" % n)) for n in ("encoder", "decoder", "exploder"))


@pytest.mark.parametrize("fmt,mode", [("PTM_FORMAT_LRGB", 0), ("PTM_FORMAT_RGB", 1), ("PTM_FORMAT_JPEG_LRGB", 0)])
def test_reference_mains_linked_against_our_library(oracle, dome1, tmp_path, fmt, mode):
    """The drop-in claim itself: the reference's UNMODIFIED ptm-encoder.c / ptm-decoder.c /
    ptm-exploder.c, compiled against the reference's own ptmlib.h, but linked against OUR libptm.so
    instead of ptmlib.o (oracle/Makefile, target dropin).  Every ptm_* call of those mains then runs on
    the GPU; the output must match the all-reference build within the usual +-1 LSB and must be
    IDENTICAL to our own command line (same kernels, same floats)."""
    if not (have_ref() and have_dropin()):
        pytest.skip("oracle/_ref drop-in mains not built")
    d = str(tmp_path)
    W, H = 80, 44
    stack = oracle.synth_stack(0xD40, mode, W, H, dome1)
    lp = write_inputs(oracle, d, stack, dome1)
    outs = {}
    # "ours" is pinned to the FFMA engine here: the reference main fits channel by channel through
    # ptm_fit_poly_jsample (FFMA kernels), our own main fuses the RGB planes on the tensor-core engine by
    # default, and the two engines round differently (both inside the +-1 LSB bar, checked below)
    for tag, exe, env in (("ref", os.path.join(REF, "ptm-encoder-ref"), None),
                          ("dropin", os.path.join(REF, "ptm-encoder-dropin"), None),
                          ("ours", os.path.join(BIN, "ptm-encoder"), {"PTM_FIT_ENGINE": "ffma"}),
                          ("ours_default", os.path.join(BIN, "ptm-encoder"), None)):
        path = os.path.join(d, tag + ".ptm")
        r = run([exe, "-f", fmt, "-o", path, lp], env=env)
        assert r.returncode == 0, (tag, r.stderr.decode())
        outs[tag] = open(path, "rb").read()
    assert outs["dropin"] == outs["ours"], "reference main on our library differs from our own main"
    if "JPEG" not in fmt:
        o, q = np.frombuffer(outs["ours_default"], np.uint8), np.frombuffer(outs["ours"], np.uint8)
        h1, h2 = outs["ours_default"].split(b"\n", 6)[:6], outs["ours"].split(b"\n", 6)[:6]
        assert o.size == q.size and h1[:4] == h2[:4]
        if h1[4:6] == h2[4:6]:
            dd = np.abs(o.astype(int) - q.astype(int))
            assert dd.max() <= 1 and (dd != 0).mean() < 0.01
    a = np.frombuffer(outs["dropin"], np.uint8)
    b = np.frombuffer(outs["ref"], np.uint8)
    assert a.size == b.size
    nl = 6 if "JPEG" not in fmt else 14
    ha, hb = outs["dropin"].split(b"\n", nl)[:nl], outs["ref"].split(b"\n", nl)[:nl]
    assert ha[:4] == hb[:4] and ha[6:] == hb[6:]
    if ha[4:6] == hb[4:6]:
        diff = np.abs(a.astype(int) - b.astype(int))
        assert diff.max() <= 1 and (diff != 0).mean() < 0.01
    # decoder and exploder mains of the reference on our library: bit-identical rasters
    ref_ptm = os.path.join(d, "ref.ptm")
    ra = run([os.path.join(REF, "ptm-decoder-dropin"), ref_ptm, "0.4", "-0.3"])
    rb = run([os.path.join(REF, "ptm-decoder-ref"), ref_ptm, "0.4", "-0.3"])
    assert ra.returncode == 0 and rb.returncode == 0, ra.stderr.decode()
    assert ra.stdout == rb.stdout
    if fmt == "PTM_FORMAT_LRGB":
        os.makedirs(os.path.join(d, "a"))
        os.makedirs(os.path.join(d, "b"))
        ra = run([os.path.join(REF, "ptm-exploder-dropin"), ref_ptm, lp, os.path.join(d, "a", "o.jpg")])
        rb = run([os.path.join(REF, "ptm-exploder-ref"), ref_ptm, lp, os.path.join(d, "b", "o.jpg")])
        assert ra.returncode == 0 and rb.returncode == 0
        for f in sorted(os.listdir(os.path.join(d, "b")))[::7]:
            assert open(os.path.join(d, "a", f), "rb").read() == open(os.path.join(d, "b", f), "rb").read(), f


def test_jpeg_ptm_plane_prediction_and_side_info(oracle, tmp_path):
    """A PTM_FORMAT_JPEG_LRGB file that uses everything the reference's reader supports beyond what its writer
    emits (rti-builder/ptmlib.c:249-284,421-440): a non-identity plane order, planes predicted from other planes
    (reference_planes >= 0), one of them from the INVERTED reference (transforms bit 0), and side-info patches
    (5-byte records, one with an out-of-range offset that must be ignored).  Decoded by libptm.so's reader and by
    the reference's unmodified decoder: the relit rasters are byte-identical, and equal the oracle's relight of
    the planes reconstructed independently here."""
    import struct
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(11)
    W, H, n = 24, 16, 9
    P = W * H
    stored = rng.integers(0, 256, (n, P), dtype=np.uint8)            # what the streams hold
    order = [3, 0, 1, 2, 5, 4, 8, 6, 7]                              # processing rank of each component
    ref_planes = [-1, 0, -1, -1, 1, -1, -1, 6, -1]                   # 1 <- 0, 4 <- 1 (inverted), 7 <- 6
    transforms = [0, 0, 0, 0, 1, 0, 0, 0, 0]
    side = {2: [(5, 200), (P - 1, 7), (P + 40, 99)], 4: [(0, 1), (17, 255)]}
    side_bytes = {i: b"".join(struct.pack(">IB", off, val) for off, val in recs) for i, recs in side.items()}
    # independent reconstruction, in processing order (rti-builder/ptmlib.c:421-440)
    planes = stored.copy()
    for rank in range(n):
        comp = order.index(rank)
        r = ref_planes[comp]
        if r > -1:
            src = planes[r].astype(np.int32)
            if transforms[comp] & 1:
                src = 255 - src
            planes[comp] = (planes[comp].astype(np.int32) + src - 128).astype(np.uint8)   # JSAMPLE wrap-around
        for off, val in side.get(comp, []):
            if off < P:
                planes[comp][off] = val
    coef = np.stack([planes[k] for k in range(6)], -1).reshape(H, W, 6)
    rgb = np.stack([planes[6 + k] for k in range(3)], -1).reshape(H, W, 3)
    scale = np.array([0.9, 1.1, 0.7, 1.3, 0.8, 1.0], np.float32)
    bias = np.array([120, 131, 117, 125, 140, 10], np.int32)
    streams = [b"RAWJ" + struct.pack("<III", W, H, 1) + stored[i].tobytes() for i in range(n)]

    def ints(v):
        return (" ".join(str(int(x)) for x in v) + "\n").encode()

    blob = b"PTM_1.2\nPTM_FORMAT_JPEG_LRGB\n%d\n%d\n" % (W, H)
    blob += (" ".join("%f" % s for s in scale) + "\n").encode() + ints(bias)
    blob += ints([90]) + ints(transforms) + ints([0] * n) + ints([0] * n) + ints(order) + ints(ref_planes)
    blob += ints([len(s) for s in streams]) + ints([len(side_bytes.get(i, b"")) for i in range(n)])
    for i in range(n):
        blob += streams[i] + side_bytes.get(i, b"")
    path = os.path.join(str(tmp_path), "predicted.ptm")
    open(path, "wb").write(blob)
    for uv in (("0.25", "-0.5"), ("0", "0")):
        ra = run([os.path.join(BIN, "ptm-decoder"), path, *uv])
        rb = run([os.path.join(REF, "ptm-decoder-ref"), path, *uv])
        assert ra.returncode == 0 and rb.returncode == 0, ra.stderr.decode() + rb.stderr.decode()
        assert ra.stdout == rb.stdout, "libptm.so and the reference decoder disagree on a predicted / patched PTM"
        want = oracle.relight_lrgb(coef, rgb, oracle.header_roundtrip(scale), bias, float(uv[0]), float(uv[1]))
        np.testing.assert_array_equal(oracle.read_rawj(ra.stdout), want)
