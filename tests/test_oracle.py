"""CPU tests: the oracle against the reference's numeric example, against the
committed outputs of the unmodified reference binaries, and (in the build
container only) against the reference run live."""
import hashlib
import json
import os
import tempfile

import numpy as np
import pytest

from conftest import golden_path

REF_PRESENT = os.path.exists("/root/reference/rti-builder/ptmlib.c")


def test_octave_known_answer(oracle):
    # doc_src/draw_light_illustrations.m:1-59 -- the reference's only pinned numbers
    kat = json.load(open(golden_path("octave_kat.json")))
    A = np.array(kat["A"], np.float32)
    b = np.array(kat["b"], np.float32)
    M = oracle.pinv_from_design(A)
    X = M.astype(np.float64) @ b.astype(np.float64)
    np.testing.assert_allclose(X, kat["X_float64"], rtol=2e-5, atol=2e-4)
    np.testing.assert_allclose(oracle.singular_values(A), kat["singular_values_float64"], rtol=1e-5)
    # SURVEY.md section 4 quotes the same answer
    np.testing.assert_allclose(X, [-50.76406, -61.452795, 19.997354, -50.396717, 20.683573, 120.549492],
                               rtol=2e-5, atol=2e-4)


def test_dome1_pinv_matches_reference_ptm_svd(oracle, dome1):
    g = np.load(golden_path("dome1_pinv.npz"))
    np.testing.assert_array_equal(g["lights"], dome1)
    M = oracle.pinv(dome1)
    # same LAPACKE/BLAS build => identical; other BLAS builds => float32 noise only
    np.testing.assert_allclose(M, g["M"], rtol=0, atol=2e-6)
    np.testing.assert_allclose(M, oracle.pinv_f64(dome1), atol=5e-7)
    np.testing.assert_allclose(np.abs(M).sum(1), [3.163441, 3.163441, 3.08259, 1.503854, 1.503854, 1.830169],
                               rtol=1e-5)


def test_synth_generator_is_pinned(oracle, dome1):
    from rti_b200 import lights
    hashes = json.load(open(golden_path("synth_hashes.json")))
    st = oracle.synth_stack(0x5EED0001, 0, 64, 48, dome1)
    assert hashlib.sha256(st.tobytes()).hexdigest() == hashes["lrgb_64x48"]
    assert st[..., 0].min() >= 1
    s16 = oracle.synth_stack(0x5EED0004, 1, 24, 16, lights.dome100_lights(), dtype=np.uint16)
    assert hashlib.sha256(s16.tobytes()).hexdigest() == hashes["u16_rgb_24x16_dome100"]
    # slabs are independent of the partition
    top = oracle.synth_stack(0x5EED0001, 0, 64, 48, dome1, row0=0, rows=20)
    bot = oracle.synth_stack(0x5EED0001, 0, 64, 48, dome1, row0=20, rows=28)
    np.testing.assert_array_equal(np.concatenate([top, bot], axis=1), st)


@pytest.mark.parametrize("name", ["lrgb_64x48", "lrgb_37x29"])
def test_oracle_vs_reference_encoder_golden_lrgb(oracle, dome1, M60, name):
    g = np.load(golden_path(name + ".npz"))
    st = oracle.synth_stack(int(g["seed"]), int(g["mode"]), int(g["width"]), int(g["height"]), dome1)
    enc = oracle.encode_lrgb(st, M60, blas=False)
    # scale goes through "%f" in the PTM header (rti-builder/ptmlib.c:156-163)
    np.testing.assert_allclose(oracle.header_roundtrip(enc["scale"]), g["scale_text"], atol=2e-6)
    assert np.abs(enc["bias"].astype(int) - g["bias"]).max() <= 1
    np.testing.assert_array_equal(enc["rgb"], g["block1"])          # integer + unfused float path: exact
    d = np.abs(enc["coef"].astype(int) - g["block0"].astype(int))
    assert d.max() <= 1 and (d != 0).mean() < 0.01                   # summation order only
    # relight: restated loop on the reference's own bytes must be exact
    for (u, v), want in zip(g["directions"], g["rasters"]):
        got = oracle.relight_lrgb(g["block0"], g["block1"], g["scale_text"], g["bias"], float(u), float(v))
        np.testing.assert_array_equal(got, want)


def test_oracle_vs_reference_encoder_golden_rgb(oracle, dome1, M60):
    g = np.load(golden_path("rgb_40x24.npz"))
    st = oracle.synth_stack(int(g["seed"]), int(g["mode"]), int(g["width"]), int(g["height"]), dome1)
    enc = oracle.encode_rgb(st, M60, blas=False)
    np.testing.assert_allclose(oracle.header_roundtrip(enc["scale"]), g["scale_text"], atol=2e-6)
    assert np.abs(enc["bias"].astype(int) - g["bias"]).max() <= 1
    ref = np.stack([g["block0"], g["block1"], g["block2"]])
    d = np.abs(enc["coef"].astype(int) - ref.astype(int))
    assert d.max() <= 1 and (d != 0).mean() < 0.01
    for (u, v), want in zip(g["directions"], g["rasters"]):
        got = oracle.relight_rgb(ref, g["scale_text"], g["bias"], float(u), float(v))
        np.testing.assert_array_equal(got, want)


def test_rgb_format_extrema_come_from_first_plane_only(oracle):
    # rti-builder/ptmlib.c:831: the min/max row pointer has no block index
    rng = np.random.default_rng(1)
    c = rng.normal(size=(3, 50, 6)).astype(np.float32)
    c[1] *= 100.0  # must NOT widen the range
    _, sc, _, mm = oracle.scale(c, planes=3)
    np.testing.assert_array_equal(mm[:6], c[0].min(0))
    np.testing.assert_array_equal(mm[6:], c[0].max(0))
    np.testing.assert_allclose(sc, (c[0].max(0) - c[0].min(0)) / 256.0, rtol=1e-6)


def test_clip_truncates(oracle):
    # rti-builder/ptmlib.h:156-158
    c = np.zeros((1, 6), np.float32)
    c[0] = [0, 0, 0, 0, 0, 0]
    coeffs = np.array([[0.0] * 6, [256.0] * 6, [254.9 / 1.0] * 6], np.float32)
    blocks, sc, bi, _ = oracle.scale(coeffs, 1)
    assert sc[0] == 1.0 and bi[0] == 0
    assert blocks[0][:, 0].tolist() == [0, 255, 254]


def test_exact_biquadratic_is_recovered(oracle, dome1, M60):
    # SURVEY.md section 4, property test 3
    u, v = dome1[:, 0].astype(np.float64), dome1[:, 1].astype(np.float64)
    truth = np.array([-40.0, -30.0, 10.0, 25.0, -15.0, 120.0])
    y = truth[0] * u * u + truth[1] * v * v + truth[2] * u * v + truth[3] * u + truth[4] * v + truth[5]
    st = np.zeros((60, 1, 4, 3), np.uint8)
    st[:, 0, :, 0] = np.round(y)[:, None]
    got = oracle.fit_u8(st, M60)[0, 0]
    np.testing.assert_allclose(got, truth, atol=1.5)   # samples were rounded to integers


def test_vertical_flip_commutes(oracle, dome1, M60):
    st = oracle.synth_stack(7, 0, 16, 12, dome1)
    a = oracle.fit_u8(st, M60)
    b = oracle.fit_u8(np.ascontiguousarray(st[:, ::-1]), M60)
    np.testing.assert_array_equal(a[::-1], b)


@pytest.mark.skipif(not REF_PRESENT, reason="needs /root/reference (build container only)")
def test_oracle_vs_reference_live(oracle, dome1, M60):
    """Byte-exact against the reference's own code, run here: unmodified ptmlib.c functions
    through ctypes, and the unmodified encoder / decoder mains end to end."""
    oracle.build()
    assert oracle.ref_available()
    st = oracle.synth_stack(99, 0, 48, 40, dome1)
    np.testing.assert_array_equal(oracle.ref_fit_u8(st, M60), oracle.fit_u8(st, M60, blas=True))
    np.testing.assert_array_equal(oracle.ref_cbcr_avg(st), oracle.chroma_avg(st))
    enc = oracle.encode_lrgb(st, M60, blas=True)
    with tempfile.TemporaryDirectory() as d:
        ref, path = oracle.ref_encode(st, dome1, "PTM_FORMAT_LRGB", d)
        np.testing.assert_array_equal(ref["blocks"][0], enc["coef"])
        np.testing.assert_array_equal(ref["blocks"][1], enc["rgb"])
        np.testing.assert_array_equal(ref["bias"], enc["bias"])
        np.testing.assert_array_equal(ref["scale"], oracle.header_roundtrip(enc["scale"]))
        for u, v in [(0.0, 0.0), (0.5, 0.5), (-0.3, 0.8)]:
            np.testing.assert_array_equal(oracle.ref_decode(path, u, v),
                                          oracle.relight_lrgb(ref["blocks"][0], ref["blocks"][1], ref["scale"],
                                                              ref["bias"], u, v))
    st = oracle.synth_stack(98, 1, 20, 12, dome1)
    enc = oracle.encode_rgb(st, M60, blas=True)
    with tempfile.TemporaryDirectory() as d:
        ref, path = oracle.ref_encode(st, dome1, "PTM_FORMAT_RGB", d)
        np.testing.assert_array_equal(np.stack(ref["blocks"]), enc["coef"])
        np.testing.assert_array_equal(ref["bias"], enc["bias"])
        np.testing.assert_array_equal(oracle.ref_decode(path, 0.2, -0.4),
                                      oracle.relight_rgb(np.stack(ref["blocks"]), ref["scale"], ref["bias"], 0.2, -0.4))


# ------------------------------------------------------------------ round 2: f-4 pieces and the RGB front end

def test_rgb_to_ycbcr_is_the_jfif_conversion(oracle):
    """orc_rgb_to_ycbcr restates libjpeg's integer converter (third party, absent from the reference tree): known
    triplets, and within 1 of the real-valued JFIF matrix for every corner / random colour."""
    known = {(255, 255, 255): (255, 128, 128), (0, 0, 0): (0, 128, 128), (255, 0, 0): (76, 85, 255),
             (0, 255, 0): (150, 44, 21), (0, 0, 255): (29, 255, 107), (128, 128, 128): (128, 128, 128)}
    for rgb, ycc in known.items():
        got = oracle.rgb_to_ycbcr(np.array(rgb, np.uint8).reshape(1, 1, 1, 3)).reshape(3)
        assert tuple(int(x) for x in got) == ycc, (rgb, got)
    rng = np.random.default_rng(1)
    rgb = rng.integers(0, 256, (1, 64, 64, 3), dtype=np.uint8)
    got = oracle.rgb_to_ycbcr(rgb).astype(np.float64)
    r, g, b = (rgb[..., i].astype(np.float64) for i in range(3))
    want = np.stack([0.299 * r + 0.587 * g + 0.114 * b,
                     -0.168735892 * r - 0.331264108 * g + 0.5 * b + 128,
                     0.5 * r - 0.418687589 * g - 0.081312411 * b + 128], -1)
    assert np.abs(got - np.clip(want, 0, 255)).max() <= 1.0


def test_lum_relight_and_normal_map_restatements(oracle, dome1, M60):
    """orc_relight_lum against a plain numpy statement of rti-builder/ptmlib.c:601-609, orc_normal_map against
    the scalar orc_normal (which tests/golden pins)."""
    rng = np.random.default_rng(2)
    W, H = 9, 5
    coef = rng.integers(0, 256, (H, W, 6), dtype=np.uint8)
    crcb = rng.integers(0, 256, (H, W, 2), dtype=np.uint8)
    scale = np.array([0.9, 1.1, 0.7, 1.3, 0.8, 1.0], np.float32)
    bias = np.array([120, 131, 117, 125, 140, 10], np.int32)
    u, v = np.float32(0.3), np.float32(-0.6)
    got = oracle.relight_lum(coef, crcb, scale, bias, float(u), float(v))
    light = [u * u, v * v, u * v, u, v]
    for y in range(H):
        for x in range(W):
            c = coef[H - 1 - y, x].astype(np.int32)
            s = np.float32(0)
            for k in range(5):
                term = np.float32(np.float32(scale[k] * np.float32(c[k] - bias[k])) * light[k])
                s = term if k == 0 else np.float32(s + term)
            s = np.float32(s + np.float32(scale[5] * np.float32(c[5] - bias[5])))
            want_y = 255 if s > 255 else (0 if s < 0 else int(s))
            assert tuple(got[y, x]) == (want_y, crcb[H - 1 - y, x, 1], crcb[H - 1 - y, x, 0])
    coeffs = rng.normal(0, 50, (20, 6)).astype(np.float32)
    nm = oracle.normal_map(coeffs=coeffs)
    for p in range(20):
        want = np.array(oracle.normal(coeffs[p]), np.float32)
        assert np.array_equal(nm[p], want, equal_nan=True)


def test_lsum_branch_restatement(oracle, dome1, M60):
    """orc_fit_lsum: mean chroma = ptm_cbcr_avg's truncated mean scaled by 256 / L_0; the median option is the
    textbook median ((lower + upper middle) / 2)."""
    stack = oracle.synth_stack(9, 1, 7, 5, dome1)
    mean = oracle.fit_lsum(stack, M60, median=False)
    med = oracle.fit_lsum(stack, M60, median=True)
    L0 = stack[0].astype(np.int64).sum(-1)
    avg = stack.astype(np.int64).sum(0) // stack.shape[0]
    np.testing.assert_array_equal(mean["colour"], ((256 * avg) // L0[..., None]).astype(np.uint8))
    m = np.sort(stack.astype(np.int64), 0)
    n = stack.shape[0]
    mid = (m[(n - 1) // 2] + m[n // 2]) // 2
    np.testing.assert_array_equal(med["colour"], ((256 * mid) // L0[..., None]).astype(np.uint8))
    np.testing.assert_array_equal(mean["coeffs"], med["coeffs"])
    # the fit is ptm_fit_poly_uint on L = R + G + B
    L = stack.astype(np.uint32).sum(-1, dtype=np.uint32)[..., None]
    np.testing.assert_array_equal(mean["coeffs"], oracle.fit_u32(np.ascontiguousarray(L), M60, 0))
