"""CPU tests of the host-side logic around the kernels."""
import numpy as np

from rti_b200 import lights
from rti_b200.device import key_to_float, light_q14
from rti_b200.encoder import slab_rows


def float_to_key(f):
    b = np.asarray(f, np.float32).view(np.int32)
    return b ^ ((b >> 31) & 0x7FFFFFFF)


def test_keys_preserve_float_order():
    rng = np.random.default_rng(0)
    f = np.concatenate([rng.normal(scale=1e3, size=1000), [0.0, -0.0, 1e-38, -1e-38, 3.4e38, -3.4e38,
                                                              np.inf, -np.inf]]).astype(np.float32)
    k = float_to_key(f)
    # total order of the keys = float order, with -0.0 sorted just below +0.0
    order_f = np.lexsort((~np.signbit(f), f))
    assert np.all(np.diff(k[order_f].astype(np.int64)) >= 0)
    np.testing.assert_array_equal(key_to_float(k).view(np.int32), f.view(np.int32))
    # max over keys == key of max, so an integer MAX reduction is an exact float max
    assert key_to_float(np.array([k.max()]))[0] == f.max()
    assert -key_to_float(np.array([float_to_key(-f).max()]))[0] == f.min()


def test_magic_division_is_exact():
    # K1 divides byte sums by the light count with one multiply-high (ptm_device.cuh div_magic)
    for d in list(range(2, 258)) + [300, 1000, 4096]:
        magic = (1 << 32) // d + 1
        x = np.arange(0, 255 * d + 1, dtype=np.uint64)
        np.testing.assert_array_equal((x * magic) >> 32, x // d)


def test_slab_partition_covers_rows():
    for H in (1, 7, 4000, 4001, 5792):
        for G in (1, 2, 3, 4, 8):
            parts = [slab_rows(H, G, r) for r in range(G)]
            assert parts[0][0] == 0 and parts[-1][1] == H
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))


def test_light_sets():
    d = lights.dome1_lights()
    assert d.shape == (60, 3) and d.dtype == np.float32
    file_rows = lights.read_skeleton_lp(lights.dome1_lp_path())
    np.testing.assert_array_equal(d, file_rows)
    np.testing.assert_allclose((d.astype(np.float64) ** 2).sum(1), 1.0, atol=2e-6)
    assert lights.dome100_lights().shape == (100, 3)


def test_lp_grammar(tmp_path):
    # rti-builder/ptm-encoder.c:152-157: count line skipped, w optional
    p = tmp_path / "s.lp"
    p.write_text("3\na.jpg 0.1 0.2 0.97\nb.jpg -0.5 0.25\n\njunk\nc.jpg 0 0 1\n")
    names, L = lights.read_lp(str(p))
    assert names == ["a.jpg", "b.jpg", "c.jpg"]
    np.testing.assert_allclose(L, [[0.1, 0.2, 0.97], [-0.5, 0.25, 0.0], [0, 0, 1]], rtol=1e-6)


def test_light_q14_matches_oracle():
    from oracle import oracle as O
    d = lights.dome1_lights()
    np.testing.assert_array_equal(light_q14(d), O.light_q14(d))


def test_mma_digit_packing_reproduces_the_dot_product_exactly(oracle, dome1, M60):
    """Host half of the tensor-core engine (rti_b200/csrc/ptm_fit_mma.cu:pack_mma_digits): the int8 digit
    image, pushed through an integer GEMM in numpy exactly as tcgen05.mma kind::i8 would, gives the
    dot product with M to within 2^-31 of each row's largest entry."""
    import ctypes as C
    from rti_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(5)
    for M in (M60, oracle.pinv(np.asarray(dome1)[:13]), (rng.standard_normal((6, 100)) * 1e-3).astype(np.float32)):
        M = np.ascontiguousarray(M, dtype=np.float32)
        n = M.shape[1]
        kpad = C.c_uint()
        image = np.zeros(32 * ((n + 31) // 32 * 32), dtype=np.int8)
        scale = np.zeros(6, dtype=np.float32)
        _lib.check(lib.ptmb_mma_pack_digits(M.ctypes.data, n, image.ctypes.data, image.nbytes, scale.ctypes.data,
                                            C.byref(kpad)))
        assert kpad.value == (n + 31) // 32 * 32
        B = image.reshape(kpad.value // 16, 4, 8, 16).transpose(0, 3, 1, 2).reshape(kpad.value, 32).astype(np.int64)
        assert (B[n:] == 0).all() and (B[:n, 24] == 1).all() and (B[:, 25:] == 0).all()
        y = rng.integers(0, 256, size=(50, kpad.value)).astype(np.int64)      # padding lights hold garbage
        D = y @ B
        np.testing.assert_array_equal(D[:, 24], y[:, :n].sum(1))
        for k in range(6):
            t = (D[:, 4 * k] << 24) + (D[:, 4 * k + 1] << 16) + (D[:, 4 * k + 2] << 8) + D[:, 4 * k + 3]
            got = t.astype(np.float64) * float(scale[k])
            want = y[:, :n].astype(np.float64) @ M[k].astype(np.float64)
            tol = 2.0 ** -31 * np.abs(M[k]).max() * y[:, :n].sum(1)
            assert (np.abs(got - want) <= tol).all()
    bad = M60.copy()
    bad[2, 5] = np.inf
    assert lib.ptmb_mma_pack_digits(bad.ctypes.data, 60, image.ctypes.data, image.nbytes, scale.ctypes.data,
                                    C.byref(kpad)) != 0


def test_weighted_row_cuts_cover_the_image():
    """Bandwidth-proportional row slabs of the host-buffer legs (bench.weighted_cuts, same formula as
    ptm_multi.cu:cut): boundaries are monotone, cover every row once, follow the weights, and equal weights
    reproduce the floor(i*H/n) slabs up to rounding."""
    import bench
    rng = np.random.default_rng(5)
    for H in (1, 2, 7, 500, 4000, 5792):
        for n in (1, 2, 3, 8):
            for w in ([1.0] * n, list(rng.uniform(0.2, 5.0, n)), [23.3] * (n // 2) + [35.3] * (n - n // 2)):
                cuts = bench.weighted_cuts(H, w)
                assert len(cuts) == n + 1 and cuts[0] == 0 and cuts[-1] == H
                assert all(a <= b for a, b in zip(cuts, cuts[1:]))
                tot = sum(w)
                for i in range(n):
                    assert abs((cuts[i + 1] - cuts[i]) - H * w[i] / tot) <= 1.0
    assert bench.weighted_cuts(4000, [23.3] * 4 + [35.3] * 4) == [0, 398, 795, 1193, 1590, 2193, 2795, 3398, 4000]


def test_golden_block_hash_is_wired_into_the_bench():
    """bench.py compares the single-GPU blocks of its default workload with a committed SHA-256 (N = 1 and N > 1 lines)."""
    import bench
    h = bench.golden_blocks_sha256(bench.WIDTH, bench.HEIGHT)
    assert isinstance(h, str) and len(h) == 64 and int(h, 16) >= 0
    assert bench.golden_blocks_sha256(64, 48) is None
