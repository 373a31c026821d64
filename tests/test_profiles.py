"""The committed bench lines under profiles/ must be internally consistent (no GPU needed): throughput follows from
the step time, the roofline fraction from SURVEY.md 8(d)'s 189 bytes per pixel and the recorded peak, every parity
gate of the run is green, and the sharded runs hash to the committed golden blocks."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINES = sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_bench_n[0-9].json")) +
               glob.glob(os.path.join(ROOT, "profiles", "r02_bench_n[0-9]_driver_c*.json")))
PX = 6000 * 4000


@pytest.mark.parametrize("path", LINES, ids=[os.path.basename(p) for p in LINES])
def test_bench_line_is_consistent(path):
    d = json.load(open(path))
    n = d["n_gpus"]
    assert d["unit"] == "Mpixel/s" and d["higher_is_better"] is True and d["data"] == "synthetic"
    assert d["value"] == pytest.approx(PX / (d["ms_per_step"] * 1e-3) / 1e6, rel=1e-6)
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["algorithmic_bytes_per_px"] == 189
    assert r["achieved"] == pytest.approx(189 * PX / (d["ms_per_step"] * 1e-3) / 1e9 / n, rel=1e-6)
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9)
    assert 0.5 < r["frac"] < 1.0
    assert d["gpu_launches"] >= 2 * d["steps"]
    par = d["parity"]
    if n == 1:
        assert par["ok"] and par["rgb_identical"] and par["max_abs"] <= 1 and par["bias_delta"] == 0
        assert par["coeff_rel_err_16_rows"] <= 1e-5 and par["scale_rel"] <= 1e-5
        assert par.get("matches_golden_sha256", True) is True
        assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
        assert d["e2e_dropin"]["identical_to_device_path"] and d["e2e_dropin"]["registered"]["identical_to_device_path"]
    else:
        assert par["sharded_identical"] is True
        gold = json.load(open(os.path.join(ROOT, "tests", "golden", "blocks_24mp_lrgb_sha256.json")))["sha256"]
        if "single_gpu_blocks_sha256" in par:
            assert par["single_gpu_blocks_sha256"] == gold
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == PX * 180 and e["d2h_bytes_per_step"] == PX * 9
    assert e["value"] == pytest.approx(PX / (e["ms_per_step"] * 1e-3) / 1e6, rel=1e-6)
    assert e.get("identical_to_device_path", e.get("identical_to_single_gpu")) is True
    assert e["value"] < d["value"]          # the host-buffer leg is PCIe bound and can never beat the resident one


def test_traffic_file_matches_the_kernel_sources():
    """profiles/traffic.json is quoted by bench.py only while the hash of K1's / K2's sources matches."""
    import sys
    sys.path.insert(0, ROOT)
    import bench
    per, src = bench.read_traffic(["fit_lrgb_u8_tma_kernel", "quantise_kernel"])
    if per is None:     # the kernels were edited after the last capture: bench.py reports the capture as stale, too
        pytest.skip(src)
    whole = per.pop("_whole_encode")
    alg = 189 * PX
    assert alg < whole["float_plane_scratch_tail_96MB"] < whole["float_plane_kept"] < 1.3 * alg
    assert 1.2 * alg < sum(per.values()) < 1.3 * alg     # cold-cache per-kernel capture: the all-HBM two-pass traffic
