#!/usr/bin/env python
"""Benchmark of the PTM fit hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # B200 arm (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path, same workload

A step is one LRGB PTM fit of a 24 MP (6000 x 4000) x 60-light (dome1.lp) uint8 stack:
K1 (fit + chroma + colour/normalise + extrema) -> hand-off of the 12 extrema keys (to every
GPU's mailbox over NVLink when N > 1) -> K2 (scale/bias + quantise).  For N > 1 the SAME 24 MP
image is row-sharded over the ranks (BASELINE.json configs[2]) -- strong scaling.

value        pixels / device time, stack already resident in HBM (CUDA events, max over ranks)
e2e          the same through the host-buffer C-ABI call (ptmb_encode_host_*): pinned host stack
             -> H2D -> K1 -> exchange -> K2 -> D2H of the blocks, all inside the timed region;
             next to it the measured pinned H2D rate of the same buffers (all ranks at once)
e2e_dropin   (N = 1) the reference's OWN call sequence through libptm.so (ptm_fit_poly_jsample,
             ptm_cbcr_avg, colour loop, ptm_scale_coefficients) on pageable host memory
roofline     the whole fit path on SURVEY.md 8(d)'s 189 algorithmic bytes per pixel against the
             measured HBM peak; K1 (dominant kernel) timed live with CUDA events beside it
parity       (N = 1) every byte of the GPU blocks against the reference's CPU result for the same
             24 MP stack; (N > 1) every rank's sharded blocks against a single-GPU encode of the
             whole image made on that rank
extra        (N = 1) BASELINE.json configs[3] / configs[4]: 50 MP x 100 x 16-bit LRGB and RGB
             encodes, the 24 MP RGB-format encode, 360-direction relight sweeps
cpu_baseline the reference's own CPU functions on this box's host cores, whole 24 MP image

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT, N_LIGHTS = 6000, 4000, 60
SEED = 0x5EED0001
METRIC = "PTM fit Mpixel/s at 24MP x dome1.lp lights (LRGB, uint8)"
WORKLOAD = "%.0fMP (%dx%d) x %d lights dome1.lp, LRGB uint8"
ALG_BYTES_PATH = N_LIGHTS * 3 + 6 + 3          # 189 B/px: SURVEY.md 8(d), whole fit path
ALG_BYTES_K1 = N_LIGHTS * 3 + 24 + 3           # 207 B/px: K1's own compulsory input + output
FALLBACK_HBM_GBS = 6650.0                      # /opt/skills/guides/B200_PROFILING.md


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--driver", default="torchrun", choices=["torchrun", "c"],
                    help="N > 1: one process per GPU (torchrun), or ONE process driving all GPUs through the "
                         "C ABI's multi-GPU entry (ptmb_multi_*; launch with plain `python bench.py --gpus N --driver c`)")
    ap.add_argument("--width", type=int, default=WIDTH)
    ap.add_argument("--height", type=int, default=HEIGHT)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-dropin", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline passes")
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def kernel_sources_sha256():
    """Hash of the sources of the two kernels profiles/traffic.json speaks about (K1 = ptm_fit_tma.cu, K2 =
    ptm_quantise.cu, and the headers they include); traffic.json records the hash its ncu capture was taken with,
    so a capture of other kernels is reported as stale instead of being quoted."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "rti_b200", "csrc")
    for name in ("ptm_fit_tma.cu", "ptm_quantise.cu", "ptm_device.cuh", "ptm_tma.cuh", "ptm_kernels.h", "ptm_internal.h"):
        h.update(name.encode())
        h.update(open(os.path.join(d, name), "rb").read())
    return h.hexdigest()


def read_traffic(kernels):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum from an `ncu --set full` capture of
    this bench command) for the named kernels, or (None, why)."""
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(tpath))
    except Exception:
        return None, "profiles/traffic.json missing"
    if t.get("kernel_sources_sha256") != kernel_sources_sha256():
        return None, "profiles/traffic.json was captured with other kernel sources (stale)"
    try:
        per = {k: float(t[k]["dram_bytes_read"]) + float(t[k]["dram_bytes_write"]) for k in kernels}
    except Exception:
        return None, "profiles/traffic.json lacks " + ", ".join(kernels)
    whole = t.get("whole_encode_range_replay")
    if whole:   # HBM bytes of a WHOLE encode with L2 left alone between K1 and K2 (ncu range replay), both plane modes
        per["_whole_encode"] = {k: float(v["dram_bytes_read"]) + float(v["dram_bytes_write"])
                                for k, v in whole.items() if isinstance(v, dict)}
    return per, t.get("source", "profiles/traffic.json")


# --------------------------------------------------------------------- clocks

class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        self.error = None

    def run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {
                getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8): "hw_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
                getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4): "sw_power_cap",
                getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
            }
            while not self._stop_evt.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                    for bit, name in names.items():
                        if bit and (r & bit):
                            self.reasons.add(name)
                except Exception:
                    pass
                self._stop_evt.wait(0.001)
        except Exception as e:  # NVML missing: report it, do not fake numbers
            self.error = repr(e)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": self.error or "no samples"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def gpu_cpu_affinity(index):
    """CPU cores NVML reports as local to the GPU (its NUMA node), or an empty set."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        mask = nv.nvmlDeviceGetCpuAffinity(h, n_words)
        return {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
    except Exception:
        return set()


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPU cores local to the GPU BEFORE any pinned host buffer is allocated, so that the
    buffers of the e2e leg are first-touched on the GPU's NUMA node (matters with one rank per GPU)."""
    try:
        cpus = gpu_cpu_affinity(index) & set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return 0


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


# ------------------------------------------------------------------- CPU legs

def reference_pass(O, stack, M, threads):
    dt, kind, res = O.cpu_encode_lrgb_timed(stack, M, threads=threads)
    return dt, kind, res


def cpu_baseline(stack_np, lights, M, target_seconds):
    """The reference's CPU sequence over the WHOLE image of this workload on all host cores, repeated until about
    `target_seconds` have been spent.  Returns (cpu_baseline object, the reference result of the last pass)."""
    from oracle import oracle as O
    cores = O.set_threads(O.host_cores())
    n, H, W, _ = stack_np.shape
    reference_pass(O, stack_np[:, :16], M, cores)           # warm-up (thread pools, page faults)
    total, passes, kind, res = 0.0, 0, "port", None
    while passes < 1 or (total < target_seconds and passes < 200):
        dt, kind, res = reference_pass(O, stack_np, M, cores)
        total += dt
        passes += 1
    return {"value": W * H * passes / total / 1e6, "unit": "Mpixel/s", "cores": cores, "kind": kind,
            "sample": "%d pass(es) over the whole %d x %d x %d-light stack (same bytes as the GPU run), %.1f s in "
                      "total; OpenBLAS %s, OMP threads = %d" % (passes, W, H, n, total,
                                                                O.lib().orc_blas_corename().decode(), cores)}, res


def run_reference(args):
    """--impl reference: the reference's own CPU implementation on ALL host cores, one whole 24 MP stack per step
    (same workload as the B200 arm).  Under torchrun rank 0 alone runs; the launcher's OMP_NUM_THREADS=1 is
    overridden through the OpenMP runtime itself."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    from rti_b200 import lights as LS
    L = LS.dome1_lights()
    M = O.pinv(L)
    try:
        os.sched_setaffinity(0, range(os.cpu_count() or 1))   # undo any launcher pinning
    except Exception:
        pass
    cores = O.set_threads(O.host_cores())
    W, H = args.width, args.height
    st = O.synth_stack(SEED, 0, W, H, L)
    kind = "port"
    for _ in range(max(1, min(args.warmup, 2))):
        _, kind, _ = reference_pass(O, st, M, cores)
    t = 0.0
    for _ in range(args.steps):
        dt, kind, _ = reference_pass(O, st, M, cores)
        t += dt
    value = W * H * args.steps / t / 1e6
    sample = "the whole %d x %d x %d-light stack per step (same workload as the B200 arm), %d OMP threads, OpenBLAS %s" % (
        W, H, len(L), cores, O.lib().orc_blas_corename().decode())
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": max(1, min(args.warmup, 2)), "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD % (W * H / 1e6, W, H, N_LIGHTS), "arm": "reference CPU path, whole image per step"},
        "cpu_baseline": {"value": value, "unit": "Mpixel/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------- parity

def parity_vs_reference(enc, ref, M, stack_np):
    """Every byte of the device-resident result against the reference's CPU result for the same stack
    (SURVEY.md 8(d) "parity gates run with every measurement")."""
    import torch
    H, W = ref["coef"].shape[:2]
    got_coef = enc.coef_bytes[0].cpu().numpy().reshape(H, W, 6)
    got_rgb = enc.rgb.cpu().numpy().reshape(H, W, 3)
    scale, bias = enc.scale_bias()
    d = np.abs(got_coef.astype(np.int16) - ref["coef"].astype(np.int16))
    # float coefficients: |delta| relative to the coefficient's scale sum_n |M_kn| y_n * 256 / mean(y)
    rows = np.linspace(0, H - 1, 16).astype(int)
    got_f = enc.coeffs[0].reshape(H, W, 6)[torch.from_numpy(rows).cuda()].cpu().numpy()
    y = stack_np[:, rows][..., 0].astype(np.float64)          # [n][16][W]
    bound = np.einsum("kn,nrw->rwk", np.abs(M).astype(np.float64), y) * (256.0 / np.maximum(np.floor(y.sum(0) / y.shape[0]), 1.0))[..., None]
    rel = float((np.abs(got_f.astype(np.float64) - ref["coeffs"][rows].astype(np.float64)) / np.maximum(bound, 1e-30)).max())
    out = {
        "against": "reference CPU result, whole image (%d x %d)" % (W, H),
        "coef_bytes_differing": int((d != 0).sum()), "coef_bytes_total": int(d.size), "max_abs": int(d.max()),
        "rgb_identical": bool(np.array_equal(got_rgb, ref["rgb"])),
        "scale_rel": float(np.max(np.abs(scale - ref["scale"]) / np.abs(ref["scale"]))),
        "bias_delta": int(np.max(np.abs(bias.astype(np.int64) - ref["bias"].astype(np.int64)))),
        "coeff_rel_err_16_rows": rel,
    }
    out["ok"] = bool(out["max_abs"] <= (2 if out["bias_delta"] else 1) and out["rgb_identical"] and
                     out["bias_delta"] <= 1 and out["scale_rel"] <= 1e-5 and rel <= 1e-5)
    return out


def golden_blocks_sha256(W, H):
    """The committed hash of the single-GPU blocks of the default workload (tests/golden), or None."""
    if (W, H) != (WIDTH, HEIGHT):
        return None
    try:
        return json.load(open(os.path.join(ROOT, "tests", "golden", "blocks_24mp_lrgb_sha256.json")))["sha256"]
    except Exception:
        return None


def sharded_identity(enc, M, L, W, H, r0, r1, world):
    """This rank's sharded blocks against ITS OWN single-GPU encode of the whole image: byte-identical or not."""
    import torch
    import torch.distributed as dist
    from rti_b200.device import U8, light_q14
    from rti_b200.encoder import PtmEncoder
    P = W * H
    full = torch.empty((N_LIGHTS, H, W, 3), dtype=torch.uint8, device="cuda")
    single = PtmEncoder(W, H, M, planes=1, exchange="none", group=None, single=True)
    single.ctx.synth(SEED, 0, U8, 0, P, light_q14(L), full, P * 3)
    single.encode(full)
    torch.cuda.synchronize()
    same = (torch.equal(single.coef_bytes[0][r0 * W:r1 * W], enc.coef_bytes[0]) and
            torch.equal(single.rgb[r0 * W:r1 * W], enc.rgb) and torch.equal(single.sb, enc.sb))
    digest = hashlib.sha256(single.coef_bytes[0].cpu().numpy().tobytes() + single.rgb.cpu().numpy().tobytes()).hexdigest()
    flag = torch.tensor([1 if same else 0], dtype=torch.int32, device="cuda")
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    blocks = (single.coef_bytes[0].clone(), single.rgb.clone())
    del full, single
    torch.cuda.empty_cache()
    return bool(flag.item()), digest, blocks


def h2d_shares(world, seconds=0.012, chunk=8 << 20):
    """GB/s every rank gets when ALL ranks copy pinned host memory to their GPU for the same span of time (the GPUs
    of one box do not share the host's PCIe bandwidth equally: on this pool's 8-GPU boxes four get 23 GB/s and four
    35 GB/s).  The e2e leg splits the image's rows in these proportions."""
    import torch
    import torch.distributed as dist
    host = torch.empty(2 * chunk, dtype=torch.uint8).pin_memory()
    dev = torch.empty(2 * chunk, dtype=torch.uint8, device="cuda")
    dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(), torch.cuda.Event()]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    moved, k = 0, 0
    dev[:chunk].copy_(host[:chunk], non_blocking=True)
    ev[0].record()
    while True:
        k ^= 1
        dev[k * chunk:(k + 1) * chunk].copy_(host[k * chunk:(k + 1) * chunk], non_blocking=True)
        ev[k].record()
        ev[k ^ 1].synchronize()
        moved += chunk
        dt = time.perf_counter() - t0
        if dt >= seconds:
            break
    torch.cuda.synchronize()
    rate = torch.tensor([moved / dt / 1e9], dtype=torch.float64, device="cuda")
    rates = [torch.zeros_like(rate) for _ in range(world)]
    if world > 1:
        dist.all_gather(rates, rate)
    else:
        rates = [rate]
    return [float(r.item()) for r in rates]


def weighted_cuts(height, weights):
    """Row boundaries [c_0 = 0, ..., c_n = height] with slab i ~ weights[i] (formula of ptm_multi.cu: cut)."""
    total = float(sum(weights))
    cuts, upto = [0], 0.0
    for w in weights[:-1]:
        upto += w
        cuts.append(min(height, int(height * (upto / total) + 0.5)))
    cuts.append(height)
    return cuts


# ------------------------------------------------------------------- extras (configs[3], configs[4])

def timed_ms(fn, reps, warm=2):
    import torch
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def run_extras(M60, L60, peak, gpu_index):
    """BASELINE.json configs[3] (50 MP x 100 lights x 16 bit, LRGB + RGB format) and configs[4] (360-direction
    relight sweep over a 24 MP PTM), plus the 24 MP RGB-format encode; a few steps each, clocks sampled."""
    import torch
    from rti_b200 import lights as LS
    from rti_b200.device import U8, U16, light_q14
    from rti_b200.encoder import PtmEncoder, PtmRelighter
    from rti_b200.ptmlib import ptm_svd
    out = []
    keep = os.environ.get("PTM_KEEP_COEFFS", "0") == "1"

    def record(name, fn, reps, units, alg_bytes, unit_name, **more):
        s = ClockSampler(gpu_index)
        s.start()
        ms = timed_ms(fn, reps)
        clocks = s.stop()
        gbs = alg_bytes / (ms * 1e-3) / 1e9
        out.append(dict({"config": name, "ms": ms, "steps": reps, "value": units / (ms * 1e-3) / 1e6,
                         "unit": unit_name, "algorithmic_bytes": alg_bytes, "achieved_gbs": gbs,
                         "frac_of_hbm_peak": gbs / peak, "clocks": clocks}, **more))

    def encode_case(name, W, H, L, dtype, planes, mode, reps):
        st, td = (U8, torch.uint8) if dtype == "u8" else (U16, torch.uint16)
        n, P = len(L), W * H
        M = ptm_svd(L)
        stack = torch.empty((n, H, W, 3), dtype=td, device="cuda")
        enc = PtmEncoder(W, H, M, planes=planes, keep_coeffs=keep)     # float plane as in the headline run
        enc.ctx.synth(SEED, mode, st, 0, P, light_q14(L), stack, P * 3 * stack.element_size())
        bpp = n * 3 * stack.element_size() + (18 if planes == 3 else 9)
        record(name, lambda: enc.encode(stack), reps, P, P * bpp, "Mpixel/s", algorithmic_bytes_per_px=bpp,
               l2="stack larger than L2, streamed once per step")
        return enc, stack

    # configs[3]: 8688 x 5792 = 50.3 MP, 100 lights, 16-bit linear samples
    L100 = LS.dome100_lights()
    for name, planes, mode in (("C4 LRGB u16, 50.3MP x 100 lights", 1, 0), ("C4 RGB-format u16, 50.3MP x 100 lights", 3, 1)):
        enc, stack = encode_case(name, 8688, 5792, L100, "u16", planes, mode, 3)
        del enc, stack
        torch.cuda.empty_cache()
    enc, stack = encode_case("RGB-format u8, 24MP x 60 lights", WIDTH, HEIGHT, L60, "u8", 3, 1, 5)
    del stack
    # configs[4]: 360 directions on a ring of radius 0.7; RGB-format PTM first (its blocks are at hand)
    W, H = WIDTH, HEIGHT
    P = W * H
    D = 360
    th = np.deg2rad(np.arange(D) * 360.0 / D)
    uv = np.stack([0.7 * np.cos(th), 0.7 * np.sin(th)], 1).astype(np.float32)
    rel = PtmRelighter(W, H)
    rasters = torch.empty((D, H, W, 3), dtype=torch.uint8, device="cuda")
    scale, bias = enc.scale_bias()
    cb = enc.coef_bytes
    for mode in ("fast", "exact"):
        rel.ctx.set_relight_mode(mode)
        record("C5 relight sweep, RGB-format 24MP PTM x 360 directions (%s)" % mode,
               lambda: rel.relight_rgb(cb[0], cb[1], cb[2], scale, bias, uv, rasters), 2, P * D, P * (18 + 3 * D),
               "Mpixel-directions/s", tolerance="+-1 per byte" if mode == "fast" else "bit-exact",
               l2="26 GB of rasters written per step")
    del enc, cb
    torch.cuda.empty_cache()
    stack = torch.empty((N_LIGHTS, H, W, 3), dtype=torch.uint8, device="cuda")
    enc = PtmEncoder(W, H, M60, planes=1)
    enc.ctx.synth(SEED, 0, U8, 0, P, light_q14(L60), stack, P * 3)
    enc.encode(stack)
    torch.cuda.synchronize()
    scale, bias = enc.scale_bias()
    for mode in ("fast", "exact"):
        rel.ctx.set_relight_mode(mode)
        record("C5 relight sweep, LRGB 24MP PTM x 360 directions (%s)" % mode,
               lambda: rel.relight_lrgb(enc.coef_bytes[0], enc.rgb, scale, bias, uv, rasters), 2, P * D,
               P * (9 + 3 * D), "Mpixel-directions/s", tolerance="+-1 per byte" if mode == "fast" else "bit-exact",
               l2="26 GB of rasters written per step")
    del rasters, enc
    torch.cuda.empty_cache()
    # LAST (it heats the board): the headline workload held for 500 fits back to back -- the 1 kW power cap settles
    # the SM clock below the burst clock of the 20-step headline run; both fit engines (the FFMA ring is the default
    # for LRGB), each after a pause so that neither starts on the other's heat
    bpp = N_LIGHTS * 3 + 9
    for engine in ("ffma", "mma"):
        time.sleep(3.0)
        e2 = PtmEncoder(W, H, M60, planes=1, engine=engine, keep_coeffs=keep)
        record("headline 24MP x 60 LRGB u8, 500 fits back to back (sustained clocks), %s engine" % engine,
               lambda: e2.encode(stack), 500, P, P * bpp, "Mpixel/s", algorithmic_bytes_per_px=bpp,
               l2="stack larger than L2, streamed once per step")
        del e2
    del stack
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------- drop-in path

def run_dropin(stack_np, L, want_blocks):
    """The reference encoder's own call sequence (rti-builder/ptm-encoder.c:316-353,406) through the ptmlib.h
    drop-in (libptm.so), host memory in and out.  Two legs: plain pageable buffers (what the unmodified main
    allocates), and the same buffers after the two-line change INTEGRATION.md shows (ptmb_host_register on the
    stack and the coefficient buffer right after they are allocated: the DMA engines then read / write the
    caller's memory directly instead of going through the library's pinned staging ring)."""
    from rti_b200 import _lib, ptmlib
    n, H, W, _ = stack_np.shape
    M = ptmlib.ptm_svd(L)
    stack_np = np.array(stack_np, copy=True)               # a plain malloc'd (pageable) copy, like the reference's buffer

    buffers = ptmlib.stepwise_buffers(W, H)                # caller-owned outputs, allocated and touched up front

    def once():
        t0 = time.perf_counter()
        res = ptmlib.encode_lrgb_stepwise(stack_np, M, buffers)
        return time.perf_counter() - t0, res

    def leg():
        once()
        times = []
        res = None
        for _ in range(2):
            dt, res = once()
            times.append((dt, res["ms"]))
        s, per_call = min(times, key=lambda x: x[0])
        same = bool(np.array_equal(res["coef"].reshape(-1, 6), want_blocks[0]) and
                    np.array_equal(res["rgb"].reshape(-1, 3), want_blocks[1]))
        return {"value": W * H / s / 1e6, "unit": "Mpixel/s", "ms_per_step": s * 1e3, "ms_per_call": per_call,
                "identical_to_device_path": same}

    out = leg()
    out.update({"h2d_bytes_per_step": int(stack_np.nbytes), "d2h_bytes_per_step": int(W * H * 9),
                "calls": "ptm_fit_poly_jsample -> ptm_cbcr_avg -> ptm_lrgb_finish -> ptm_scale_coefficients (libptm.so)",
                "host_memory": "pageable (plain numpy arrays for the stack, the float coefficients and the blocks)"})
    # the same caller buffers, registered once by the caller (cost reported, not part of the step)
    lib = _lib.load()
    owned = [stack_np, buffers["coeffs"]] + list(buffers["blocks"])
    t0 = time.perf_counter()
    done = []
    try:
        for a in owned:
            _lib.check(lib.ptmb_host_register(a.ctypes.data, a.nbytes))
            done.append(a)
        reg_s = time.perf_counter() - t0
        reg = leg()
        reg["host_memory"] = ("the same buffers after ptmb_host_register (cudaHostRegister) by the caller: %.0f ms once "
                              "for %.2f GB, not part of the step" % (reg_s * 1e3, sum(a.nbytes for a in owned) / 1e9))
        out["registered"] = reg
    except Exception as e:      # registration refused (locked-memory limit): the pageable leg stands on its own
        out["registered"] = {"unavailable": repr(e)[:200]}
    finally:
        for a in done:
            lib.ptmb_host_unregister(a.ctypes.data)
    return out


# ------------------------------------------------------------------- B200 arm

def run_b200(args):
    import torch
    import torch.distributed as dist

    from rti_b200 import lights as LS
    from rti_b200.device import U8, Context, light_q14
    from rti_b200.encoder import PtmEncoder, allreduce_keys, slab_rows
    from rti_b200.ptmlib import ptm_svd

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    numa_cpus = bind_to_gpu_numa_node(physical_gpu_index(local)) if world > 1 else 0
    if world > 1:
        # NCCL announces its version on stdout at communicator creation when NCCL_DEBUG=VERSION is set in the
        # environment; stdout must carry the one JSON line only, so it points at stderr until NCCL is up
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)
    assert world == args.gpus, "launch with torchrun --nproc-per-node %d (WORLD_SIZE=%d)" % (args.gpus, world)

    W, H = args.width, args.height
    L = LS.dome1_lights()
    M = ptm_svd(L)                                            # host LAPACKE pseudo-inverse
    r0, r1 = slab_rows(H, world, rank)
    rows = r1 - r0
    P = rows * W

    stack = torch.empty((N_LIGHTS, rows, W, 3), dtype=torch.uint8, device="cuda")
    # the float coefficients are an intermediate of the encode (the reference's encoder frees them after
    # ptm_scale_coefficients): declared scratch, K2 may drop their L2 lines instead of sending them to HBM and back
    # (ptmb_set_coeffs_scratch; PTM_KEEP_COEFFS=1 measures with the float plane kept as an output)
    keep_coeffs = os.environ.get("PTM_KEEP_COEFFS", "0") == "1"
    enc = PtmEncoder(W, rows, M, planes=1, keep_coeffs=keep_coeffs, exchange=os.environ.get("PTM_EXCHANGE", "p2p"))
    enc.ctx.synth(SEED, 0, U8, r0 * W, P, light_q14(L), stack, P * 3)
    torch.cuda.synchronize()
    # the stack is resident and complete before the first encode is enqueued: K1's copy warp may start while the
    # previous K2 drains (ptmb_set_stack_stable; PTM_STACK_STABLE=0 measures without it)
    stack_stable = os.environ.get("PTM_STACK_STABLE", "1") == "1"
    enc.ctx.set_stack_stable(stack_stable)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(values):
        t = torch.tensor(values, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.tolist()

    def step():
        enc.encode(stack)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- device-resident timing.  Nothing but the encodes is enqueued between the two events: an event record
    # between K1 and K2 would break their programmatic (PDL) chaining, so K1 is timed in a SECOND pass of the same
    # loop, where the library brackets its fit stage with CUDA events on the launching stream (ptmb_probe_*).
    # N > 1: one untimed encode right after the barrier lines the ranks up on the device (every K2 waits for all
    # peers' keys), so the max over ranks does not include the host barrier's release jitter.
    fused = enc.launches_per_encode == 2
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(physical_gpu_index(local)) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    if world > 1:
        step()
    t_start.record()
    for i in range(args.steps):
        step()
    t_end.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_total = t_start.elapsed_time(t_end)

    p_start, p_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if world > 1:
        step()
    if fused:
        enc.ctx.probe_arm(args.steps)
    else:
        k1_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                     for _ in range(args.steps)]
    p_start.record()
    for i in range(args.steps):
        if fused:
            step()
        else:
            k1_events[i][0].record()
            enc.fit(stack)
            k1_events[i][1].record()
            enc.exchange()
            enc.quantise()
    p_end.record()
    barrier()
    probed_ms = p_start.elapsed_time(p_end) / args.steps
    if fused:
        k1_ms = float(np.mean(enc.ctx.probe_read(args.steps)))
    else:
        k1_ms = float(np.mean([a.elapsed_time(b) for a, b in k1_events]))
    (probed_ms,) = allmax([probed_ms])
    ms_total, k1_ms = allmax([ms_total, k1_ms])
    ms_step = ms_total / args.steps
    value = W * H / (ms_step * 1e-3) / 1e6

    # ---- N > 1: the sharded blocks must equal a single-GPU encode of the whole image
    sharded, single_blocks = None, None
    if world > 1:
        same, digest, single_blocks = sharded_identity(enc, M, L, W, H, r0, r1, world)
        gold = golden_blocks_sha256(W, H)
        sharded = {"sharded_identical": same, "single_gpu_blocks_sha256": digest,
                   "matches_golden_sha256": None if gold is None else bool(gold == digest),
                   "how": "every rank encodes the whole image on its own GPU (no exchange) and compares its slab"}

    # ---- end to end through the host-buffer C-ABI call.  The rows are split over the ranks in proportion to the
    # host bandwidth each GPU gets when all of them copy at once (measured here, PTM_E2E_BALANCE=0: equal slabs):
    # this leg is bound by PCIe, not by the kernels, so its slabs follow the links.
    e2e = None
    host_stack = None
    if not args.no_e2e:
        ctx = Context(local, stream=torch.cuda.current_stream().cuda_stream)
        shares = h2d_shares(world) if world > 1 else [1.0]
        balanced = world > 1 and os.environ.get("PTM_E2E_BALANCE", "1") == "1"
        cuts = weighted_cuts(H, shares) if balanced else [slab_rows(H, world, i)[0] for i in range(world)] + [H]
        refined = 0
        while True:
            e0, e1 = cuts[rank], cuts[rank + 1]
            erows = e1 - e0
            if (e0, e1) == (r0, r1):
                dev_slab = stack
            else:
                dev_slab = torch.empty((N_LIGHTS, erows, W, 3), dtype=torch.uint8, device="cuda")
                enc.ctx.synth(SEED, 0, U8, e0 * W, erows * W, light_q14(L), dev_slab, erows * W * 3)
                torch.cuda.synchronize()
            host_stack = torch.empty((N_LIGHTS, erows, W, 3), dtype=torch.uint8).pin_memory()
            host_stack.copy_(dev_slab)                        # same synthetic bytes, now in pinned host memory

            # the ceiling: the same pinned buffers copied to the device with nothing else going on, all ranks at once
            scratch = dev_slab if dev_slab is not stack else torch.empty_like(stack)
            scratch.copy_(host_stack, non_blocking=True)
            barrier()
            ca, cb_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ca.record()
            for _ in range(2):
                scratch.copy_(host_stack, non_blocking=True)
            cb_.record()
            barrier()
            my_ms = ca.elapsed_time(cb_) / 2
            del scratch, dev_slab
            per_rank = torch.zeros(world, dtype=torch.float64, device="cuda")
            per_rank[rank] = my_ms
            if world > 1:
                dist.all_reduce(per_rank, op=dist.ReduceOp.SUM)
            per_rank = per_rank.tolist()
            h2d_ms = max(per_rank)
            # one refinement with the real buffers: the probe's small copies do not load the links exactly like
            # slab-sized ones; rows follow rows_i / t_i when the ranks finished more than 4 % apart
            if balanced and refined < 1 and min(per_rank) > 0 and h2d_ms / min(per_rank) > 1.04:
                cuts = weighted_cuts(H, [max(1e-9, (cuts[i + 1] - cuts[i]) / per_rank[i]) for i in range(world)])
                refined += 1
                del host_stack
                continue
            break
        out_coef = torch.empty((erows, W, 6), dtype=torch.uint8).pin_memory()
        out_rgb = torch.empty((erows, W, 3), dtype=torch.uint8).pin_memory()
        keys = torch.empty(12, dtype=torch.int32, device="cuda")
        h2d_peak_gbs = N_LIGHTS * W * H * 3 / (h2d_ms * 1e-3) / 1e9     # aggregate over all ranks

        def e2e_step():
            ctx.encode_host_fit(host_stack.data_ptr(), (N_LIGHTS, erows, W), U8, M, planes=1, keys_out=keys,
                                early_rgb=out_rgb)
            allreduce_keys(keys)
            return ctx.encode_host_finish(keys, [out_coef, out_rgb])

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()                                        # synchronous: returns after the D2H
        barrier()
        (e2e_s,) = allmax([time.perf_counter() - t0])
        e2e_s /= args.e2e_steps
        # the host path must give the same bytes as the device-resident one (N > 1: as a single-GPU encode of the
        # whole image, whatever the split)
        if single_blocks is not None:
            want_coef, want_rgb = single_blocks[0][e0 * W:e1 * W].cpu(), single_blocks[1][e0 * W:e1 * W].cpu()
        else:
            want_coef, want_rgb = enc.coef_bytes[0].cpu(), enc.rgb.cpu()
        assert torch.equal(out_coef.reshape(-1, 6), want_coef), "e2e bytes differ from device path"
        assert torch.equal(out_rgb.reshape(-1, 3), want_rgb), "e2e colour block differs from device path"
        h2d_gbs = N_LIGHTS * W * H * 3 / e2e_s / 1e9
        e2e = {"value": W * H / e2e_s / 1e6, "unit": "Mpixel/s",
               "h2d_bytes_per_step": int(N_LIGHTS * W * H * 3), "d2h_bytes_per_step": int(W * H * 9),
               "ms_per_step": e2e_s * 1e3, "steps": args.e2e_steps,
               "h2d_gbs": h2d_gbs, "h2d_peak_gbs": h2d_peak_gbs, "frac_of_h2d_peak": h2d_gbs / h2d_peak_gbs,
               "h2d_peak_note": "aggregate rate of plain pinned cudaMemcpyAsync of the same buffers, all ranks at once",
               "h2d_share_gbs_per_rank": shares if world > 1 else None,
               "rows_per_rank": [cuts[i + 1] - cuts[i] for i in range(world)], "h2d_ms_per_rank": per_rank,
               "rebalanced_with_real_buffers": refined,
               "numa_bound_cpus": numa_cpus,
               "identical_to_device_path": True,
               "note": "whole-job bytes; every rank: pinned host slab -> ptmb_encode_host_fit/finish -> pinned host "
                       "blocks; slabs proportional to each GPU's measured share of the host's PCIe bandwidth"}
        ctx.close()
    del single_blocks

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    k1_gbs = ALG_BYTES_K1 * P / (k1_ms * 1e-3) / 1e9
    path_gbs = ALG_BYTES_PATH * W * H / (ms_step * 1e-3) / 1e9 / world
    traffic, traffic_src, whole_encode = None, "single-GPU captures only", None
    if world == 1 and (W, H) == (WIDTH, HEIGHT):
        per, traffic_src = read_traffic(["fit_lrgb_u8_tma_kernel", "quantise_kernel"])
        whole_encode = per.pop("_whole_encode", None) if per else None
        traffic = None if per is None else sum(per.values())
    out = {
        "metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD % (W * H / 1e6, W, H, N_LIGHTS), "sharding": "row slabs over %d GPU(s)" % world,
                   "rows_per_gpu": rows, "l2": "inputs larger than L2 (%.0f MB stack per GPU, streamed once per step)"
                                               % (stack.numel() / 1e6),
                   "driver": "one process per GPU (torchrun)" if world > 1 else "one process",
                   "stack_stable": stack_stable,
                   "float_plane": "kept as an output" if keep_coeffs else
                                  "scratch (ptmb_set_coeffs_scratch): K2 quantises its L2-resident end first and "
                                  "discards those lines; parity of the floats is checked on a separate encode that keeps them",
                   "timed_region": "K back-to-back encodes between two CUDA events, nothing else enqueued" +
                                   ("; one untimed encode after the barrier lines the ranks up on the device first"
                                    if world > 1 else ""),
                   "exchange": {"none": "none", "nccl": "int32 MAX all-reduce of 12 keys (NCCL)",
                                "p2p": "12 int32 keys written into every peer's mailbox over NVLink by K1's last block, picked up in K2's prologue"}
                               [enc.exchange_kind]},
        "roofline": {"bound": "hbm", "kernel": "fit path K1 + K2 (fit_lrgb_u8_tma_kernel dominant)",
                     "achieved": path_gbs, "peak": peak, "unit": "GB/s", "frac": path_gbs / peak,
                     "algorithmic_bytes_per_px": ALG_BYTES_PATH, "peak_source": peak_src,
                     "traffic": traffic, "traffic_ratio": None if traffic is None else traffic / (ALG_BYTES_PATH * W * H),
                     "traffic_source": traffic_src,
                     "traffic_whole_encode_range_replay": whole_encode,
                     "k1_ms": k1_ms, "k1_timed_in": "a second pass of the same loop with CUDA events around K1 "
                                                     "(%.4f ms per step there)" % probed_ms,
                     "k1_achieved": k1_gbs, "k1_frac": k1_gbs / peak,
                     "k1_algorithmic_bytes_per_px": ALG_BYTES_K1,
                     "k1_share_of_step": k1_ms / ms_step},
        "clocks": clocks,
        "gpu_launches": args.steps * enc.launches_per_encode,
    }
    if e2e:
        out["e2e"] = e2e
    if sharded:
        out["parity"] = sharded
    if world == 1:
        if not args.no_extra and (W, H) == (WIDTH, HEIGHT):
            # release the headline stack first: configs[3] alone is 30 GB + 26 GB of rasters for configs[4]
            out["extra"] = run_extras(M, L, peak, physical_gpu_index(local))
        if not args.no_cpu_baseline:
            from oracle import oracle as O  # noqa: F401  (cpu_baseline leg: the one place bench may run the oracle)
            stack_np = host_stack.numpy() if host_stack is not None else stack.cpu().numpy()
            out["cpu_baseline"], ref = cpu_baseline(stack_np, L, M, args.cpu_seconds)
            if not keep_coeffs:
                # the timed encodes left the float plane undefined: their bytes must equal those of an encode that
                # keeps it, and that encode's floats are the ones compared with the reference
                timed_bytes, timed_rgb, timed_sb = enc.coef_bytes.clone(), enc.rgb.clone(), enc.sb.clone()
                enc.ctx.set_coeffs_scratch(False)
                enc.encode(stack)
                torch.cuda.synchronize()
                assert torch.equal(timed_bytes, enc.coef_bytes) and torch.equal(timed_rgb, enc.rgb) and \
                    torch.equal(timed_sb, enc.sb), "scratch-plane encode differs from the encode that keeps its floats"
                del timed_bytes, timed_rgb, timed_sb
            out["parity"] = parity_vs_reference(enc, ref, M, stack_np)
            digest = hashlib.sha256(enc.coef_bytes[0].cpu().numpy().tobytes() + enc.rgb.cpu().numpy().tobytes()).hexdigest()
            gold = golden_blocks_sha256(W, H)
            out["parity"]["blocks_sha256"] = digest
            out["parity"]["matches_golden_sha256"] = None if gold is None else bool(gold == digest)
            if not args.no_dropin:
                out["e2e_dropin"] = run_dropin(stack_np, L, (enc.coef_bytes[0].cpu().numpy(), enc.rgb.cpu().numpy()))
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.driver == "c" and args.gpus > 1:
        from rti_b200.multi import bench_multi          # single process, all GPUs through the C ABI
        bench_multi(args, sys.modules[__name__])
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
