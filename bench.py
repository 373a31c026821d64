#!/usr/bin/env python
"""Benchmark of the PTM fit hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # B200 arm (torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path

A step is one LRGB PTM fit of a 24 MP (6000 x 4000) x 60-light (dome1.lp) uint8 stack:
keys-init -> K1 (fit + chroma + colour/normalise + extrema) -> [int32 MAX all-reduce of the
12 extrema keys when N > 1] -> K2 (scale/bias + quantise).  For N > 1 the SAME 24 MP image is
row-sharded over the ranks (BASELINE.json configs[2]) -- strong scaling.

value   = pixels / device time, stack already resident in HBM (CUDA events, max over ranks)
e2e     = the same through the host-buffer C-ABI call (ptmb_encode_host_*): pinned host
          stack -> H2D -> K1 -> exchange -> K2 -> D2H of the blocks, all inside the timed region
roofline= K1 (dominant kernel) against the measured HBM peak, timed live with CUDA events
cpu_baseline = the reference's own CPU functions on this box's host cores (bounded sample)

Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
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
ALG_BYTES_PATH = N_LIGHTS * 3 + 6 + 3          # 189 B/px: SURVEY.md 8(d), whole fit path
ALG_BYTES_K1 = N_LIGHTS * 3 + 24 + 3           # 207 B/px: K1's own compulsory input + output
FALLBACK_HBM_GBS = 6650.0                      # /opt/skills/guides/B200_PROFILING.md


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--width", type=int, default=WIDTH)
    ap.add_argument("--height", type=int, default=HEIGHT)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    return ap.parse_args()


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


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


def bind_to_gpu_numa_node(index):
    """Pin this process to the CPU cores NVML reports as local to the GPU, so that the pinned host
    buffers of the e2e leg are first-touched on the GPU's NUMA node (matters with one rank per GPU)."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        n_words = (os.cpu_count() + 63) // 64
        mask = nv.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
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

def host_sample_stack(width, rows, lights):
    from oracle import oracle as O
    return O.synth_stack(SEED, 0, width, rows, lights, row0=0, rows=rows)


def cpu_baseline(width, lights, M, target_seconds):
    """Reference CPU sequence on a bounded row slab of the same workload, repeated until about
    `target_seconds` of CPU wall time have been spent."""
    from oracle import oracle as O
    cores = os.cpu_count() or 1
    rows = 1200                                              # 7.2 MP slab: 1.3 GB of host memory
    st = host_sample_stack(width, rows, lights)
    O.cpu_encode_lrgb_timed(st[:, :16], M, threads=cores)    # warm-up (thread pools, page faults)
    total, passes, kind = 0.0, 0, "port"
    while total < target_seconds and passes < 200:
        dt, kind, _ = O.cpu_encode_lrgb_timed(st, M, threads=cores)
        total += dt
        passes += 1
    return {"value": width * rows * passes / total / 1e6, "unit": "Mpixel/s", "cores": cores, "kind": kind,
            "sample": "%d passes over %d rows x %d px x %d lights of the same synthetic stack, %.1f s in total; "
                      "OpenBLAS %s, OMP threads = cores" % (passes, rows, width, len(lights), total,
                                                            O.lib().orc_blas_corename().decode())}


def run_reference(args):
    """--impl reference: the reference's own CPU implementation, all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle as O
    from rti_b200 import lights as LS
    L = LS.dome1_lights()
    M = O.pinv(L)
    cores = os.cpu_count() or 1
    rows = 200                                                # 1.2 MP per step: seconds, not minutes
    st = host_sample_stack(args.width, rows, L)
    kind = "port"
    for _ in range(max(1, args.warmup)):
        _, kind, _ = O.cpu_encode_lrgb_timed(st, M, threads=cores)
    t = 0.0
    for _ in range(args.steps):
        dt, kind, _ = O.cpu_encode_lrgb_timed(st, M, threads=cores)
        t += dt
    value = args.width * rows * args.steps / t / 1e6
    sample = "%d rows x %d px x %d lights per step (bounded sample of the 24 MP workload)" % (rows, args.width, len(L))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "24MP (6000x4000) x 60 lights dome1.lp, LRGB uint8 -- CPU on a bounded sample",
                   "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Mpixel/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


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
    if world > 1:
        bind_to_gpu_numa_node(physical_gpu_index(local))
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
    enc = PtmEncoder(W, rows, M, planes=1, keep_coeffs=os.environ.get("PTM_KEEP_COEFFS", "1") == "1",
                     exchange=os.environ.get("PTM_EXCHANGE", "p2p"))
    enc.ctx.synth(SEED, 0, U8, r0 * W, P, light_q14(L), stack, P * 3)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        enc.encode(stack)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- device-resident timing; the library brackets its fit stage (K1) with CUDA events on the
    # same stream inside the timed region (ptmb_probe_*), which is what the roofline object uses
    fused = enc.launches_per_encode == 2
    if fused:
        enc.ctx.probe_arm(args.steps)
    else:
        k1_events = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                     for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(physical_gpu_index(local)) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    t_start.record()
    for i in range(args.steps):
        if fused:
            step()
        else:
            k1_events[i][0].record()
            enc.fit(stack)
            k1_events[i][1].record()
            enc.exchange()
            enc.quantise()
    t_end.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_total = t_start.elapsed_time(t_end)
    if fused:
        k1_ms = float(np.mean(enc.ctx.probe_read(args.steps)))
    else:
        k1_ms = float(np.mean([a.elapsed_time(b) for a, b in k1_events]))
    tt = torch.tensor([ms_total, k1_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total, k1_ms = tt.tolist()
    ms_step = ms_total / args.steps
    value = W * H / (ms_step * 1e-3) / 1e6

    # ---- end to end through the host-buffer C-ABI call
    e2e = None
    if not args.no_e2e:
        ctx = Context(local, stream=torch.cuda.current_stream().cuda_stream)
        host_stack = torch.empty((N_LIGHTS, rows, W, 3), dtype=torch.uint8).pin_memory()
        host_stack.copy_(stack)                               # same synthetic bytes, now in pinned host memory
        out_coef = torch.empty((rows, W, 6), dtype=torch.uint8).pin_memory()
        out_rgb = torch.empty((rows, W, 3), dtype=torch.uint8).pin_memory()
        keys = torch.empty(12, dtype=torch.int32, device="cuda")

        def e2e_step():
            ctx.encode_host_fit(host_stack.data_ptr(), (N_LIGHTS, rows, W), U8, M, planes=1, keys_out=keys,
                                early_rgb=out_rgb)
            allreduce_keys(keys)
            return ctx.encode_host_finish(keys, [out_coef, out_rgb])

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            res = e2e_step()                                  # synchronous: returns after the D2H
        barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e_s = dt.item() / args.e2e_steps
        # the host path must give the same bytes as the device-resident one
        assert torch.equal(out_coef.reshape(-1, 6), enc.coef_bytes[0].cpu()), "e2e bytes differ from device path"
        assert torch.equal(out_rgb.reshape(-1, 3), enc.rgb.cpu()), "e2e colour block differs from device path"
        assert torch.equal(out_rgb.reshape(-1, 3), enc.rgb.cpu())
        e2e = {"value": W * H / e2e_s / 1e6, "unit": "Mpixel/s",
               "h2d_bytes_per_step": int(host_stack.numel()), "d2h_bytes_per_step": int(out_coef.numel() + out_rgb.numel()),
               "ms_per_step": e2e_s * 1e3, "steps": args.e2e_steps,
               "note": "per-rank bytes; pinned host stack -> ptmb_encode_host_fit/finish -> pinned host blocks"}
        del host_stack
        ctx.close()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if world == 1 and os.path.exists(tpath) and (W, H) == (WIDTH, HEIGHT):
        try:
            traffic = float(json.load(open(tpath))["fit_lrgb_u8_tma_kernel"]["dram_bytes_per_launch"])
        except Exception:
            traffic = None
    k1_gbs = ALG_BYTES_K1 * P / (k1_ms * 1e-3) / 1e9
    path_gbs = ALG_BYTES_PATH * W * H / (ms_step * 1e-3) / 1e9 / world
    out = {
        "metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "24MP (%dx%d) x %d lights dome1.lp, LRGB uint8, row-sharded over %d GPU(s)"
                               % (W, H, N_LIGHTS, world),
                   "rows_per_gpu": rows, "l2": "inputs larger than L2 (%.0f MB stack per GPU, streamed once per step)"
                                               % (stack.numel() / 1e6),
                   "exchange": {"none": "none", "nccl": "int32 MAX all-reduce of 12 keys (NCCL)",
                                "p2p": "12 int32 keys written into every peer's mailbox over NVLink by K1's last block, picked up in K2's prologue"}
                               [enc.exchange_kind]},
        "roofline": {"bound": "hbm", "kernel": "fit_lrgb_u8 (K1)", "achieved": k1_gbs, "peak": peak, "unit": "GB/s",
                     "frac": k1_gbs / peak, "traffic": traffic, "peak_source": peak_src,
                     "algorithmic_bytes_per_px": ALG_BYTES_K1, "k1_ms": k1_ms,
                     "path_achieved": path_gbs, "path_frac": path_gbs / peak,
                     "path_algorithmic_bytes_per_px": ALG_BYTES_PATH},
        "clocks": clocks,
        "gpu_launches": args.steps * enc.launches_per_encode,
    }
    if e2e:
        out["e2e"] = e2e
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O  # noqa: F401  (cpu_baseline leg: the one place bench may run the oracle)
        out["cpu_baseline"] = cpu_baseline(W, L, M, args.cpu_seconds)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
