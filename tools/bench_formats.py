#!/usr/bin/env python
"""Throughput of the other encode configurations (not the headline bench line):
RGB-format uint8 at 24 MP x 60 lights, and the 16-bit / 100-light configuration at a size given
on the command line (default 12.5 MP so that it runs in seconds).  One JSON line each."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rti_b200 import lights  # noqa: E402
from rti_b200.device import U8, U16, light_q14  # noqa: E402
from rti_b200.encoder import PtmEncoder  # noqa: E402
from rti_b200.ptmlib import ptm_svd  # noqa: E402


def timed(fn, reps):
    for _ in range(2):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def run(name, W, H, L, dtype, planes, mode, reps, peak):
    st, td = (U8, torch.uint8) if dtype == "u8" else (U16, torch.uint16)
    n, P = len(L), W * H
    M = ptm_svd(L)
    stack = torch.empty((n, H, W, 3), dtype=td, device="cuda")
    enc = PtmEncoder(W, H, M, planes=planes)
    enc.ctx.synth(0x5EED0001, mode, st, 0, P, light_q14(L), stack, P * 3 * stack.element_size())
    ms = timed(lambda: enc.encode(stack), reps)
    bpp = n * 3 * stack.element_size() + (18 if planes == 3 else 9)
    print(json.dumps({"config": name, "pixels": P, "lights": n, "ms": ms, "Mpixel_per_s": P / ms / 1e3,
                      "algorithmic_B_per_px": bpp, "frac_of_hbm_peak": P * bpp / (ms * 1e-3) / 1e9 / peak}))
    del stack, enc
    torch.cuda.empty_cache()


def main():
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
    mp16 = float(sys.argv[1]) if len(sys.argv) > 1 else 12.5
    run("RGB-format u8, 24 MP x 60", 6000, 4000, lights.dome1_lights(), "u8", 3, 1, 5, peak)
    H16 = int(mp16 * 1e6 / 8688)
    run("RGB-format u16, %.1f MP x 100" % (8688 * H16 / 1e6), 8688, H16, lights.dome100_lights(), "u16", 3, 1, 2, peak)
    run("LRGB u16 (extension), %.1f MP x 100" % (8688 * H16 / 1e6), 8688, H16, lights.dome100_lights(), "u16", 1, 0, 2, peak)


if __name__ == "__main__":
    main()
