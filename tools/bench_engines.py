#!/usr/bin/env python
"""FFMA engine against the tensor-core engine on the same resident stack: fit alone (K1) and the
whole encode, LRGB and RGB formats, uint8, 24 MP x dome1.lp by default.  One JSON line per case.

Tuning: PTM_MMA_VARIANT=1 runs four epilogue groups; variants 2-4 (ceilings: load + MMA only, TMA
delivery only, MMA only) exist only in a library built with `make EXTRA_NVFLAGS=-DPTM_TUNING` and do
not produce results.  PTM_MMA_SCHED=0 strides units over the grid, PTM_MMA_LOAD=0 uses plain 2-D boxes."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rti_b200 import lights  # noqa: E402
from rti_b200.device import U8, light_q14  # noqa: E402
from rti_b200.encoder import PtmEncoder  # noqa: E402
from rti_b200.ptmlib import ptm_svd  # noqa: E402


def timed(fn, reps):
    for _ in range(3):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (6000, 4000)
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = json.load(open(pk))["hbm_gbs"] if os.path.exists(pk) else 6650.0
    L = lights.dome1_lights()
    if os.environ.get("N_LIGHTS"):          # any other light count: a spiral of that many directions
        import numpy as np
        nl = int(os.environ["N_LIGHTS"])
        th = np.arange(nl) * 2.399963
        r = 0.15 + 0.8 * (np.arange(nl) + 0.5) / nl
        L = np.stack([r * np.cos(th), r * np.sin(th), np.sqrt(1 - r * r)], 1).astype(np.float32)
    n, P = len(L), W * H
    M = ptm_svd(L)
    stack = torch.empty((n, H, W, 3), dtype=torch.uint8, device="cuda")
    for planes, mode, name in ((1, 0, "LRGB"), (3, 1, "RGB")):
        for engine in os.environ.get("ENGINES", "ffma,mma").split(","):
            enc = PtmEncoder(W, H, M, planes=planes, engine=engine)
            enc.ctx.synth(0x5EED0001, mode, U8, 0, P, light_q14(L), stack, P * 3)
            before = enc.ctx.mma_launches()
            fit = timed(lambda: enc.fit(stack), reps)
            whole = timed(lambda: enc.encode(stack), reps)
            ran_mma = enc.ctx.mma_launches() > before
            bpp = n * 3 + (18 if planes == 3 else 9)
            k1_bpp = n * 3 + (72 if planes == 3 else 27)
            print(json.dumps({"format": name, "engine": engine, "tensor_core_kernel_ran": ran_mma, "pixels": P,
                              "fit_ms": fit, "encode_ms": whole, "encode_Mpixel_per_s": P / whole / 1e3,
                              "fit_GBs": P * k1_bpp / fit / 1e6, "fit_frac_of_hbm_peak": P * k1_bpp / fit / 1e6 / peak,
                              "encode_frac_of_hbm_roofline": P * bpp / whole / 1e6 / peak}))
            del enc
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
