#!/usr/bin/env python
"""Relight throughput (BASELINE.json configs[4]): a 24 MP LRGB PTM relit for D directions on a
ring of radius 0.7.  Prints one JSON line per D with pixel-directions/s and the fraction of the
HBM roofline ((9 + 3 D) / D bytes per pixel-direction)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rti_b200 import lights  # noqa: E402
from rti_b200.device import U8, light_q14  # noqa: E402
from rti_b200.encoder import PtmEncoder, PtmRelighter  # noqa: E402
from rti_b200.ptmlib import ptm_svd  # noqa: E402


def main():
    W, H = 6000, 4000
    P = W * H
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    L = lights.dome1_lights()
    M = ptm_svd(L)
    stack = torch.empty((60, H, W, 3), dtype=torch.uint8, device="cuda")
    enc = PtmEncoder(W, H, M)
    enc.ctx.synth(0x5EED0001, 0, U8, 0, P, light_q14(L), stack, P * 3)
    enc.encode(stack)
    torch.cuda.synchronize()
    scale, bias = enc.scale_bias()
    del stack
    torch.cuda.empty_cache()
    rel = PtmRelighter(W, H)
    for D in (1, 8, 60, 360):
        th = np.deg2rad(np.arange(D) * 360.0 / D)
        uv = np.stack([0.7 * np.cos(th), 0.7 * np.sin(th)], 1).astype(np.float32)
        out = torch.empty((D, H, W, 3), dtype=torch.uint8, device="cuda")
        for _ in range(2):
            rel.relight_lrgb(enc.coef_bytes[0], enc.rgb, scale, bias, uv, out)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5 if D < 100 else 2
        a.record()
        for _ in range(reps):
            rel.relight_lrgb(enc.coef_bytes[0], enc.rgb, scale, bias, uv, out)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        pd = P * D / (ms * 1e-3)
        bytes_pd = (9 + 3 * D) / D
        print(json.dumps({"directions": D, "ms": ms, "pixel_directions_per_s": pd,
                          "achieved_GBs": pd * bytes_pd / 1e9, "frac_of_hbm_peak": pd * bytes_pd / 1e9 / peak}))
        del out
    # RGB-format PTM (three coefficient planes, CLIP(POLY) per channel): (18 + 3 D) / D bytes per pixel-direction
    del enc
    torch.cuda.empty_cache()
    stack = torch.empty((60, H, W, 3), dtype=torch.uint8, device="cuda")
    enc = PtmEncoder(W, H, M, planes=3)
    enc.ctx.synth(0x5EED0001, 1, U8, 0, P, light_q14(L), stack, P * 3)
    enc.encode(stack)
    torch.cuda.synchronize()
    scale, bias = enc.scale_bias()
    del stack
    torch.cuda.empty_cache()
    for D in (1, 60, 360):
        th = np.deg2rad(np.arange(D) * 360.0 / D)
        uv = np.stack([0.7 * np.cos(th), 0.7 * np.sin(th)], 1).astype(np.float32)
        out = torch.empty((D, H, W, 3), dtype=torch.uint8, device="cuda")
        cb = enc.coef_bytes
        for _ in range(2):
            rel.relight_rgb(cb[0], cb[1], cb[2], scale, bias, uv, out)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5 if D < 100 else 2
        a.record()
        for _ in range(reps):
            rel.relight_rgb(cb[0], cb[1], cb[2], scale, bias, uv, out)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / reps
        pd = P * D / (ms * 1e-3)
        bytes_pd = (18 + 3 * D) / D
        print(json.dumps({"format": "RGB", "directions": D, "ms": ms, "pixel_directions_per_s": pd,
                          "achieved_GBs": pd * bytes_pd / 1e9, "frac_of_hbm_peak": pd * bytes_pd / 1e9 / peak}))
        del out


if __name__ == "__main__":
    main()
