#!/usr/bin/env python
"""One 24 MP x D-direction relight call (for ncu): python tools/relight_once.py [D] [lrgb|rgb] [auto|exact|fast]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rti_b200.encoder import PtmRelighter  # noqa: E402

D = int(sys.argv[1]) if len(sys.argv) > 1 else 60
fmt = sys.argv[2] if len(sys.argv) > 2 else "lrgb"
mode = sys.argv[3] if len(sys.argv) > 3 else "auto"
W, H = 6000, 4000
g = torch.Generator(device="cuda").manual_seed(1)
planes = [torch.randint(0, 256, (H * W, 6), dtype=torch.uint8, device="cuda", generator=g) for _ in range(3 if fmt == "rgb" else 1)]
rgb = torch.randint(0, 256, (H * W, 3), dtype=torch.uint8, device="cuda", generator=g)
scale = np.array([0.58, 0.60, 0.69, 1.42, 1.42, 0.38], np.float32)
bias = np.array([130, 128, 127, 120, 131, -20], np.int32)
th = np.deg2rad(np.arange(D) * 360.0 / D)
uv = np.stack([0.7 * np.cos(th), 0.7 * np.sin(th)], 1).astype(np.float32)
rel = PtmRelighter(W, H)
rel.ctx.set_relight_mode(mode)
out = torch.empty((D, H, W, 3), dtype=torch.uint8, device="cuda")


def once():
    if fmt == "rgb":
        rel.relight_rgb(planes[0], planes[1], planes[2], scale, bias, uv, out)
    else:
        rel.relight_lrgb(planes[0], rgb, scale, bias, uv, out)


for _ in range(3):
    once()
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
once()
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b)
bpd = ((18 if fmt == "rgb" else 9) + 3 * D) * W * H
print("%s D=%d mode=%s: %.3f ms, %.0f GB/s algorithmic" % (fmt, D, mode, ms, bpd / ms / 1e6))
