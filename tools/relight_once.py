#!/usr/bin/env python
"""One 24 MP x D-direction LRGB relight call (for ncu): python tools/relight_once.py [D]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rti_b200.encoder import PtmRelighter  # noqa: E402

D = int(sys.argv[1]) if len(sys.argv) > 1 else 60
W, H = 6000, 4000
g = torch.Generator(device="cuda").manual_seed(1)
coef = torch.randint(0, 256, (H * W, 6), dtype=torch.uint8, device="cuda", generator=g)
rgb = torch.randint(0, 256, (H * W, 3), dtype=torch.uint8, device="cuda", generator=g)
scale = np.array([0.58, 0.60, 0.69, 1.42, 1.42, 0.38], np.float32)
bias = np.array([130, 128, 127, 120, 131, -20], np.int32)
th = np.deg2rad(np.arange(D) * 360.0 / D)
uv = np.stack([0.7 * np.cos(th), 0.7 * np.sin(th)], 1).astype(np.float32)
rel = PtmRelighter(W, H)
out = torch.empty((D, H, W, 3), dtype=torch.uint8, device="cuda")
for _ in range(3):
    rel.relight_lrgb(coef, rgb, scale, bias, uv, out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
rel.relight_lrgb(coef, rgb, scale, bias, uv, out)
b.record()
torch.cuda.synchronize()
print("D=%d: %.3f ms" % (D, a.elapsed_time(b)))
