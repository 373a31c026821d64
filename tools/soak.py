#!/usr/bin/env python
"""Soak test of the fused encode loop: many back-to-back encodes (K1 -> hand-off -> K2, each launched
programmatically dependent on its predecessor, tiles drawn dynamically) on both engines and two slab sizes;
every result must equal the first one bit for bit."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rti_b200 import lights  # noqa: E402
from rti_b200.device import U8, light_q14  # noqa: E402
from rti_b200.encoder import PtmEncoder  # noqa: E402
from rti_b200.ptmlib import ptm_svd  # noqa: E402

L = lights.dome1_lights()
M = ptm_svd(L)
for (W, H, reps) in ((6000, 500, 3000), (6000, 4000, 300), (2000, 37, 3000)):
    P = W * H
    stack = torch.empty((len(L), H, W, 3), dtype=torch.uint8, device="cuda")
    for engine in ("ffma", "mma"):
        enc = PtmEncoder(W, H, M, engine=engine)
        enc.ctx.synth(0x5EED0001, 0, U8, 0, P, light_q14(L), stack, P * 3)
        enc.encode(stack)
        torch.cuda.synchronize()
        ref = (enc.coef_bytes.clone(), enc.rgb.clone(), enc.coeffs.clone(), enc.sb.clone())
        bad = 0
        for i in range(reps):
            enc.encode(stack)
            if i % 97 == 0 or i == reps - 1:
                torch.cuda.synchronize()
                ok = (torch.equal(enc.coef_bytes, ref[0]) and torch.equal(enc.rgb, ref[1]) and
                      torch.equal(enc.coeffs, ref[2]) and torch.equal(enc.sb, ref[3]))
                bad += 0 if ok else 1
        print("%dx%d %s: %d encodes, %d mismatching checks" % (W, H, engine, reps, bad))
        assert bad == 0
        del enc
print("soak ok")
