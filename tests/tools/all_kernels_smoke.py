#!/usr/bin/env python
"""Small end-to-end pass over every kernel family, odd and even sizes, so that out-of-bounds
accesses in tile / tail handling show up under compute-sanitizer memcheck where that tool is
available (it is closed on this pool; the pass is also a quick all-kernel smoke on its own):
    python tests/tools/all_kernels_smoke.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402  (tools/ is test tooling)
from rti_b200 import lights  # noqa: E402
from rti_b200.device import U8, U16, Context, light_q14  # noqa: E402
from rti_b200.encoder import PtmEncoder  # noqa: E402
from rti_b200.ptmlib import ptm_svd  # noqa: E402

torch.cuda.set_device(0)
L = lights.dome1_lights()
M = ptm_svd(L)
ctx = Context(0)
for (W, H) in [(64, 48), (37, 29), (1040, 3), (16, 1), (3, 1)]:
    P = W * H
    stack = torch.empty((60, H, W, 3), dtype=torch.uint8, device="cuda")
    ctx.synth(1, 0, U8, 0, P, light_q14(L), stack, P * 3)
    enc = PtmEncoder(W, H, M)
    enc.encode(stack)            # fused
    enc.fit(stack)
    enc.exchange()
    enc.quantise()               # stepwise
    torch.cuda.synchronize()
    sc, bi = enc.scale_bias()
    out = torch.empty((5, H, W, 3), dtype=torch.uint8, device="cuda")
    uv = np.array([[0, 0], [0.5, 0.5], [-0.7, 0.1], [0.2, -0.9], [1, 0]], np.float32)
    ctx.relight_lrgb(enc.coef_bytes[0], enc.rgb, W, H, sc, bi, uv, out)
    e3 = PtmEncoder(W, H, M, planes=3)
    e3.encode(stack)
    ctx.relight_rgb(e3.coef_bytes[0], e3.coef_bytes[1], e3.coef_bytes[2], W, H, sc, bi, uv, out)
    host = ctx.encode_host(stack.cpu().numpy(), M, planes=1, want_coeffs=True)
    torch.cuda.synchronize()
L100 = lights.dome100_lights()
M100 = ptm_svd(L100)
s16 = torch.from_numpy(O.synth_stack(2, 1, 24, 10, L100, dtype=np.uint16).view(np.int16)).cuda().view(torch.uint16)
e16 = PtmEncoder(24, 10, M100, planes=3)
e16.encode(s16)
e16l = PtmEncoder(24, 10, M100, planes=1)
e16l.encode(s16)
torch.cuda.synchronize()
print("sanitize smoke ok")
