#!/usr/bin/env python
"""Per-CTA timeline of K1 on a small slab (needs a library built with EXTRA_NVFLAGS=-DPTM_TUNING):
when each CTA entered, got its first stage, stored its last tile and left, relative to the earliest entry."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rti_b200 import _lib, lights  # noqa: E402
from rti_b200.device import U8, light_q14  # noqa: E402
from rti_b200.encoder import PtmEncoder  # noqa: E402
from rti_b200.ptmlib import ptm_svd  # noqa: E402

W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (6000, 500)
L = lights.dome1_lights()
M = ptm_svd(L)
P = W * H
stack = torch.empty((len(L), H, W, 3), dtype=torch.uint8, device="cuda")
enc = PtmEncoder(W, H, M, engine="ffma")
enc.ctx.synth(0x5EED0001, 0, U8, 0, P, light_q14(L), stack, P * 3)
for _ in range(5):
    enc.encode(stack)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20):
    enc.encode(stack)
b.record()
torch.cuda.synchronize()
print("encode %.4f ms" % (a.elapsed_time(b) / 20))
n = 296
t = np.zeros(4 * n, dtype=np.uint64)
lib = _lib.load()
assert lib.ptmb_tuning_cta_times(t.ctypes.data_as(C.c_void_p), C.c_int(4 * n)) == 0
t = t.reshape(n, 4).astype(np.int64)
t0 = t[:, 0].min()
r = (t - t0) / 1e3
for name, col in (("entry", 0), ("first stage", 1), ("last tile stored", 2), ("exit", 3)):
    c = r[:, col]
    print("%-18s min %7.2f  p50 %7.2f  p90 %7.2f  max %7.2f us" % (name, c.min(), np.percentile(c, 50), np.percentile(c, 90), c.max()))
busy = r[:, 2] - r[:, 1]
print("first stage -> last tile per CTA: min %.2f p50 %.2f max %.2f us" % (busy.min(), np.percentile(busy, 50), busy.max()))
