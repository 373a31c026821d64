#!/usr/bin/env python
"""DRAM traffic of whole encodes (K1 + hand-off + K2) with the L2 left alone between the kernels.

ncu's per-kernel collection serialises the kernels and writes back / invalidates L2 between them, so it cannot show
what K2 finds in L2 after K1.  This script brackets N back-to-back encodes with cudaProfilerStart/Stop; run it under

    ncu --replay-mode range --cache-control none --clock-control none \
        --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum python tools/range_traffic.py ROWS [N]

and the range's counters are the encodes' real HBM traffic (divide by N)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from rti_b200 import lights as LS  # noqa: E402
from rti_b200.device import U8, light_q14  # noqa: E402
from rti_b200.encoder import PtmEncoder  # noqa: E402
from rti_b200.ptmlib import ptm_svd  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
W = 6000
L = LS.dome1_lights()
M = ptm_svd(L)
P = rows * W
stack = torch.empty((60, rows, W, 3), dtype=torch.uint8, device="cuda")
enc = PtmEncoder(W, rows, M, planes=1)
enc.ctx.synth(0x5EED0001, 0, U8, 0, P, light_q14(L), stack, P * 3)
enc.ctx.set_stack_stable(True)
for _ in range(6):
    enc.encode(stack)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for _ in range(n):
    enc.encode(stack)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("%d encodes of %d x %d x 60: algorithmic %.1f MB each (K1 in %.1f + out %.1f, K2 in %.1f + out %.1f)"
      % (n, W, rows, P * 189 / 1e6, P * 180 / 1e6, P * 27 / 1e6, P * 24 / 1e6, P * 6 / 1e6))
