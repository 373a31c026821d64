#!/usr/bin/env python
"""Write-only HBM ceiling: how fast can 26 GB (the 360-direction raster set of BASELINE.json configs[4]) be
written at all?  cudaMemset-style fill through torch (16-byte stores) and a device-to-device copy for comparison."""
import torch

n = 360 * 6000 * 4000 * 3
out = torch.empty(n, dtype=torch.uint8, device="cuda")


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


ms = timed(lambda: out.zero_())
print("fill  %5.1f GB: %.3f ms = %.0f GB/s written" % (n / 1e9, ms, n / ms / 1e6))
v32 = out.view(torch.int32)
ms = timed(lambda: v32.fill_(0x01020304))
print("fill32 %5.1f GB: %.3f ms = %.0f GB/s written" % (n / 1e9, ms, n / ms / 1e6))
half = n // 2
ms = timed(lambda: out[:half].copy_(out[half:2 * half]))
print("copy  %5.1f GB read + %5.1f GB written: %.3f ms = %.0f GB/s (read + write)" % (half / 1e9, half / 1e9, ms, 2 * half / ms / 1e6))
